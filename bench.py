#!/usr/bin/env python
"""bench.py -- point-op pipeline scenes/sec @20k pts (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      # this framework (libgbops, sm_100a)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # the reference's CPU path (oracle port) on host cores

A "step" is one forward+backward pass of the GraspBalance operator pipeline (graspbalance_b200/pipeline.py: SA1-4 with
FPS/gather/ball query/group, 15 InvResMLP groupings, FP1/FP2 and the 20k-point up-sampling, the 16 cylinder-query grasp
crops, the 1024-grasp collision test) over `--batch` synthetic 20k-point scenes per GPU (default 32 = BASELINE config 5).
Scenes shard by batch across ranks with no collective on the op path ("scaling": "weak": per-GPU work is fixed); each
step ends with one NCCL all_gather of the small per-scene outputs.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for how every field is produced.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU job that should use every host thread
    # (numpy's BLAS pool reads the variable when numpy is imported, i.e. below)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point-op pipeline scenes/sec @20k pts"
UNIT = "scenes/s"
N_POINTS = 20000


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in-process (pynvml: two light queries every 50 ms);
    a looping `nvidia-smi` process beside the benchmark was measured to stall the GPU for 60-80 ms now and then (one step
    of ten taking 70-85 ms instead of 13.3), so it is only the fallback when pynvml is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.nvml, self._stop = [], None, gpu_index, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        while not self._stop:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.time(), sm, self.max_sm, [name for name, b in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.strip().split(",")]
            try:
                reasons = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6])
                           if v.lower().startswith("active")]
                self.rows.append((time.time(), float(p[0]), float(p[1]), reasons))
            except Exception:
                continue

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        time.sleep(0.05 if self.nvml is not None else 0.15)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, clk, mxclk, rs in list(self.rows):
            if ts < t0 or ts > t1 + (0.0 if self.nvml is not None else 0.1):
                continue
            sm.append(clk); mx.append(mxclk); reasons.update(rs)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs (host side, pinned)
# ------------------------------------------------------------------------------------------------------------------
def make_host_inputs(scene_ids, pin=True):
    import torch
    from graspbalance_b200 import pipeline, scenes
    from graspbalance_b200.collision_detector import voxel_down_sample_gpu
    B = len(scene_ids)
    xyz = scenes.scene_batch(scene_ids, N_POINTS, "tabletop")
    rot = pipeline.make_view_rotations(B, seed=int(scene_ids[0])).astype(np.float32)
    pts, Ts, Rs, thrs = [], [], [], []
    fw, fl, ad = 0.01, 0.06, 0.03
    for b, sid in enumerate(scene_ids):
        # what ModelFreeCollisionDetector.__init__ does (the product's GPU down-sampling; input preparation, untimed)
        p = voxel_down_sample_gpu(torch.from_numpy(xyz[b].astype(np.float64)).cuda(), 0.01).cpu().numpy()
        g = scenes.grasp_set(int(sid) + 1000, p, pipeline.NUM_GRASP)
        h, d, w = g["heights"][:, None], g["depths"][:, None], g["widths"][:, None]
        thr = np.concatenate([-h / 2, h / 2, d - fl, d, -(w / 2 + fw), -w / 2, (w / 2 + fw), w / 2, d - fl - fw, d - fl - fw - ad], axis=1)
        pts.append(p); Ts.append(g["translations"]); Rs.append(g["rotation_matrices"]); thrs.append(thr)
    offs = np.cumsum([0] + [p.shape[0] for p in pts])
    host = {
        "xyz": torch.from_numpy(xyz),
        "rot": torch.from_numpy(np.ascontiguousarray(rot)),
        "scene_points": torch.from_numpy(np.ascontiguousarray(np.concatenate(pts, axis=0))),
        "T": torch.from_numpy(np.ascontiguousarray(np.stack(Ts))),
        "R": torch.from_numpy(np.ascontiguousarray(np.stack(Rs))),
        "thr": torch.from_numpy(np.ascontiguousarray(np.stack(thrs))),
    }
    if pin:
        host = {k: v.pin_memory() for k, v in host.items()}
    return host, offs


def to_device(host, offs, dev):
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    grasps = {"scene_points": [d["scene_points"][offs[b]:offs[b + 1]] for b in range(len(offs) - 1)],
              "T": d["T"], "R": d["R"], "thr": d["thr"]}
    return d["xyz"], d["rot"], grasps


# ------------------------------------------------------------------------------------------------------------------
# the reference's CPU path: oracle port of the same chain, one scene (used by cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def cpu_pipeline_scene(scene_id, backward=True):
    """One scene through the same op chain on the host cores: oracle/gb_oracle.c (pthreads over independent
    scenes/queries, i.e. every host thread) for the CUDA ops' arithmetic and the whole-array numpy restatement of
    collision_detector.detect for the collision test.  Returns seconds."""
    import oracle
    from graspbalance_b200 import pipeline, scenes
    rng = np.random.default_rng(scene_id)
    xyz = scenes.scene_batch([scene_id], N_POINTS, "tabletop")
    rot = pipeline.make_view_rotations(1, seed=scene_id).reshape(1, pipeline.NUM_SEED, 9)
    pts = oracle.voxel_down_sample(xyz[0].astype(np.float64), 0.01)
    g = scenes.grasp_set(scene_id + 1000, pts, pipeline.NUM_GRASP)
    feats = {}
    for lvl, (m, _, ns, c_in) in enumerate(pipeline.SA_SPECS):
        n_in = N_POINTS if lvl == 0 else pipeline.SA_SPECS[lvl - 1][0]
        if c_in:
            feats[("sa", lvl)] = (rng.normal(size=(1, c_in, n_in)).astype(np.float32), rng.normal(size=(1, c_in, m, ns)).astype(np.float32))
        blocks, c, _, nsb = pipeline.IRM_SPECS[lvl]
        feats[("irm", lvl)] = (rng.normal(size=(1, c, m)).astype(np.float32), rng.normal(size=(1, c, m, nsb)).astype(np.float32))
    fp = [(rng.normal(size=(1, 256, mm)).astype(np.float32), rng.normal(size=(1, 256, nn)).astype(np.float32))
          for (nn, mm) in ((512, 256), (1024, 512), (N_POINTS, 1024))]

    t0 = time.perf_counter()
    cur, levels = xyz, []
    for lvl, (m, radius, ns, c_in) in enumerate(pipeline.SA_SPECS):
        inds = oracle.furthest_point_sample(cur, m, "A")
        cur_t = np.ascontiguousarray(cur.transpose(0, 2, 1))
        new_xyz = np.ascontiguousarray(oracle.gather_operation(cur_t, inds).transpose(0, 2, 1))
        idx = oracle.ball_query(radius, ns, cur, new_xyz)
        gx = oracle.grouping_operation(cur_t, idx)
        gx -= new_xyz.transpose(0, 2, 1)[..., None]
        gx /= radius
        if c_in:
            f, go = feats[("sa", lvl)]
            oracle.grouping_operation(f, idx)
            if backward:
                oracle.grouping_operation_grad(go, idx, f.shape[2])
        blocks, c, r2, nsb = pipeline.IRM_SPECS[lvl]
        f, go = feats[("irm", lvl)]
        new_t = np.ascontiguousarray(new_xyz.transpose(0, 2, 1))
        for _ in range(blocks):
            idx = oracle.ball_query(r2, nsb, new_xyz, new_xyz)
            dp = oracle.grouping_operation(new_t, idx)
            dp = dp - new_xyz.transpose(0, 2, 1)[..., None]
            oracle.grouping_operation(f, idx)
            if backward:
                oracle.grouping_operation_grad(go, idx, m)
        cur = new_xyz
        levels.append(new_xyz)
    for (unk, kn), (f, go) in zip(((levels[2], levels[3]), (levels[1], levels[2]), (xyz, levels[1])), fp):
        dist, idx = oracle.three_nn(unk, kn)
        recip = 1.0 / (dist + 1e-8)
        w = (recip / recip.sum(axis=2, keepdims=True)).astype(np.float32)
        oracle.three_interpolate(f, idx, w)
        if backward:
            oracle.three_interpolate_grad(go, idx, w, f.shape[2])
    seed = levels[1]
    xyz_t = np.ascontiguousarray(xyz.transpose(0, 2, 1))
    rot33 = rot.reshape(1, pipeline.NUM_SEED, 3, 3)
    for r in pipeline.CROP_RADII:
        for hmax in pipeline.CROP_HMAX:
            idx = oracle.cylinder_query(r, pipeline.CROP_HMIN, hmax, 64, xyz, seed, rot)
            gx = oracle.grouping_operation(xyz_t, idx)
            gx -= seed.transpose(0, 2, 1)[..., None]
            np.matmul(gx.transpose(0, 2, 3, 1), rot33)
    oracle.collision_detect_numpy(pts, 0.01, g["translations"], g["rotation_matrices"], g["heights"], g["depths"], g["widths"])
    return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = oracle.num_threads()
    for w in range(args.warmup):
        cpu_pipeline_scene(10_000 + w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_pipeline_scene(20_000 + k)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = "1 scene per step (the full op chain fwd+bwd of one 20k-point scene) out of the 32-scene batch"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "scenes_per_step": 1, "n_points": N_POINTS},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------------------------
# reported baselines beside the headline (rank 0, outside every timed region of the product)
# ------------------------------------------------------------------------------------------------------------------
def _event_ms(fn, iters, warm, dev):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gpu_baseline_leg(args, dev):
    """The reference's own kernels (oracle/_ref = PointNet/_ext_src + pointnet2_batch/src compiled unmodified for sm_100a) on
    this GPU, driven op for op through the same chain with the reference's torch glue (oracle/ref_chain.py), and the
    product on exactly those ops (no collision test: the reference has no GPU collision kernel) at the same batch."""
    import torch
    from graspbalance_b200 import pipeline
    from oracle import ref_chain
    ra, rb, _ = ref_chain.load_refs()
    if ra is None or rb is None:
        return {"unavailable": "oracle/_ref is not built (oracle/build_ref.py needs /root/reference)"}
    Bb = max(1, min(args.batch, args.baseline_batch))
    host, offs = make_host_inputs(list(range(Bb)), pin=False)
    xyz, rot, _ = to_device(host, offs, dev)
    pipe = pipeline.OpPipeline(Bb, N_POINTS, dev, seed=0, backward=not args.no_backward, overlap=not args.no_overlap,
                               fused_crops=not args.unfused_crops)
    with torch.no_grad():
        ms_ref = _event_ms(lambda: ref_chain.run(pipe, xyz, rot, ra, rb, backward=not args.no_backward), 2, 1, dev)
    ms_ours = _event_ms(lambda: pipe.run(xyz, rot, None), 5, 3, dev)
    return {"value": Bb / (ms_ref * 1e-3), "unit": UNIT, "scenes": Bb, "ms_per_step": ms_ref,
            "kind": "reference kernels, sm_100a (oracle/_ref: PointNet/_ext_src + pointnet2_batch/src, unmodified) with the reference's torch glue",
            "product_same_ops": {"value": Bb / (ms_ours * 1e-3), "ms_per_step": ms_ours},
            "excluded": "collision test (numpy on the CPU in the reference)"}


def configs_leg(args, dev):
    """BASELINE configs 1-4 as written (B = 4 scenes): product time, the reference's own implementation beside it."""
    import torch
    import oracle
    from graspbalance_b200 import knn_modules, pipeline, scenes
    from graspbalance_b200 import pointnet2_utils as pu
    from graspbalance_b200.collision_detector import ModelFreeCollisionDetector
    from oracle import ref_chain
    ra, rb, rc = ref_chain.load_refs()
    peak, _ = read_peaks()
    B, N, m, ns = 4, N_POINTS, 1024, 64
    out = {}
    xyz_np = scenes.scene_batch(range(100, 100 + B), N, "tabletop")
    xyz = torch.from_numpy(xyz_np).to(dev)
    xyz_t = xyz.transpose(1, 2).contiguous()
    gen = torch.Generator(device="cpu").manual_seed(7)

    # cfg1: ModelFreeCollisionDetector(scene, voxel_size=0.01).detect(1024 grasps), host arrays in, bool mask out
    pts64 = xyz_np[0].astype(np.float64)
    det = ModelFreeCollisionDetector(pts64, voxel_size=0.01, device=dev)
    gs = scenes.GraspGroupStandIn(**scenes.grasp_set(77, det.scene_points, pipeline.NUM_GRASP))

    def cfg1():
        d = ModelFreeCollisionDetector(pts64, voxel_size=0.01, device=dev)
        return d.detect(gs, approach_dist=0.05, collision_thresh=0.01)
    for _ in range(2):
        cfg1()
    t0 = time.perf_counter()
    for _ in range(5):
        cfg1()
    ms1 = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    det.detect(gs, approach_dist=0.05, collision_thresh=0.01)
    ms1_detect = (time.perf_counter() - t0) * 1e3
    oracle.build()
    t0 = time.perf_counter()
    p_cpu = oracle.voxel_down_sample(pts64, 0.01)
    oracle.collision_detect_numpy(p_cpu, 0.01, gs.translations, gs.rotation_matrices, gs.heights, gs.depths, gs.widths,
                                  approach_dist=0.05, collision_thresh=0.01)
    out["cfg1"] = {"what": "constructor (voxel 0.01) + detect, 20k-point scene, 1024 grasps, host arrays in / mask out",
                   "ms": ms1, "detect_only_ms": ms1_detect, "ref_cpu_ms": (time.perf_counter() - t0) * 1e3,
                   "ref": "numpy restatement of collision_detector.py:16-64 on the host cores"}

    # cfg2: SA chain, B = 4: FPS 20000 -> 1024, ball_query r = 0.05 ns = 64, group C = 3 + 128
    feats = torch.randn((B, 128, N), generator=gen).to(dev)

    def cfg2(mod_fps, mod_gather, mod_ball, mod_group):
        inds = mod_fps(xyz, m)
        new_xyz = mod_gather(xyz_t, inds).transpose(1, 2).contiguous()
        idx = mod_ball(new_xyz, xyz, 0.05, ns)
        return torch.cat([mod_group(xyz_t, idx), mod_group(feats, idx)], dim=1)
    from graspbalance_b200 import _ext as A
    ms2 = _event_ms(lambda: cfg2(A.furthest_point_sampling, A.gather_points, A.ball_query, A.group_points), 5, 2, dev)
    algo2 = B * ((12 * N + 4 * m) + (12 * N + 16 * m) + (12 * N + 12 * m + 4 * m * ns) + 2 * 4 * m * ns + 4 * 131 * N + 4 * 131 * m * ns)
    out["cfg2"] = {"what": "FPS 20000->1024 + gather + ball_query(0.05, 64) + group C=3 and C=128 + cat, B=4", "ms": ms2,
                   "hbm_frac": algo2 / (ms2 * 1e-3) / 1e9 / peak}
    if ra is not None:
        out["cfg2"]["ref_gpu_ms"] = _event_ms(lambda: cfg2(ra.furthest_point_sampling, ra.gather_points, ra.ball_query, ra.group_points), 2, 1, dev)

    # cfg3: FP chain, B = 4: three_nn + weights + three_interpolate forward / backward, 1024 -> 20000, C = 256
    known = pu.gather_operation(xyz_t, pu.furthest_point_sample(xyz, m)).transpose(1, 2).contiguous()
    kf = torch.randn((B, 256, m), generator=gen).to(dev)
    go = torch.randn((B, 256, N), generator=gen).to(dev)

    def cfg3_ours():
        _, idx, w = pu.three_nn_weights(xyz, known)
        A.three_interpolate(kf, idx, w)
        return A.three_interpolate_grad(go, idx, w, m)

    def cfg3_ref():
        d2, idx = ra.three_nn(xyz, known)
        r = 1.0 / (torch.sqrt(d2) + 1e-8)
        w = r / torch.sum(r, dim=2, keepdim=True)
        ra.three_interpolate(kf, idx, w)
        return ra.three_interpolate_grad(go, idx, w, m)
    ms3 = _event_ms(cfg3_ours, 5, 2, dev)
    algo3 = B * ((12 * N + 12 * m + 36 * N) + 2 * (4 * 256 * m + 24 * N + 4 * 256 * N))
    out["cfg3"] = {"what": "three_nn + weights + three_interpolate fwd + bwd, 1024 -> 20000, C=256, B=4", "ms": ms3,
                   "hbm_frac": algo3 / (ms3 * 1e-3) / 1e9 / peak}
    if ra is not None:
        out["cfg3"]["ref_gpu_ms"] = _event_ms(cfg3_ref, 2, 1, dev)

    # cfg4: grasp-crop stage, B = 4: 12 approach-view sets x 4 depths = 48 cylinder queries + grouped coordinates,
    # KNN k = 1 and 64 (R = 20000, Q = 1024), and the batched collision test
    seeds = known
    rng = np.random.default_rng(3)
    views = rng.normal(size=(B, m, 3)).astype(np.float32)
    rots = [torch.from_numpy(np.ascontiguousarray(scenes.viewpoint_rotations(-views, np.full((B, m), i * np.pi / 12, np.float32))
                                                  .reshape(B, m, 9))).to(dev) for i in range(12)]
    ref_cf, qry_cf = xyz_t, seeds.transpose(1, 2).contiguous()

    def crops(query, group, fused):
        for rot in rots:
            if fused:
                idx = pu.cylinder_query_multi(0.05, -0.02, pipeline.CROP_HMAX, ns, xyz, seeds, rot)
                pu._FusedQueryGroup.apply(xyz, seeds, idx.view(B, m, 4 * ns), rot, None, None)
                continue
            for hmax in pipeline.CROP_HMAX:
                idx = query(seeds, xyz, rot, 0.05, -0.02, hmax, ns)
                g = group(xyz_t, idx)
                g -= seeds.transpose(1, 2).unsqueeze(-1)
                torch.matmul(g.permute(0, 2, 3, 1).contiguous(), rot.view(B, m, 3, 3)).permute(0, 3, 1, 2).contiguous()

    def crops_ours_unfused():
        for rot in rots:
            for hmax in pipeline.CROP_HMAX:
                idx = A.cylinder_query(seeds, xyz, rot, 0.05, -0.02, hmax, ns)
                pu._FusedQueryGroup.apply(xyz, seeds, idx, rot, None, None)

    def knn_ours():
        knn_modules.knn_k(ref_cf, qry_cf, 1)
        knn_modules.knn_k(ref_cf, qry_cf, 64)

    def knn_ref():
        for k in (1, 64):
            idx = torch.empty((B, k, m), dtype=torch.int64, device=dev)
            rc.knn(ref_cf, qry_cf, idx)
    ms4c = _event_ms(crops_ours_unfused, 3, 1, dev)
    ms4f = _event_ms(lambda: crops(None, None, True), 3, 1, dev)
    ms4k = _event_ms(knn_ours, 3, 1, dev)
    out["cfg4"] = {"what": "48 cylinder_query(0.05, -0.02, hmax, 64) + grouped rotated coordinates; knn k=1 and k=64 (R=20000, Q=1024); B=4",
                   "crops_48_calls_ms": ms4c, "crops_fused_depths_ms": ms4f, "knn_ms": ms4k, "ms": ms4f + ms4k}
    if ra is not None:
        out["cfg4"]["ref_gpu_crops_ms"] = _event_ms(lambda: crops(ra.cylinder_query, ra.group_points, False), 1, 1, dev)
    if rc is not None:
        out["cfg4"]["ref_gpu_knn_ms"] = _event_ms(knn_ref, 1, 1, dev)
    if "ref_gpu_crops_ms" in out["cfg4"] and "ref_gpu_knn_ms" in out["cfg4"]:
        out["cfg4"]["ref_gpu_ms"] = out["cfg4"]["ref_gpu_crops_ms"] + out["cfg4"]["ref_gpu_knn_ms"]
    return out


def arithmetic_peaks():
    """Measured DFMA / FFMA lane-operations per second (tests/ubench/peaks.cu), or None when the helper is not built."""
    import ctypes
    path = os.path.join(ROOT, "tests", "ubench", "libgb_peaks.so")
    if not os.path.exists(path):
        return None
    L = ctypes.CDLL(path)
    L.ub_dfma_per_s.restype = L.ub_ffma_per_s.restype = ctypes.c_double
    L.ub_dfma_per_s.argtypes = L.ub_ffma_per_s.argtypes = [ctypes.c_int]
    return {"dfma_per_s": float(L.ub_dfma_per_s(4096)), "ffma_per_s": float(L.ub_ffma_per_s(8192)),
            "how": "tests/ubench/peaks.cu: 8 independent FMA chains per thread, 2 x 1024 threads per SM, best of 3"}


# lane instructions per candidate test of the scan kernels (sub/mul/fma/compare of the reference arithmetic)
SCAN_INSTR = {"gb_ball_query": 8, "gb_cylinder_query": 17, "gb_three_nn": 12, "gb_three_nn_weights": 12, "gb_knn": 8}


def kernel_notes(prof, peaks, steps):
    """What bounds the kernels that are not HBM-bound: FPS = latency of m - 1 dependent rounds (us per round);
    scans = fp32 issue (full-scan-equivalent candidate tests per second against the measured FFMA rate: above 1 means the
    cell grid culled candidates a full scan would test); collision = FP64 pipe against the measured DFMA rate."""
    notes = {}
    for name, evs in prof.items():
        if name.startswith("gb_fps"):
            best = max(evs, key=lambda e: e[2])
            a = best[3]
            mm = a[6] if name == "gb_fps_xyz" else a[5]
            rows = [e for e in evs if e[3] == a]
            us = sum(x.elapsed_time(y) for x, y, _, _ in rows) / len(rows) * 1e3
            notes["fps"] = {"launch": f"b={a[4] if name == 'gb_fps_xyz' else a[3]} n={a[5] if name == 'gb_fps_xyz' else a[4]} m={mm}",
                            "us": us, "us_per_round": us / max(mm - 1, 1),
                            "sm": "ncu (profiles/r01e_ncu_full_summary.csv): 2 warps per scheduler, issue slots 45 % of a round, the rest is the cluster exchange latency"}
    if peaks:
        for name, evs in prof.items():
            base = name.replace("_multi_radius", "").replace("_multi", "").replace("_batched", "")
            if base in SCAN_INSTR:
                tests = 0.0
                for _, _, _, a in evs:
                    if base in ("gb_ball_query",):
                        tests += a[3] * a[4] * a[5]
                    elif base == "gb_cylinder_query":
                        tests += a[4] * a[5] * a[6]
                    elif base in ("gb_three_nn",):
                        tests += a[4] * a[5] * a[6]
                    elif base == "gb_three_nn_weights":
                        tests += a[5] * a[6] * a[7]
                    elif base == "gb_knn":
                        tests += a[3] * a[5] * a[6]
                sec = sum(x.elapsed_time(y) for x, y, _, _ in evs) * 1e-3
                n = notes.setdefault(base, {"tests_per_s": 0.0})
                n["tests_per_s"] = tests / sec if sec > 0 else 0.0
                n["fp32_issue_frac_full_scan_equiv"] = n["tests_per_s"] * SCAN_INSTR[base] / peaks["ffma_per_s"]
            if name.startswith("gb_collision_counts"):
                pairs = sum((a[2] * a[3] * a[7]) if name.endswith("batched") else (a[1] * a[5]) for _, _, _, a in evs)
                sec = sum(x.elapsed_time(y) for x, y, _, _ in evs) * 1e-3
                notes["collision"] = {"pairs_per_s_exhaustive_equiv": pairs / sec if sec > 0 else 0.0,
                                      "fp64_pipe_frac_exhaustive_equiv": (pairs / sec) * 24 / peaks["dfma_per_s"] if sec > 0 else 0.0,
                                      "ops_per_pair": "3 DADD + 9 DMUL/DFMA + 12 DSETP (collision_detector.py:23-41) if every grasp x point pair were tested; "
                                                      "above 1 = the pack-bounds culling and the height-slab reject skipped that share of the arithmetic"}
    return notes


def workload_name(args):
    return (f"BASELINE config 5: full GraspBalance backbone op pipeline forward{'+backward' if not args.no_backward else ''} "
            f"(SA1-4 + 15 InvResMLP + FP1/2 + 20k up-sampling + 16 cylinder crops + 1024-grasp collision), "
            f"{args.batch} synthetic 20k-point scenes per GPU")


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gbops", choices=["gbops", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="scenes per GPU per step")
    ap.add_argument("--no-backward", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="every step runs its own sampling chain first (no cross-step pipelining)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (32 scenes in total over the N GPUs)")
    ap.add_argument("--strong-total", type=int, default=32, help="scenes in total of the strong-scaling leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-kernels leg (oracle/_ref on this GPU)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 1-4 leg")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE", help="libgbops tuning knob for experiments (gb_set_tuning)")
    ap.add_argument("--baseline-batch", type=int, default=8, help="scenes per step of the reference-kernels leg")
    ap.add_argument("--no-overlap", action="store_true", help="run the sampling chain and the collision tests on the main stream")
    ap.add_argument("--unfused-crops", action="store_true",
                    help="grasp crops as the reference's 16 separate CylinderQueryAndGroup calls instead of 4 multi-depth scans")
    ap.add_argument("--cuda-profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gbops" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from graspbalance_b200 import _lib, pipeline, sharding

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: graspbalance_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    _lib.lib()  # fail loudly now if libgbops.so is missing
    for kv in args.tune:
        key, _, val = kv.partition("=")
        _lib.set_tuning(key, int(val))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    scene_ids = sharding.scene_ids_for_rank(rank, world, B)
    host, offs = make_host_inputs(scene_ids)
    pipe = pipeline.OpPipeline(B, N_POINTS, dev, seed=rank, backward=not args.no_backward, overlap=not args.no_overlap,
                               fused_crops=not args.unfused_crops)
    gather_buf = torch.empty((world * B, pipeline.NUM_SEED + 6 * pipeline.NUM_GRASP), dtype=torch.int64, device=dev) if world > 1 else None

    # cross-step pipelining of the sampling chain (4 x FPS + gather: ~2 ms of dependent rounds that every other operator of a
    # step waits for, but which needs the coordinates only): step k launches the chain of step k + 1 on the sampling stream
    # beside its own grouping work, into the other of two sample-buffer sets.  Every step still runs exactly one chain.
    pf = {"on": not args.no_prefetch, "k": 0, "bufs": [pipe.alloc_samples(), pipe.alloc_samples()]}

    def prime_prefetch(xyz):
        pf["k"] = 0
        pipe.sampling_chain(xyz, pf["bufs"][0])

    def step(resident, inputs=None, next_xyz=None, next_ready=None):
        """One pipeline pass over device-resident inputs; returns the per-scene result tensor and the checksums."""
        xyz, rot, grasps = inputs
        if pf["on"]:
            k = pf["k"]
            pf["k"] = k + 1
            out = pipe.run(xyz, rot, grasps, samples=pf["bufs"][k % 2],
                           prefetch=(xyz if next_xyz is None else next_xyz, pf["bufs"][(k + 1) % 2], next_ready))
        else:
            out = pipe.run(xyz, rot, grasps)
        result = torch.cat([out["seed_inds"].to(torch.int64), out["collision_counts"].reshape(B, -1)], dim=1)
        if world > 1:
            sharding.gather_scene_outputs(result, world, gather_buf)  # NCCL: gather per-scene outputs only
        if resident:
            return result
        chk = torch.stack([out["up_checksum"], out["crop_checksum"]] + ([out["grad_checksum"]] if "grad_checksum" in out else []))
        return result, chk

    # ---- e2e: every step copies its inputs from pinned host memory and its results back to pinned host memory.  The copies
    # run on a copy stream one step ahead / behind the compute stream (two device input sets, two host result sets), so a
    # step's H2D overlaps the previous step's kernels; every copy is inside the timed region and every step's result is on
    # the host when the region closes.
    copy_stream = torch.cuda.Stream(dev)
    e2e_bufs = []
    e2e_state = {"k": 0, "ready": None}

    def e2e_prepare():
        for _ in range(2):
            d = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
            grasps = {"scene_points": [d["scene_points"][offs[b]:offs[b + 1]] for b in range(len(offs) - 1)],
                      "T": d["T"], "R": d["R"], "thr": d["thr"]}
            e2e_bufs.append({"dev": d, "inputs": (d["xyz"], d["rot"], grasps), "free": None, "res_h": None, "chk_h": None})

    def e2e_steps(n_steps):
        main = torch.cuda.current_stream(dev)

        def upload(k):
            buf = e2e_bufs[k % 2]
            with torch.cuda.stream(copy_stream):
                if buf["free"] is not None:
                    copy_stream.wait_event(buf["free"])  # the step that last read this input set has finished
                for name, v in host.items():
                    buf["dev"][name].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        # A continuous stream of steps: every call picks up where the last one stopped, so a timed call of K steps holds exactly
        # K uploads, K sampling chains, K steps and K downloads (the first call primes step 0's upload and chain).
        st = e2e_state
        if st["ready"] is None:
            # the device input sets were allocated on the compute stream's pool: whatever pending compute-stream work last used
            # those blocks must finish before the copy stream writes into them
            copy_stream.wait_stream(main)
            st["ready"] = upload(0)
            if pf["on"]:
                main.wait_event(st["ready"])
                prime_prefetch(e2e_bufs[0]["inputs"][0])  # the first step's chain (every later one is launched a step ahead)
        for k in range(st["k"], st["k"] + n_steps):
            buf = e2e_bufs[k % 2]
            main.wait_event(st["ready"])
            st["ready"] = upload(k + 1)
            nxt = e2e_bufs[(k + 1) % 2]["inputs"][0]
            result, chk = step(False, buf["inputs"], nxt, st["ready"])
            done = torch.cuda.Event()
            done.record(main)
            buf["free"] = done
            with torch.cuda.stream(copy_stream):  # results of step k leave while step k + 1 computes
                copy_stream.wait_event(done)
                if buf["res_h"] is None:
                    buf["res_h"] = torch.empty(result.shape, dtype=result.dtype, pin_memory=True)
                    buf["chk_h"] = torch.empty(chk.shape, dtype=chk.dtype, pin_memory=True)
                buf["res_h"].copy_(result, non_blocking=True)
                buf["chk_h"].copy_(chk, non_blocking=True)
                result.record_stream(copy_stream), chk.record_stream(copy_stream)
        st["k"] += n_steps
        main.wait_stream(copy_stream)  # the timed region ends when the last result is on the host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, resident, inputs=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        marks = []
        if callable(resident):
            resident(n_steps)  # a whole-loop runner (the e2e path, graph replays)
        else:
            for _ in range(n_steps):
                step(resident, inputs)
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()  # per-step marks on the main stream: diagnostics only (the value is e0..e1 over all K steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        timed.per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)] if marks else [ms / n_steps]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t_wall0, time.time()

    resident_inputs = to_device(host, offs, dev)
    torch.cuda.synchronize()
    prime_prefetch(resident_inputs[0])
    for _ in range(args.warmup):
        step(True, resident_inputs)
    # allocator priming (untimed, on top of the W warm-up steps): the caching allocator keeps growing its pools for a few
    # steps because blocks handed to side streams are recycled late; a cudaMalloc inside the timed region costs
    # milliseconds.  Step until one whole step needs no new device allocation (at most 8 extra steps).
    # ... and leave it slack: a block as large as everything reserved so far (the allocator splits it on demand) plus 128 MB
    # of small-pool blocks.  A cudaMalloc needs the driver's resource-manager lock; when an NVML / nvidia-smi clock query
    # holds that lock at the same moment the step stalls for 40-120 ms (measured: one step of ten at 70-128 ms instead of
    # 13.3 ms in a third of the runs).  Kernel launches do not take the lock, so with no allocation inside the timed region
    # the clock sampling is harmless.
    main_line = getattr(pipe, "_main_stream", None) or torch.cuda.current_stream(dev)  # the stream the step's big tensors come from
    with torch.cuda.stream(main_line):
        slack = [torch.empty(max(torch.cuda.memory_reserved(dev), 1 << 30), dtype=torch.uint8, device=dev)]
        slack += [torch.empty(64 << 10, dtype=torch.uint8, device=dev) for _ in range(2048)]
    if main_line is not torch.cuda.current_stream(dev):
        slack += [torch.empty(1 << 30, dtype=torch.uint8, device=dev)]
        slack += [torch.empty(64 << 10, dtype=torch.uint8, device=dev) for _ in range(512)]
    for st in (getattr(pipe, n, None) for n in ("_fps_stream", "_col_stream", "_aux_stream")):  # pools are per stream
        if st is not None:
            with torch.cuda.stream(st):
                slack += [torch.empty(2 << 30, dtype=torch.uint8, device=dev)]
                slack += [torch.empty(64 << 10, dtype=torch.uint8, device=dev) for _ in range(512)]
    torch.cuda.synchronize()
    del slack
    prime_steps, quiet = 0, 0
    while prime_steps < 10 and quiet < 3:  # three steps in a row without a new device allocation
        before = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        step(True, resident_inputs)
        torch.cuda.synchronize()
        prime_steps += 1
        quiet = quiet + 1 if torch.cuda.memory_stats(dev).get("num_device_alloc", 0) == before else 0
    dev_allocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    if args.cuda_profiler_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms, tw0, tw1 = timed(args.steps, True, resident_inputs)
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    launches = _lib.launch_count() - launches0
    dev_allocs_timed = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - dev_allocs0
    per_step = sorted(timed.per_step)
    mstats = torch.cuda.memory_stats(dev)
    step_diag = {"min_ms": per_step[0], "median_ms": per_step[len(per_step) // 2], "max_ms": per_step[-1],
                 "reserved_gb": mstats.get("reserved_bytes.all.current", 0) / 1e9, "device_frees": mstats.get("num_device_free", 0),
                 "alloc_retries": mstats.get("num_alloc_retries", 0)}
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- the same K steps without the cross-step pipelining (every step runs its own sampling chain first) ----
    no_pf = None
    if pf["on"]:
        pf["on"] = False
        for _ in range(2):
            step(True, resident_inputs)
        ms_np, _, _ = timed(args.steps, True, resident_inputs)
        no_pf = {"ms_per_step": ms_np / args.steps, "value": world * B * args.steps / (ms_np * 1e-3), "unit": UNIT}
        pf["on"] = True

    # ---- roofline pass: the same K steps again with CUDA events around every libgbops call, on ONE stream (the side
    # streams of the overlapped schedule would make the bracketed durations overlap each other) ----
    overlap, pipe.overlap, pf_on, pf["on"] = pipe.overlap, False, pf["on"], False
    with torch.cuda.stream(main_line):  # the stream whose memory pool holds the step's tensors: no device allocation in this pass either
        step(True, resident_inputs)  # untimed: the one-stream schedule draws the side streams' tensors from this pool too
        torch.cuda.synchronize()
        _lib.PROFILER = {}
        ms_prof, _, _ = timed(args.steps, True, resident_inputs)
        prof, _lib.PROFILER = _lib.PROFILER, None
    pipe.overlap, pf["on"] = overlap, pf_on
    prime_prefetch(resident_inputs[0])

    # ---- e2e: host buffers in, results out, every step ----
    e2e = None
    if not args.no_e2e:
        e2e_prepare()
        e2e_steps(3)
        ms_e2e, _, _ = timed(args.steps, e2e_steps)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = B * (pipeline.NUM_SEED + 6 * pipeline.NUM_GRASP) * 8 + 3 * 4
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
               "copies": "pinned host -> device and results -> pinned host every step, on a copy stream one step ahead / behind"}


    # ---- strong scaling (BASELINE config 5 as written: 32 scenes in total, sharded 32 / N per GPU) ----
    # A small shard makes the step launch-bound on the host (about 170 launches from Python), so the whole step -- every
    # stream, forward and backward -- is captured once in a CUDA graph and replayed; the all_gather stays an eager call.
    strong = None
    if not args.no_strong:
        total = args.strong_total
        lo, hi = sharding.shard_batch(total, world)[rank]
        Bs = hi - lo
        if Bs == B:
            s_pipe, s_inputs = pipe, resident_inputs
        else:
            s_host, s_offs = make_host_inputs(list(range(lo, hi)), pin=False)
            s_pipe = pipeline.OpPipeline(Bs, N_POINTS, dev, seed=rank, backward=not args.no_backward, overlap=not args.no_overlap,
                                         fused_crops=not args.unfused_crops)
            s_inputs = to_device(s_host, s_offs, dev)
        s_gather = torch.empty((total, pipeline.NUM_SEED + 6 * pipeline.NUM_GRASP), dtype=torch.int64, device=dev) if world > 1 else None
        equal_shards = all(e - b == Bs for b, e in sharding.shard_batch(total, world))

        s_bufs = [s_pipe.alloc_samples(), s_pipe.alloc_samples()]

        def s_step(k):
            if args.no_prefetch:
                out = s_pipe.run(*s_inputs)
            else:  # graph k % 2 consumes sample set k % 2 and fills the other one for the next replay
                out = s_pipe.run(*s_inputs, samples=s_bufs[k % 2], prefetch=(s_inputs[0], s_bufs[(k + 1) % 2]))
            return torch.cat([out["seed_inds"].to(torch.int64), out["collision_counts"].reshape(Bs, -1)], dim=1)

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            s_pipe.sampling_chain(s_inputs[0], s_bufs[0])
            for k in range(4):
                s_step(k)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graphs, s_results = [], []
        launches_g0 = _lib.launch_count()
        for k in range(1 if args.no_prefetch else 2):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=side):
                s_results.append(s_step(k))
            graphs.append(g_)
        launches_graph = (_lib.launch_count() - launches_g0) // len(graphs)
        torch.cuda.synchronize()

        def s_loop(n_steps):
            for k in range(n_steps):
                graphs[k % len(graphs)].replay()
                if world > 1 and equal_shards:
                    sharding.gather_scene_outputs(s_results[k % len(graphs)], world, s_gather)

        s_loop(max(args.warmup, 3))
        ms_s, _, _ = timed(args.steps, s_loop)
        strong = {"scenes_total": total, "scenes_per_gpu": Bs, "ms_per_step": ms_s / args.steps,
                  "value": total * args.steps / (ms_s * 1e-3), "unit": UNIT, "cuda_graph": True,
                  "launches_per_step": int(launches_graph),
                  "note": "same step as `value`, one CUDA-graph replay per step; efficiency = value(N) / value(1) across the driver's runs"}
        strong["prefetch"] = not args.no_prefetch
        del graphs

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (largest share of the summed launch time) ----
    peak, peak_src = read_peaks()
    fam = {}
    for name, evs in prof.items():
        tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in evs)
        base = name.replace("gb_group_xyz_feat", "gb_group_fwd").replace("_set", "").replace("_strided", "").replace("_multi_radius", "").replace("_multi", "").replace("_batched", "")  # entry-point variants of one op share its kernels
        f = fam.setdefault(base, {"launches": 0, "ms": 0.0, "bytes": 0, "big": None})
        f["launches"] += len(evs)
        f["ms"] += tot_ms
        f["bytes"] += sum(x for _, _, x, _ in evs)
        for a, b, x, _ in evs:  # the launch that moves the most bytes: small launches of a family are latency-bound
            if f["big"] is None or x > f["big"][0]:
                f["big"] = (x, a.elapsed_time(b))
    total_ms = sum(f["ms"] for f in fam.values()) or 1.0
    per_op = []
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0.0
        big_gbs = f["big"][0] / (f["big"][1] * 1e-3) / 1e9 if f["big"] and f["big"][1] > 0 else 0.0
        per_op.append({"kernel": name, "launches_per_step": f["launches"] / args.steps, "ms_per_step": f["ms"] / args.steps,
                       "share": f["ms"] / total_ms, "achieved_gbs": gbs, "hbm_frac": gbs / peak,
                       "largest_launch": {"algorithmic_bytes": int(f["big"][0]) if f["big"] else 0,
                                          "us": f["big"][1] * 1e3 if f["big"] else 0.0, "hbm_frac": big_gbs / peak}})
    top = per_op[0]
    # dram__bytes_read + dram__bytes_write per launch of that family, from the committed ncu capture of this same step
    # (profiles/dram_traffic.json, written by tests/ubench/dram_traffic.py); null when the capture is for another batch size
    traffic, algo_per_launch = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("batch") == B and tj.get("backward") == (not args.no_backward):
            traffic = tj["families"].get(top["kernel"], {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    if top["launches_per_step"] > 0:
        algo_per_launch = fam[top["kernel"]]["bytes"] / fam[top["kernel"]]["launches"]
    roofline = {"kernel": top["kernel"], "bound": "hbm", "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["achieved_gbs"] / peak, "traffic": traffic, "algorithmic_bytes_per_launch": algo_per_launch,
                "peak_source": peak_src,
                "avg_launch_us": top["ms_per_step"] / max(top["launches_per_step"], 1e-9) * 1e3, "share_of_step": top["share"]}

    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cpu_pipeline_scene(9_999)  # warm the page cache / thread pool
        n_cpu = 2
        t_cpu = sum(cpu_pipeline_scene(30_000 + i) for i in range(n_cpu))
        cpu = {"value": n_cpu / t_cpu, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
               "sample": f"{n_cpu} scenes (full op chain fwd+bwd each) of the {B}-scene batch, oracle/gb_oracle.c + numpy detect"}

    peaks = arithmetic_peaks()
    notes = kernel_notes(prof, peaks, args.steps)
    gpu_base = None if args.no_gpu_baseline else gpu_baseline_leg(args, dev)
    cfgs = None if args.no_configs else configs_leg(args, dev)

    algo = pipeline.algorithmic_bytes_per_scene(N_POINTS, not args.no_backward)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "scenes_per_gpu": B, "n_points": N_POINTS, "parallelism": f"scene-sharded x{world}",
                       "headline": "`value` / `e2e` = 32 scenes per GPU (weak scaling: the configuration that fills a B200); `strong` = BASELINE config 5 as written, 32 scenes in total over the N GPUs",
                       "streams": ("sampling chain, collision tests, crops + interpolation on side streams; the main line (group forward / backward) on a stream of higher priority"
                                   if pipe.overlap else "single stream"),
                       "sampling": ("the sampling chain of step k + 1 runs beside step k (it needs the coordinates only); `no_prefetch` = every step samples first"
                                    if pf["on"] else "every step runs its own sampling chain first"),
                       "l2": "per-step working set (>10 GB of grouped features) exceeds the 126 MB L2; no explicit flush",
                       "algorithmic_bytes_per_scene": int(sum(algo.values())),
                       "allocator_priming_steps": prime_steps, "device_allocs_in_timed_region": int(dev_allocs_timed),
                       "step_diagnostics": step_diag},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "no_prefetch": no_pf, "strong": strong, "gpu_baseline": gpu_base, "configs": cfgs, "arithmetic_peaks": peaks, "kernel_notes": notes,
            "roofline_pass_ms_per_step": ms_prof / args.steps,
            "pipeline_hbm_frac": (sum(algo.values()) * world * B * args.steps / (ms * 1e-3) / 1e9) / (peak * world),
            "per_op": per_op}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
