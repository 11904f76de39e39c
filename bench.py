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
            time.sleep(0.05)

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
    from graspbalance_b200.collision_detector import voxel_down_sample
    B = len(scene_ids)
    xyz = scenes.scene_batch(scene_ids, N_POINTS, "tabletop")
    rot = pipeline.make_view_rotations(B, seed=int(scene_ids[0])).astype(np.float32)
    pts, Ts, Rs, thrs = [], [], [], []
    fw, fl, ad = 0.01, 0.06, 0.03
    for b, sid in enumerate(scene_ids):
        p = voxel_down_sample(xyz[b].astype(np.float64), 0.01)  # what ModelFreeCollisionDetector.__init__ does (host)
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


def workload_name(args):
    return (f"BASELINE config 5: full GraspBalance backbone op pipeline forward{'+backward' if not args.no_backward else ''} "
            f"(SA1-4 + 15 InvResMLP + FP1/2 + 20k up-sampling + 16 cylinder crops + 1024-grasp collision), "
            f"{args.batch} synthetic 20k-point scenes per GPU")


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gbops", choices=["gbops", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="scenes per GPU per step")
    ap.add_argument("--no-backward", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    def step(resident, inputs=None):
        """One pipeline pass.  resident=True: inputs already on the device.  resident=False: the e2e path -- pinned host
        buffers are copied in, the per-scene results are copied back to the host."""
        if resident:
            xyz, rot, grasps = inputs
        else:
            xyz, rot, grasps = to_device(host, offs, dev)
        out = pipe.run(xyz, rot, grasps)
        result = torch.cat([out["seed_inds"].to(torch.int64), out["collision_counts"].reshape(B, -1)], dim=1)
        if world > 1:
            sharding.gather_scene_outputs(result, world, gather_buf)  # NCCL: gather per-scene outputs only
        if not resident:
            res_h = result.to("cpu", non_blocking=True)
            chk_h = torch.stack([out["up_checksum"], out["crop_checksum"]] + ([out["grad_checksum"]] if "grad_checksum" in out else [])).to("cpu", non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return res_h, chk_h
        return result

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
        for _ in range(n_steps):
            step(resident, inputs)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()  # per-step marks on the main stream: diagnostics only (the value is e0..e1 over all K steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        timed.per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t_wall0, time.time()

    resident_inputs = to_device(host, offs, dev)
    torch.cuda.synchronize()
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
    slack = [torch.empty(max(torch.cuda.memory_reserved(dev), 1 << 30), dtype=torch.uint8, device=dev)]
    slack += [torch.empty(64 << 10, dtype=torch.uint8, device=dev) for _ in range(2048)]
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

    # ---- roofline pass: the same K steps again with CUDA events around every libgbops call, on ONE stream (the side
    # streams of the overlapped schedule would make the bracketed durations overlap each other) ----
    _lib.PROFILER = {}
    overlap, pipe.overlap = pipe.overlap, False
    ms_prof, _, _ = timed(args.steps, True, resident_inputs)
    pipe.overlap = overlap
    prof, _lib.PROFILER = _lib.PROFILER, None

    # ---- e2e: host buffers in, results out, every step ----
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step(False)
        ms_e2e, _, _ = timed(args.steps, False)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = B * (pipeline.NUM_SEED + 6 * pipeline.NUM_GRASP) * 8 + 3 * 4
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (largest share of the summed launch time) ----
    peak, peak_src = read_peaks()
    fam = {}
    for name, evs in prof.items():
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
        base = name.replace("_set", "").replace("_strided", "").replace("_multi_radius", "").replace("_multi", "").replace("_batched", "")  # entry-point variants of one op share its kernels
        f = fam.setdefault(base, {"launches": 0, "ms": 0.0, "bytes": 0, "big": None})
        f["launches"] += len(evs)
        f["ms"] += tot_ms
        f["bytes"] += sum(x for _, _, x in evs)
        for a, b, x in evs:  # the launch that moves the most bytes: small launches of a family are latency-bound
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

    algo = pipeline.algorithmic_bytes_per_scene(N_POINTS, not args.no_backward)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "scenes_per_gpu": B, "n_points": N_POINTS, "parallelism": f"scene-sharded x{world}",
                       "streams": "sampling chain + collision tests on side streams" if pipe.overlap else "single stream",
                       "l2": "per-step working set (>10 GB of grouped features) exceeds the 126 MB L2; no explicit flush",
                       "algorithmic_bytes_per_scene": int(sum(algo.values())),
                       "allocator_priming_steps": prime_steps, "device_allocs_in_timed_region": int(dev_allocs_timed),
                       "step_diagnostics": step_diag},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "roofline_pass_ms_per_step": ms_prof / args.steps,
            "pipeline_hbm_frac": (sum(algo.values()) * world * B * args.steps / (ms * 1e-3) / 1e9) / (peak * world),
            "per_op": per_op}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
