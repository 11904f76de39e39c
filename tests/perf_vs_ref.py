#!/usr/bin/env python
"""Per-op timing of libgbops against the reference's own CUDA extensions (oracle/_ref) on the same B200.

Test infrastructure (lives under tests/, may load oracle/_ref): run on the GPU box as
    python tests/perf_vs_ref.py [--out gpurun_out/perf_vs_ref.json] [--sweep]
Every op is timed with CUDA events on the launching stream over ITER launches after WARM warm-up launches, with a
256 MB write between launches to flush L2 where the working set is smaller than L2 (flag `flush`).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import _load_ref  # noqa: E402
from graspbalance_b200 import _ext as A, _lib, knn_modules, pointnet2_batch_cuda as Bm, scenes  # noqa: E402
from graspbalance_b200.collision_detector import ModelFreeCollisionDetector, collision_counts  # noqa: E402

HBM = 6542.4  # GB/s, MEASURED_PEAKS.json
dev = torch.device("cuda:0")
_flush_buf = None


def flush_l2():
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    _flush_buf.fill_(1)


def timeit(fn, iters=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]  # median, microseconds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "perf_vs_ref.json"))
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--sweep2", action="store_true", help="cell-grid edge and sorted-backward channel sweeps")
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--only", default="", help="regex: time only the ops whose name matches")
    ap.add_argument("--iters", type=int, default=0, help="override the number of timed launches per op")
    ap.add_argument("--no-ref", action="store_true", help="do not time the reference extensions")
    args = ap.parse_args()
    import re
    only = re.compile(args.only) if args.only else None
    rA, rB, rC = _load_ref("gbref_pointnet2_ext"), _load_ref("gbref_pointnet2_batch"), _load_ref("gbref_knn")
    if args.no_ref:
        rA = rB = rC = None
    B, N, m, ns = args.B, 20000, 1024, 64
    rows = []

    def add(name, mine, ref_a=None, ref_b=None, bytes_per_scene=None, scenes_n=B, **kw):
        if only is not None and not only.search(name):
            return
        if args.iters:
            kw["iters"] = args.iters
        t = timeit(mine, **kw)
        row = {"op": name, "us": round(t, 2), "us_per_scene": round(t / scenes_n, 2)}
        if bytes_per_scene:
            gbs = bytes_per_scene * scenes_n / (t * 1e-6) / 1e9
            row.update(algo_bytes_per_scene=bytes_per_scene, gbs=round(gbs, 1), hbm_frac=round(gbs / HBM, 4))
        if ref_a is not None and rA is not None:
            row["ref_a_us"] = round(timeit(ref_a, **kw), 2)
            row["speedup_vs_a"] = round(row["ref_a_us"] / t, 2)
        if ref_b is not None and rB is not None:
            row["ref_b_us"] = round(timeit(ref_b, **kw), 2)
            row["speedup_vs_b"] = round(row["ref_b_us"] / t, 2)
        rows.append(row)
        print(json.dumps(row), flush=True)

    xyz = torch.from_numpy(scenes.scene_batch(range(B), N, "tabletop")).to(dev)
    xyz_t = xyz.transpose(1, 2).contiguous()
    g = torch.Generator(device="cpu").manual_seed(0)

    # ---- cfg2: SA chain ----
    for mm in (1024, 2048):
        temp = torch.full((B, N), 1e10, device=dev)
        out = torch.empty((B, mm), dtype=torch.int32, device=dev)
        add(f"fps {N}->{mm} (A)", lambda: A.furthest_point_sampling(xyz, mm),
            (lambda: rA.furthest_point_sampling(xyz, mm)) if rA else None,
            (lambda: (temp.fill_(1e10), rB.furthest_point_sampling_wrapper(B, N, mm, xyz, temp, out))) if rB else None,
            bytes_per_scene=12 * N + 4 * mm, iters=10)
    fidx = A.furthest_point_sampling(xyz, m)
    new_xyz = A.gather_points(xyz_t, fidx).transpose(1, 2).contiguous()
    for nn, mm in ((2048, 1024), (1024, 512), (512, 256)):
        sub = xyz[:, :nn].contiguous()
        add(f"fps {nn}->{mm} (A)", lambda: A.furthest_point_sampling(sub, mm),
            (lambda: rA.furthest_point_sampling(sub, mm)) if rA else None, bytes_per_scene=12 * nn + 4 * mm, iters=10)
    add("gather C=3", lambda: A.gather_points(xyz_t, fidx), (lambda: rA.gather_points(xyz_t, fidx)) if rA else None,
        bytes_per_scene=4 * 3 * N + 4 * m + 4 * 3 * m)
    idxb = torch.zeros((B, m, ns), dtype=torch.int32, device=dev)
    add("ball_query r=.05 ns=64", lambda: A.ball_query(new_xyz, xyz, 0.05, ns),
        (lambda: rA.ball_query(new_xyz, xyz, 0.05, ns)) if rA else None,
        (lambda: rB.ball_query_wrapper(B, N, m, 0.05, ns, new_xyz, xyz, idxb)) if rB else None,
        bytes_per_scene=12 * N + 12 * m + 4 * m * ns)
    uni = torch.from_numpy(scenes.scene_batch(range(B), N, "uniform")).to(dev)
    uq = uni[:, :m].contiguous()
    add("ball_query uniform r=.05 (full scans)", lambda: A.ball_query(uq, uni, 0.05, ns),
        (lambda: rA.ball_query(uq, uni, 0.05, ns)) if rA else None,
        (lambda: rB.ball_query_wrapper(B, N, m, 0.05, ns, uq, uni, idxb)) if rB else None,
        bytes_per_scene=12 * N + 12 * m + 4 * m * ns)
    idx = A.ball_query(new_xyz, xyz, 0.05, ns)
    feats = torch.randn((B, 128, N), generator=g).to(dev)
    for name, f in (("C=3", xyz_t), ("C=128", feats)):
        C = f.shape[1]
        outg = torch.empty((B, C, m, ns), device=dev)
        nbytes = 4 * C * N + 4 * m * ns + 4 * C * m * ns
        add(f"group fwd {name}", lambda: A.group_points(f, idx), (lambda: rA.group_points(f, idx)) if rA else None,
            (lambda: rB.group_points_wrapper(B, C, N, m, ns, f, idx, outg)) if rB else None, bytes_per_scene=nbytes)
        gout = torch.randn((B, C, m, ns), generator=g).to(dev)
        gin = torch.zeros((B, C, N), device=dev)
        add(f"group bwd {name}", lambda: A.group_points_grad(gout, idx, N),
            (lambda: rA.group_points_grad(gout, idx, N)) if rA else None,
            (lambda: (gin.zero_(), rB.group_points_grad_wrapper(B, C, N, m, ns, gout, idx, gin))) if rB else None,
            bytes_per_scene=nbytes)
    # InvResMLP-sized group (the bulk of the backbone's bytes): C=128, N=m=2048, ns=64
    sub = xyz[:, :2048].contiguous()
    idx2 = A.ball_query(sub, sub, 0.08, 64)
    f2 = torch.randn((B, 128, 2048), generator=g).to(dev)
    out2 = torch.empty((B, 128, 2048, 64), device=dev)
    nbytes = 4 * 128 * 2048 + 4 * 2048 * 64 + 4 * 128 * 2048 * 64
    add("group fwd C=128 N=m=2048 ns=64", lambda: A.group_points(f2, idx2), (lambda: rA.group_points(f2, idx2)) if rA else None,
        (lambda: rB.group_points_wrapper(B, 128, 2048, 2048, 64, f2, idx2, out2)) if rB else None, bytes_per_scene=nbytes)
    g2 = torch.randn((B, 128, 2048, 64), generator=g).to(dev)
    gin2 = torch.zeros((B, 128, 2048), device=dev)
    add("group bwd C=128 N=m=2048 ns=64", lambda: A.group_points_grad(g2, idx2, 2048), None,
        (lambda: (gin2.zero_(), rB.group_points_grad_wrapper(B, 128, 2048, 2048, 64, g2, idx2, gin2))) if rB else None,
        bytes_per_scene=nbytes)
    add("ball_query N=m=2048 r=.08 ns=64 (B)", lambda: A.ball_query(sub, sub, 0.08, 64), None,
        (lambda: rB.ball_query_wrapper(B, 2048, 2048, 0.08, 64, sub, sub, torch.zeros_like(idx2))) if rB else None,
        bytes_per_scene=12 * 2048 * 2 + 4 * 2048 * 64)

    # ---- cfg3: FP chain ----
    d2, i3 = A.three_nn(xyz, new_xyz)
    d2b, i3b = torch.empty_like(d2), torch.empty_like(i3)
    add("three_nn 20000x1024", lambda: A.three_nn(xyz, new_xyz), (lambda: rA.three_nn(xyz, new_xyz)) if rA else None,
        (lambda: rB.three_nn_wrapper(B, N, m, xyz, new_xyz, d2b, i3b)) if rB else None, bytes_per_scene=12 * N + 12 * m + 24 * N)
    dist = torch.sqrt(d2)
    recip = 1.0 / (dist + 1e-8)
    w = (recip / recip.sum(dim=2, keepdim=True)).contiguous()
    fk = torch.randn((B, 256, m), generator=g).to(dev)
    oi = torch.empty((B, 256, N), device=dev)
    nbytes = 4 * 256 * m + 12 * N + 12 * N + 4 * 256 * N
    add("three_interpolate fwd C=256", lambda: A.three_interpolate(fk, i3, w), (lambda: rA.three_interpolate(fk, i3, w)) if rA else None,
        (lambda: rB.three_interpolate_wrapper(B, 256, m, N, fk, i3, w, oi)) if rB else None, bytes_per_scene=nbytes)
    go = torch.randn((B, 256, N), generator=g).to(dev)
    gi = torch.zeros((B, 256, m), device=dev)
    add("three_interpolate bwd C=256", lambda: A.three_interpolate_grad(go, i3, w, m),
        (lambda: rA.three_interpolate_grad(go, i3, w, m)) if rA else None,
        (lambda: (gi.zero_(), rB.three_interpolate_grad_wrapper(B, 256, N, m, go, i3, w, gi))) if rB else None, bytes_per_scene=nbytes)

    # ---- cfg4: grasp crop ----
    rng = np.random.default_rng(0)
    v = rng.normal(size=(B, m, 3)).astype(np.float32)
    rot = torch.from_numpy(scenes.viewpoint_rotations(-v, np.full((B, m), 0.3, np.float32)).reshape(B, m, 9)).to(dev)
    for hmax in (0.01, 0.04):
        add(f"cylinder_query r=.05 hmax={hmax}", lambda: A.cylinder_query(new_xyz, xyz, rot, 0.05, -0.02, hmax, ns),
            (lambda: rA.cylinder_query(new_xyz, xyz, rot, 0.05, -0.02, hmax, ns)) if rA else None,
            bytes_per_scene=12 * N + 48 * m + 4 * m * ns)
    ref3 = xyz_t
    q3 = new_xyz.transpose(1, 2).contiguous()
    for k in (1, 64):
        want = torch.empty((B, k, m), dtype=torch.int64, device=dev)
        rows_before = len(rows)
        add(f"knn R=20000 Q=1024 k={k}", lambda: knn_modules.knn_k(ref3, q3, k), None, None, bytes_per_scene=12 * (N + m) + 8 * k * m)
        if rC is not None and len(rows) > rows_before:
            t = timeit(lambda: rC.knn(ref3, q3, want))
            rows[rows_before]["ref_c_us"] = round(t, 2)
            rows[rows_before]["speedup_vs_c"] = round(t / rows[rows_before]["us"], 2)
            print(json.dumps(rows[rows_before]), flush=True)

    # ---- cfg1: collision ----
    raw = scenes.tabletop_scene(3, N).astype(np.float64)
    det = ModelFreeCollisionDetector(raw, voxel_size=0.01, device=dev)
    gs = scenes.grasp_set(4, det.scene_points, 1024)
    gg = scenes.GraspGroupStandIn(**gs)
    import oracle
    thr = oracle.collision_thresholds(gs["heights"], gs["depths"], gs["widths"], 0.03)
    Td, Rd, thd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (gs["translations"], gs["rotation_matrices"], thr))
    npts = det.scene_points.shape[0]
    add(f"collision counts G=1024 N'={npts} (kernel)", lambda: collision_counts(det._scene_dev, Td, Rd, thd), scenes_n=1,
        bytes_per_scene=24 * npts + 120 * 1024 + 1024)
    t0 = time.perf_counter()
    for _ in range(5):
        det.detect(gg)
    rows.append({"op": "collision detect() e2e host->host", "us": round((time.perf_counter() - t0) / 5 * 1e6, 1)})
    print(json.dumps(rows[-1]), flush=True)
    # BASELINE config 1 end to end: constructor (voxel down-sample of the raw 20k-point scene) + detect, host arrays in, mask out
    import oracle
    voxel_down_sample = oracle.voxel_down_sample
    for _ in range(2):
        ModelFreeCollisionDetector(raw, voxel_size=0.01, device=dev).detect(gg)
    t0 = time.perf_counter()
    for _ in range(5):
        ModelFreeCollisionDetector(raw, voxel_size=0.01, device=dev).detect(gg)
    t_gpu = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(3):
        voxel_down_sample(raw, 0.01)
    t_np = (time.perf_counter() - t0) / 3
    rows.append({"op": "collision __init__(20k points, voxel .01) + detect(1024 grasps) e2e", "us": round(t_gpu * 1e6, 1),
                 "host_numpy_voxel_down_sample_alone_us": round(t_np * 1e6, 1)})
    print(json.dumps(rows[-1]), flush=True)

    if args.sweep:
        for C in (1, 2, 4, 8, 16):
            for Tn in (256, 512, 1024):
                _lib.set_tuning("fps_cluster", C)
                _lib.set_tuning("fps_threads", Tn)
                try:
                    t = timeit(lambda: A.furthest_point_sampling(xyz, 1024), iters=5, warm=1)
                    rows.append({"op": f"sweep fps B={B} cluster={C} threads={Tn}", "us": round(t, 1), "us_per_iter": round(t / 1023, 4)})
                    print(json.dumps(rows[-1]), flush=True)
                except RuntimeError as e:
                    print("sweep failed", C, Tn, e)
        _lib.set_tuning("fps_cluster", 0)
        _lib.set_tuning("fps_threads", 0)
        for Bs in (1, 8, 16, 32, 64):
            xs = torch.from_numpy(scenes.scene_batch(range(Bs), N, "tabletop")).to(dev)
            t = timeit(lambda: A.furthest_point_sampling(xs, 1024), iters=5, warm=1)
            row = {"op": f"sweep fps B={Bs} auto", "us": round(t, 1), "scenes_per_s": round(Bs / (t * 1e-6), 1)}
            if rA is not None:
                row["ref_a_us"] = round(timeit(lambda: rA.furthest_point_sampling(xs, 1024), iters=3, warm=1), 1)
            rows.append(row)
            print(json.dumps(row), flush=True)
        for mode in (0, 1):
            _lib.set_tuning("group_mode", mode)
            t = timeit(lambda: A.group_points(feats, idx))
            rows.append({"op": f"sweep group fwd C=128 mode={mode}", "us": round(t, 1)})
            print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("group_mode", 0)
        for sp in (1, 2, 4):
            _lib.set_tuning("group_split", sp)
            t = timeit(lambda: A.group_points(feats, idx))
            rows.append({"op": f"sweep group fwd C=128 split={sp}", "us": round(t, 1)})
            print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("group_split", 0)
        for q in (1, 2, 4):
            _lib.set_tuning("query_qpw", q)
            t = timeit(lambda: A.ball_query(uq, uni, 0.05, ns))
            rows.append({"op": f"sweep ball uniform qpw={q}", "us": round(t, 1)})
            print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("query_qpw", 0)

    if args.sweep2:
        for pct in (100, 67, 50, 33, 25):
            _lib.set_tuning("grid_cell_pct", pct)
            for name, fn in (("ball r=.05", lambda: A.ball_query(new_xyz, xyz, 0.05, ns)),
                             ("cyl hmax=.04", lambda: A.cylinder_query(new_xyz, xyz, rot, 0.05, -0.02, 0.04, ns)),
                             ("cyl r=.02 hmax=.01", lambda: A.cylinder_query(new_xyz, xyz, rot, 0.02, -0.02, 0.01, ns))):
                t = timeit(fn, iters=5)
                rows.append({"op": f"sweep2 {name} B={B} grid_cell_pct={pct}", "us": round(t, 1)})
                print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("grid_cell_pct", 0)
        _lib.set_tuning("query_mode", 1)
        for name, fn in (("ball r=.05", lambda: A.ball_query(new_xyz, xyz, 0.05, ns)),
                         ("cyl hmax=.04", lambda: A.cylinder_query(new_xyz, xyz, rot, 0.05, -0.02, 0.04, ns))):
            t = timeit(fn, iters=5)
            rows.append({"op": f"sweep2 {name} B={B} full scan", "us": round(t, 1)})
            print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("query_mode", 0)
        for cc in (1, 2, 4, 8):
            _lib.set_tuning("scatter_cc", cc)
            for name, fn in (("group bwd C=128 N=20000", lambda: A.group_points_grad(gout, idx, N)),
                             ("group bwd C=128 N=2048", lambda: A.group_points_grad(g2, idx2, 2048)),
                             ("interp bwd C=256", lambda: A.three_interpolate_grad(go, i3, w, m))):
                t = timeit(fn, iters=5)
                rows.append({"op": f"sweep2 {name} B={B} scatter_cc={cc}", "us": round(t, 1)})
                print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("scatter_cc", 0)
        # the other InvResMLP grouping shapes of the backbone (drp.py:169-247): N = m, C = 256
        for (nn, cc_, nsb, rr) in ((1024, 256, 32, 0.2), (512, 256, 16, 0.4), (256, 256, 16, 0.6)):
            subx = xyz[:, :nn].contiguous()
            idxn = A.ball_query(subx, subx, rr, nsb)
            gn = torch.randn((B, cc_, nn, nsb), generator=g).to(dev)
            nb = 4 * cc_ * nn + 4 * nn * nsb + 4 * cc_ * nn * nsb
            for cc in (0, 2, 4, 8):
                _lib.set_tuning("scatter_cc", cc)
                t = timeit(lambda: A.group_points_grad(gn, idxn, nn), iters=5)
                rows.append({"op": f"sweep2 group bwd C={cc_} N=m={nn} ns={nsb} B={B} scatter_cc={cc}", "us": round(t, 1),
                             "hbm_frac": round(nb * B / (t * 1e-6) / 1e9 / HBM, 3)})
                print(json.dumps(rows[-1]), flush=True)
        _lib.set_tuning("scatter_cc", 0)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"gpu": torch.cuda.get_device_name(0), "B": B, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
