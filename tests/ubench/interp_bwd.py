#!/usr/bin/env python
"""three_interpolate backward (seg_sort_dense_kernel + seg_dense_kernel) at the shapes of a pipeline step, on real three_nn
indices: microseconds per call (CUDA events, L2 flushed between calls) and the fraction of the HBM peak.  Run with
GBOPS_LIB=<variant .so> to time another build in the same gpurun call."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, pointnet2_utils as pu, scenes  # noqa: E402

from graspbalance_b200 import _lib  # noqa: E402

for kv in filter(None, os.environ.get("GB_TUNE", "").split(",")):  # GB_TUNE=scatter_mode=4,scatter_cc=4
    k, v = kv.split("=")
    _lib.set_tuning(k, int(v))
HBM = 6542.4
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {}
for B in (32, 4):
    xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
    inds, lv1 = pu.furthest_point_sample_xyz(xyz, 1024)
    for (n, m, C) in ((20000, 1024, 256), (1024, 512, 256), (512, 256, 256)):
        unknown = xyz if n == 20000 else lv1[:, :n].contiguous()
        known = lv1[:, :m].contiguous()
        _, idx, w = pu.three_nn_weights(unknown, known)
        gout = torch.randn((B, C, n), device=dev)
        us = timeit(lambda: A.three_interpolate_grad(gout, idx, w, m))
        by = B * (4 * C * n + 24 * n + 4 * C * m)
        out[f"B{B}_n{n}_m{m}"] = {"us": round(us, 1), "hbm_frac": round(by / (us * 1e-6) / 1e9 / HBM, 3)}
print(json.dumps({"lib": os.environ.get("GBOPS_LIB", "default"), "tune": os.environ.get("GB_TUNE", ""), "interp_bwd": out}))
