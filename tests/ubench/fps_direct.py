#!/usr/bin/env python
"""FPS cluster exchange: every warp pushes its winner to all CTAs (direct) vs CTA-level stage + one push per CTA."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes

dev = torch.device("cuda:0")


def timeit(fn, iters=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


B = 32
xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
ok = True
for (n, m) in ((20000, 2048), (20000, 1024), (2048, 1024), (1024, 512), (5000, 700)):
    x = xyz[:, :n].contiguous()
    res = {}
    for mode in (1, 0):
        _lib.set_tuning("fps_direct", mode)
        res[mode] = (A.furthest_point_sampling(x, m), timeit(lambda: A.furthest_point_sampling(x, m)))
    same = bool(torch.equal(res[0][0], res[1][0]))
    ok &= same
    print(json.dumps({"n": n, "m": m, "staged_us": round(res[1][1], 1), "direct_us": round(res[0][1], 1), "identical": same,
                      "direct_us_per_round": round(res[0][1] / (m - 1), 4)}), flush=True)
_lib.set_tuning("fps_direct", 0)
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
