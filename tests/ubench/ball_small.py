#!/usr/bin/env python
"""Ball query on the InvResMLP shapes (queries = support points): full scan vs the cell grid (query_mode = 2 forces it)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes

dev = torch.device("cuda:0")


def timeit(fn, iters=9, warm=2):
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
fidx = A.furthest_point_sampling(xyz, 2048).long()
lv0 = torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()
ok = True
for (n, m, r, ns) in ((2048, 2048, 0.08, 64), (1024, 1024, 0.2, 32), (512, 512, 0.4, 16), (256, 256, 0.6, 16), (2048, 1024, 0.1, 32),
                      (1024, 512, 0.2, 16), (512, 256, 0.3, 16)):
    t, q = lv0[:, :n].contiguous(), lv0[:, :m].contiguous()
    res = {}
    for mode in (0, 2):
        _lib.set_tuning("query_mode", mode)
        res[mode] = (A.ball_query(q, t, r, ns), timeit(lambda: A.ball_query(q, t, r, ns)))
    same = bool(torch.equal(res[0][0], res[2][0]))
    ok &= same
    print(json.dumps({"n": n, "m": m, "r": r, "ns": ns, "scan_us": round(res[0][1], 1), "grid_us": round(res[2][1], 1), "identical": same}), flush=True)
_lib.set_tuning("query_mode", 0)
print("OK" if ok else "MISMATCH")
