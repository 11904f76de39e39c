#!/usr/bin/env python
"""Group forward on the InvResMLP shapes: store policy and work-split knobs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes

HBM = 6542.4
dev = torch.device("cuda:0")
_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=9, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


B = 32
xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
fidx = A.furthest_point_sampling(xyz, 2048).long()
lv0 = torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()
g = torch.Generator(device="cpu").manual_seed(0)
for (label, n, m, ns, C, r) in (("irm0", 2048, 2048, 64, 128, 0.08), ("irm1", 1024, 1024, 32, 256, 0.2), ("irm2", 512, 512, 16, 256, 0.4),
                                ("irm3", 256, 256, 16, 256, 0.6), ("sa2", 2048, 1024, 32, 128, 0.1), ("sa3", 1024, 512, 16, 256, 0.2)):
    t, q = lv0[:, :n].contiguous(), lv0[:, :m].contiguous()
    idx = A.ball_query(q, t, r, ns)
    f = torch.randn((B, C, n), generator=g).to(dev)
    nbytes = B * (4 * C * n + 4 * m * ns + 4 * C * m * ns)
    row = {"op": f"group fwd {label}"}
    for name, ch, kb in (("auto", 0, 0), ("old_ch24_kb384", 24, 384), ("old_ch48_kb384", 48, 384), ("ch16_kb384", 16, 384), ("ch16_kb768", 16, 768)):
        _lib.set_tuning("group_ch", ch); _lib.set_tuning("group_target_kb", kb)
        us = timeit(lambda: A.group_points(f, idx))
        row[name] = (round(us, 1), round(nbytes / (us * 1e-6) / 1e9 / HBM, 3))
    _lib.set_tuning("group_ch", 0); _lib.set_tuning("group_target_kb", 0)
    print(json.dumps(row), flush=True)
