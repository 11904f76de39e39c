#!/usr/bin/env python
"""FPS launch shape for small shards (the strong-scaling limiter): cluster size x threads per CTA; identical picks required.
    python tests/ubench/fps_small_batch.py [--B 4]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes  # noqa: E402

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


ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4)
args = ap.parse_args()
xyz = torch.from_numpy(scenes.scene_batch(range(args.B), 20000, "tabletop")).to(dev)
ok = True
for (n, m) in ((20000, 2048), (2048, 1024), (1024, 512), (512, 256)):
    x = xyz[:, :n].contiguous()
    _lib.set_tuning("fps_cluster", 0), _lib.set_tuning("fps_threads", 0)
    want = A.furthest_point_sampling(x, m)
    row = {"B": args.B, "n": n, "m": m, "auto_us": round(timeit(lambda: A.furthest_point_sampling(x, m)), 1)}
    for c in (1, 2, 4, 8, 16):
        for t in (128, 256):
            _lib.set_tuning("fps_cluster", c), _lib.set_tuning("fps_threads", t)
            try:
                got = A.furthest_point_sampling(x, m)
            except Exception as e:  # shape not available
                row[f"C{c}xT{t}"] = None
                continue
            same = bool(torch.equal(got, want))
            ok &= same
            row[f"C{c}xT{t}"] = round(timeit(lambda: A.furthest_point_sampling(x, m)), 1) if same else "MISMATCH"
    print(json.dumps(row), flush=True)
_lib.set_tuning("fps_cluster", 0), _lib.set_tuning("fps_threads", 0)
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
