#!/usr/bin/env python
"""The group backward on the backbone's shapes with the launcher's own choices (one line per shape):
    [GBOPS_LIB=.variants/libgbops_<name>.so] python tests/ubench/bwd_defaults.py [--B 32] [--tune KEY=VALUE ...]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes  # noqa: E402
from bwd_shapes import timeit, HBM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--tune", action="append", default=[])
ap.add_argument("--shapes", default="irm0,irm1,sa2,irm2,sa3,irm3,sa4")
args = ap.parse_args()
for kv in args.tune:
    k, v = kv.split("=")
    _lib.set_tuning(k, int(v))
dev = torch.device("cuda:0")
B = args.B
g = torch.Generator(device="cpu").manual_seed(0)
xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
fidx = A.furthest_point_sampling(xyz, 2048).long()
lv0 = torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()
shapes = {"irm0": (2048, 2048, 64, 128, 0.08, 3), "irm1": (1024, 1024, 32, 256, 0.2, 6), "irm2": (512, 512, 16, 256, 0.4, 3),
          "sa2": (2048, 1024, 32, 128, 0.1, 1), "sa3": (1024, 512, 16, 256, 0.2, 1), "irm3": (256, 256, 16, 256, 0.6, 3),
          "sa4": (512, 256, 16, 256, 0.3, 1)}
out, total = {}, 0.0
for label in args.shapes.split(","):
    n, m, ns, C, r, per_step = shapes[label]
    idx = A.ball_query(lv0[:, :m].contiguous(), lv0[:, :n].contiguous(), r, ns)
    gout = torch.randn((B, C, m, ns), generator=g).to(dev)
    nbytes = B * (4 * C * n + 4 * m * ns + 4 * C * m * ns)
    t = timeit(lambda: A.group_points_grad(gout, idx, n), iters=9)
    out[label] = (round(t, 1), round(nbytes / (t * 1e-6) / 1e9 / HBM, 3))
    total += per_step * t
out["step_family_ms"] = round(total / 1000, 3)
print(os.environ.get("GBOPS_LIB", "default"), json.dumps(out))
