#!/usr/bin/env python
"""One backbone backward shape, a few launches (for ncu):  python tests/ubench/bwd_one.py --shape irm0 [--vl 2 --pf -1 --mode 0]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="irm0")
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--vl", type=int, default=0)
ap.add_argument("--cw", type=int, default=0)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
shapes = {"irm0": (2048, 2048, 64, 128, 0.08), "irm1": (1024, 1024, 32, 256, 0.2), "irm2": (512, 512, 16, 256, 0.4),
          "sa2": (2048, 1024, 32, 128, 0.1), "sa3": (1024, 512, 16, 256, 0.2)}
n, m, ns, C, r = shapes[args.shape]
B = args.B
g = torch.Generator(device="cpu").manual_seed(0)
xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
fidx = A.furthest_point_sampling(xyz, 2048).long()
lv0 = torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()
idx = A.ball_query(lv0[:, :m].contiguous(), lv0[:, :n].contiguous(), r, ns)
gout = torch.randn((B, C, m, ns), generator=g).to(dev)
_lib.set_tuning("priv_vl", args.vl)
_lib.set_tuning("priv_cw", args.cw)
_lib.set_tuning("scatter_mode", args.mode)
for _ in range(args.iters):
    out = A.group_points_grad(gout, idx, n)
torch.cuda.synchronize()
print("done", float(out.abs().sum()))
