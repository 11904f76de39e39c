#!/usr/bin/env python
"""Warp-private group backward: load width (priv_vl), L2 prefetch distance (priv_pf_kb) and warps per block (scatter_cc)
on the backbone's shapes.    python tests/ubench/bwd_private_sweep.py [--B 32] [--out gpurun_out/bwd_private_sweep.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _ext as A, _lib, scenes  # noqa: E402
from bwd_shapes import timeit, HBM  # noqa: E402

dev = torch.device("cuda:0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bwd_private_sweep.json"))
    ap.add_argument("--shapes", default="irm0,sa2,irm1,irm2")
    args = ap.parse_args()
    B = args.B
    g = torch.Generator(device="cpu").manual_seed(0)
    xyz = torch.from_numpy(scenes.scene_batch(range(B), 20000, "tabletop")).to(dev)
    fidx = A.furthest_point_sampling(xyz, 2048).long()
    lv0 = torch.gather(xyz, 1, fidx[:, :, None].expand(-1, -1, 3)).contiguous()
    shapes = {"irm0": (2048, 2048, 64, 128, 0.08), "irm1": (1024, 1024, 32, 256, 0.2), "irm2": (512, 512, 16, 256, 0.4),
              "sa2": (2048, 1024, 32, 128, 0.1), "sa3": (1024, 512, 16, 256, 0.2), "irm3": (256, 256, 16, 256, 0.6),
              "sa4": (512, 256, 16, 256, 0.3)}
    rows = []
    a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    b = torch.empty_like(a)
    t = timeit(lambda: b.copy_(a), iters=5)
    rows.append({"copy_gbs": round(2 * a.numel() * 4 / (t * 1e-6) / 1e9, 1)})
    print(json.dumps(rows[-1]), flush=True)
    del a, b
    for label in args.shapes.split(","):
        n, m, ns, C, r = shapes[label]
        idx = A.ball_query(lv0[:, :m].contiguous(), lv0[:, :n].contiguous(), r, ns)
        gout = torch.randn((B, C, m, ns), generator=g).to(dev)
        nbytes = B * (4 * C * n + 4 * m * ns + 4 * C * m * ns)
        _lib.set_tuning("scatter_mode", 8)
        ref = A.group_points_grad(gout, idx, n)
        t = timeit(lambda: A.group_points_grad(gout, idx, n), iters=7)
        row = {"shape": label, "path": "sorted (control)", "us": round(t, 1), "hbm_frac": round(nbytes / (t * 1e-6) / 1e9 / HBM, 3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        _lib.set_tuning("scatter_mode", 0)
        vls = (1, 2, 4) if ns >= 32 else (1,)
        for vl in vls:
            for cw in (4, 2):
                for w, sp in ((0, 0),):
                    _lib.set_tuning("priv_split", sp)
                    _lib.set_tuning("priv_vl", vl)
                    _lib.set_tuning("priv_cw", cw)
                    _lib.set_tuning("scatter_cc", w)
                    got = A.group_points_grad(gout, idx, n)
                    ok = (got - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1.0)
                    t = timeit(lambda: A.group_points_grad(gout, idx, n), iters=7)
                    row = {"shape": label, "vl": vl, "cw": cw, "W": w, "split": sp, "us": round(t, 1), "hbm_frac": round(nbytes / (t * 1e-6) / 1e9 / HBM, 3),
                           "ok": bool(ok)}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
    for k in ("priv_vl", "priv_cw", "scatter_cc", "priv_split"):
        _lib.set_tuning(k, 0)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"B": B, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
