#!/usr/bin/env python
"""Warp-private group backward on index tensors that leave its fast path: random indices with repeats inside a row,
padded rows, out-of-range indices, kNN-ordered rows, ragged channel counts, strided sources -- against an fp64 scatter_add."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graspbalance_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
ok_all = True
_lib.set_tuning("scatter_mode", 16)  # private path for any number of tasks
for (B, C, n, m, ns, kind) in [(3, 7, 300, 64, 32, "random"), (2, 4, 2048, 128, 64, "random"), (2, 9, 1000, 50, 16, "random"),
                               (2, 17, 500, 40, 8, "random"), (2, 8, 1024, 96, 32, "padded"), (2, 8, 2048, 33, 64, "padded"),
                               (2, 5, 700, 77, 16, "padded"), (2, 6, 512, 64, 32, "sorted"), (2, 12, 256, 64, 64, "oob"),
                               (1, 4, 2400, 10, 128, "random"), (2, 130, 1024, 64, 32, "sorted")]:
    if kind == "random":
        idx = torch.randint(0, n, (B, m, ns), generator=g, dtype=torch.int32)
    elif kind == "sorted":
        idx = torch.stack([torch.stack([torch.randperm(n, generator=g)[:ns].sort().values for _ in range(m)]) for _ in range(B)]).int()
    elif kind == "padded":
        rows = []
        for _ in range(B * m):
            cnt = int(torch.randint(1, ns + 1, (1,), generator=g))
            r = torch.randperm(n, generator=g)[:cnt].sort().values
            rows.append(torch.cat([r, r[:1].expand(ns - cnt)]))
        idx = torch.stack(rows).reshape(B, m, ns).int()
    else:
        idx = torch.randint(-5, n + 5, (B, m, ns), generator=g, dtype=torch.int32)
    idx = idx.to(dev)
    gout = torch.randn((B, C + 3, m, ns), generator=g).to(dev)
    for strided in (False, True):
        src = gout[:, 3:] if strided else gout[:, 3:].contiguous()
        want = torch.zeros((B, C, n + 10), dtype=torch.float64, device=dev)
        safe = idx.long().clamp(-1, n).reshape(B, 1, m * ns)
        safe = torch.where((safe < 0) | (safe >= n), torch.full_like(safe, n + 5), safe)
        want.scatter_add_(2, safe.expand(-1, C, -1), gout[:, 3:].double().reshape(B, C, m * ns))
        want = want[:, :, :n]
        for overwrite in (1, 0):
            grad = torch.full((B, C, n), 0.5 if not overwrite else float("nan"), device=dev)
            stride = (C + 3) * m * ns if strided else C * m * ns
            ptr = gout.data_ptr() + 12 * m * ns if strided else src.data_ptr()
            l0 = _lib.launch_count()
            _lib.call("gb_group_bwd_strided", gout, ptr, idx.data_ptr(), grad.data_ptr(), B, C, n, m, ns, stride, overwrite)
            torch.cuda.synchronize()
            ref = want + (0.0 if overwrite else 0.5)
            err = (grad.double() - ref).abs().max().item()
            ok = err <= 1e-5 * max(ref.abs().max().item(), 1.0) and _lib.launch_count() - l0 == 1
            ok_all &= ok
            print(f"B={B} C={C} n={n} m={m} ns={ns} {kind} strided={strided} overwrite={overwrite}: err {err:.3g} "
                  f"launches {_lib.launch_count() - l0} {'ok' if ok else 'MISMATCH'}", flush=True)
_lib.set_tuning("scatter_mode", 0)
print("ALL_OK" if ok_all else "MISMATCH")
sys.exit(0 if ok_all else 1)
