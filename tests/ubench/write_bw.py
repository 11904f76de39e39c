#!/usr/bin/env python
"""Pure-write and copy bandwidth of the GPU (the ceiling of the store-dominated group forward)."""
import json
import torch
dev = torch.device("cuda:0")
n = 2 * 1024 ** 3 // 4
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


tf = t(lambda: a.fill_(1.0))
tz = t(lambda: a.zero_())
tc = t(lambda: b.copy_(a))
print(json.dumps({"fill_GBs": round(a.numel() * 4 / tf / 1e9, 1), "memset_GBs": round(a.numel() * 4 / tz / 1e9, 1),
                  "copy_GBs_read_plus_write": round(2 * a.numel() * 4 / tc / 1e9, 1)}))
