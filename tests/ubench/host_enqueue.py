#!/usr/bin/env python
"""How long does the host need to ENQUEUE one pipeline step (no synchronisation), next to the step's GPU time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from graspbalance_b200 import pipeline, _lib

dev = torch.device("cuda:0")
B = 32
host, offs = bench.make_host_inputs(list(range(B)))
for overlap in (True, False):
    pipe = pipeline.OpPipeline(B, bench.N_POINTS, dev, seed=0, backward=True, overlap=overlap)
    inp = bench.to_device(host, offs, dev)
    for _ in range(3):
        pipe.run(*inp)
    torch.cuda.synchronize()
    enq, tot = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        pipe.run(*inp)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    print(f"overlap={overlap}: host enqueue {min(enq):.2f} ms, step wall {min(tot):.2f} ms", flush=True)
