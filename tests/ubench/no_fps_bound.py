#!/usr/bin/env python
"""How much of a step is the sampling chain really worth?  The same pipeline with the samples given (no FPS launched at all)
against the normal and the prefetched schedules."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from graspbalance_b200 import pipeline

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
host, offs = bench.make_host_inputs(list(range(B)))
pipe = pipeline.OpPipeline(B, bench.N_POINTS, dev, seed=0, backward=True, overlap=True)
inp = bench.to_device(host, offs, dev)
bufs = [pipe.alloc_samples(), pipe.alloc_samples()]
pipe.sampling_chain(inp[0], bufs[0])
pipe.sampling_chain(inp[0], bufs[1])

def timeit(fn, n=10, warm=4):
    for k in range(warm):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("normal      ms/step", round(timeit(lambda k: pipe.run(*inp)), 3), flush=True)
print("prefetched  ms/step", round(timeit(lambda k: pipe.run(*inp, samples=bufs[k % 2], prefetch=(inp[0], bufs[(k + 1) % 2]))), 3), flush=True)
print("no FPS      ms/step", round(timeit(lambda k: pipe.run(*inp, samples=bufs[0])), 3), flush=True)
