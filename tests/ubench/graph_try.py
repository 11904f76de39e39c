#!/usr/bin/env python
"""Capture one pipeline step (all streams, forward + backward) in a CUDA graph and compare replay time with eager."""
import os, sys, time
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
for _ in range(3):
    ref = pipe.run(*inp)
torch.cuda.synchronize()
ref = {k: v.clone() for k, v in ref.items()}

def timeit(fn, n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("eager ms/step", timeit(lambda: pipe.run(*inp)), flush=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        pipe.run(*inp)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g, stream=s):
    out = pipe.run(*inp)
torch.cuda.synchronize()
print("captured", flush=True)
g.replay()
torch.cuda.synchronize()
for k in ref:
    a, b = ref[k].double().cpu(), out[k].double().cpu()
    print(k, "max abs diff", float((a - b).abs().max()), flush=True)
print("graph ms/step", timeit(g.replay), flush=True)
