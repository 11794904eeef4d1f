"""A/B of context options on the configs[1] step (sampler + SVD from range), interleaved so that clock / power drift
hits every variant alike.  CUDA events on the context stream, like bench.py.
    python tools/ab_step.py overlap=1,speculate=1 overlap=0,speculate=1 overlap=0,speculate=0"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rusty_compression_b200 import api

variants = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in v.split(",")) for v in sys.argv[1:]] or [{}]
ctx = api.Context(device=0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
m, n, k, p, it = 65536, 8192, 64, 10, 2
a = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=512, decade_every=16.0, ctx=ctx)


def step():
    q = api.sample_range_power_iteration(a, k, p, it, seed=42, ctx=ctx, device=True)
    return api.SVD.compute_from_range_estimate(q, a)


for _ in range(5):
    step()
res = {i: [] for i in range(len(variants))}
for rep in range(4):
    for i, v in enumerate(variants):
        for key, val in v.items():
            ctx.set_option(key, val)
        step(); step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        res[i].append(e0.elapsed_time(e1) / 10)
for i, v in enumerate(variants):
    print(v, " ".join(f"{t:7.3f}" for t in res[i]), "ms/step")
