"""Config 3 (f32 32768^2 adaptive range finder + column ID), warm: per-stage CUDA-event-free wall timing
after a warm-up pass.  Usage: tools/prof_config3.py [n] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api

ctx = api.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a = api.decaying_spectrum_matrix((n, n), np.float32, 1235, r0=1024, decade_every=64.0)
ctx.synchronize()


def timed(label, fn):
    ctx.synchronize(); t0 = time.perf_counter()
    out = fn()
    ctx.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"   {label:<46s} {dt:9.2f} ms", flush=True)
    return out


for rep in range(reps):
    print(f"pass {rep}")
    q, hist = timed("sample_range_adaptive(1e-4, 64)", lambda: api.sample_range_adaptive(a, 1e-4, 64, seed=42, device=True))
    qr = timed("QR::compute_from_range_estimate", lambda: api.QR.compute_from_range_estimate(q, a))
    qrc = timed("compress(ADAPTIVE(1e-4))", lambda: qr.compress(api.ADAPTIVE(1e-4)))
    cid = timed("column_id()", lambda: qrc.column_id())
    print("   hist", hist[-1], "rank", qrc.rank())
    del q, qr, qrc, cid
