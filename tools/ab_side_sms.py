"""A/B of option "side_sms" on the adaptive sampler of config 3 (f32 32768^2, tol 1e-4, s = 64): the next sketch
Y' = A Omega' on all but `side_sms` SMs beside the projection + pivoted QR of the current sketch.  0 = off.
Usage: tools/ab_side_sms.py [values ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api

ctx = api.default_context()
args = sys.argv[1:]
dtype = np.float32
if args and args[0] in ("f32", "f64", "c64"):
    dtype = {"f32": np.float32, "f64": np.float64, "c64": np.complex128}[args.pop(0)]
vals = [int(v) for v in args] or [0, 4, 8, 16]
n = 32768 if dtype == np.float32 else 16384
a = (api.helmholtz_kernel_matrix((n, n), dtype) if dtype == np.complex128 else
     api.decaying_spectrum_matrix((n, n), dtype, 1235, r0=1024, decade_every=64.0))
print(f"{np.dtype(dtype).name} {n} x {n}")
ref_hist = None
for v in vals + vals[:1]:
    ctx.set_option("side_sms", v)
    best = 1e30
    for rep in range(4):
        ctx.synchronize(); t0 = time.perf_counter()
        q, hist = api.sample_range_adaptive(a, 1e-4, 64, seed=42, device=True)
        ctx.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
        del q
    if ref_hist is None:
        ref_hist = hist
    same = [r for r, _ in hist] == [r for r, _ in ref_hist]
    dev = max(abs(e - e0) / e0 for (_, e), (_, e0) in zip(hist, ref_hist))
    print(f"side_sms {v:3d}: sample_range_adaptive {best:8.3f} ms  (rank history {'identical' if same else 'DIFFERS'}, "
          f"residual history within {dev:.1e}; final rank {hist[-1][0]})", flush=True)
