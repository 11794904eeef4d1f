"""Stage timer of one configs[1] step (option "trace": the library synchronises and prints wall time at every mark, so
the stages are serialised -- use it to see where the non-GEMM time of a step goes, not for absolute numbers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api

ctx = api.default_context()
m, n, k, p, it = 65536, 8192, 64, 10, 2
a = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=512, decade_every=16.0)
for _ in range(3):
    q = api.sample_range_power_iteration(a, k, p, it, seed=42, device=True)
    svd = api.SVD.compute_from_range_estimate(q, a)
y = a.matmat(api.DeviceMatrix.random_gaussian((n, k + p), np.float64, 42))


def timed(label, fn, reps=20):
    fn(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ctx.synchronize()
    print(f"{label:<58s} {(time.perf_counter() - t0) / reps * 1e3:8.3f} ms", flush=True)
    return out


timed("pivoted QR of the 65536 x 74 sketch (QR::compute_from)", lambda: api.QR.compute_from(y))
z = timed("Z = A^H Q (8192 x 74)", lambda: a.conj_matmat(q))
timed("pivoted QR of the 8192 x 74 factor", lambda: api.QR.compute_from(z))
timed("SVD::compute_from_range_estimate (B pass + SVD of b + U)", lambda: api.SVD.compute_from_range_estimate(q, a))
timed("sample_range_power_iteration", lambda: api.sample_range_power_iteration(a, k, p, it, seed=42, device=True))
for key, val in (("overlap", 0), ("speculate", 0)):
    ctx.set_option(key, val)
    timed(f"sample_range_power_iteration with {key} = {val}", lambda: api.sample_range_power_iteration(a, k, p, it, seed=42, device=True))
ctx.set_option("overlap", 1); ctx.set_option("speculate", 1)
if "--trace" in sys.argv:
    ctx.set_option("trace", 1)
    api.QR.compute_from(y)
    api.SVD.compute_from_range_estimate(q, a)
