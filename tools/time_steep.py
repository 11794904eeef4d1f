"""configs[1] pipeline on a STEEP spectrum (sigma_j = 10^(-j/6): the 74-column sketch spans twelve decades), the
kind of operator the crate targets.  The plain Cholesky-QR2 breaks down on Y = A Omega; times the step on the
shifted Cholesky-QR3 re-run against the Householder-TSQR re-run, next to the well-conditioned configs[1] operator.
Usage: tools/time_steep.py [rows] [passes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api

ctx = api.default_context()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, k, p, it = 8192, 64, 10, 2


def run(a, label):
    best = 1e30
    for i in range(passes):
        ctx.reset_counters()
        ctx.synchronize()
        t0 = time.perf_counter()
        q = api.sample_range_power_iteration(a, k, p, it, seed=42, device=True)
        svd = api.SVD.compute_from_range_estimate(q, a)
        ctx.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    s = svd.s_f64()
    print(f"{label:58s} {best:8.2f} ms per step   cholqr used {ctx.counter('cholqr_used')}, shifted {ctx.counter('cholqr_shifted')}, "
          f"rejected {ctx.counter('cholqr_fallbacks')};  s[0] {s[0]:.3e}  s[-1] {s[-1]:.3e}", flush=True)
    return s


a = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=512, decade_every=16.0, ctx=ctx)
run(a, "configs[1] operator (one decade per 16 singular values)")
a.free()
a = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=128, decade_every=6.0, ctx=ctx)
s1 = run(a, "steep operator, shifted Cholesky-QR3 re-run")
ctx.set_option("shifted_cholqr", 0)
s0 = run(a, "steep operator, Householder-TSQR re-run")
ctx.set_option("shifted_cholqr", 1)
print(f"singular values of the two routes: max |diff| / s[0] = {np.max(np.abs(s1 - s0)) / s0[0]:.2e}")
ctx.set_option("qr_mode", 1)
run(a, "steep operator, Householder TSQR from the start (qr_mode 1)")
ctx.set_option("qr_mode", 0)
