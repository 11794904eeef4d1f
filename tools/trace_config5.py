"""Stage trace (option "trace": synchronising, so the sum is larger than the untraced pass) of config 5:
c64 16384^2 Helmholtz kernel matrix, by-rank sampling -> QR from range -> compress -> column ID -> two-sided ID."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
n, k, p = 16384, 128, 10
a = api.helmholtz_kernel_matrix((n, n), np.complex128)
def run():
    q = api.sample_range_by_rank(a, k, p, seed=42, device=True)
    qr = api.QR.compute_from_range_estimate(q, a)
    cid = qr.compress(api.RANK(k)).column_id()
    return cid.two_sided_id()
for _ in range(2):
    run()
ctx.synchronize(); t0 = time.perf_counter(); run(); ctx.synchronize()
print(f"untraced pass: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
ctx.set_option("trace", 1)
run()
ctx.set_option("trace", 0)
