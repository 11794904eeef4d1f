import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
which = sys.argv[1] if len(sys.argv) > 1 else "3"
if which == "3":
    n = 32768
    a = api.decaying_spectrum_matrix((n, n), np.float32, 1235, r0=1024, decade_every=64.0)
    run = lambda: api.QR.compute_from_range_estimate(api.sample_range_adaptive(a, 1e-4, 64, seed=42, device=True)[0], a).compress(api.ADAPTIVE(1e-4)).column_id()
else:
    n = 16384
    a = api.helmholtz_kernel_matrix((n, n), np.complex128)
    run = lambda: api.QR.compute_from_range_estimate(api.sample_range_by_rank(a, 128, 10, seed=42, device=True), a).compress(api.RANK(128)).column_id().two_sided_id()
for _ in range(2):
    run()
ctx.set_option("trace", 1)
run()
