"""One call of each small one-CTA kernel at the config-2 sizes (74 x 74 f64), for `ncu --set full -k regex:...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
rng = np.random.default_rng(0)
r = np.triu(rng.standard_normal((74, 74))) * (10.0 ** (-np.arange(74) / 16.0))[:, None]
for _ in range(3):
    qr = api.QR.compute_from(r)                       # pivqr_fused_kernel
y = rng.standard_normal((4096, 74)) * (10.0 ** (-np.arange(74) / 30.0))[None, :]
for _ in range(3):
    api.QR.compute_from(y)                            # cholqr2: chol_inv_kernel x 2 + pivqr_fused_kernel
    api.SVD.compute_from(y[:, :64].copy())            # jacobi_kernel
print("ok", np.abs(np.diag(qr.r))[:3])
