"""Runs a few tall pivoted QRs (65536 x 74 f64) -- the profiling target for tsqr/pivqr kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
y = api.DeviceMatrix.random_gaussian((65536, 74), np.float64, 1)
for _ in range(3):
    qr = api.QR.compute_from(y)
    ctx.synchronize()
print("ok", qr.rank())
