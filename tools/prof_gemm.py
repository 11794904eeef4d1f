"""Runs the two dominant contractions of config 2 (Y = A Omega, Z = A^H Y; 65536 x 8192 f64, l = 74):
the profiling target for the TMA + DMMA GEMM kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
ctx = api.default_context()
a = api.DeviceMatrix.random_gaussian((65536, 8192), np.float64, 1)
om = api.DeviceMatrix.random_gaussian((8192, 74), np.float64, 2)
for _ in range(3):
    y = a.matmat(om)
    z = a.conj_matmat(y)
ctx.synchronize()
print("ok", y.shape, z.shape)
