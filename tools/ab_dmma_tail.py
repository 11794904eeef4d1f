#!/usr/bin/env python
"""A/B of the DFMA tail path of the DMMA GEMM (context option "dmma_tail"): correctness against numpy on ragged
shapes (l % 8 in {2, 4}; NN, TN and c64), then wall-clock timing around synchronised launches at the config-2 shape."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rusty_compression_b200 import api  # noqa: E402

ctx = api.default_context()
rng = np.random.default_rng(0)
for (m, n, l) in [(4096, 2048, 74), (4096, 2048, 76), (1000, 520, 10), (1000, 520, 12), (4096, 2048, 26), (777, 512, 90)]:
    a = rng.standard_normal((m, n)); x = rng.standard_normal((n, l)); y = rng.standard_normal((m, l))
    ad = api.DeviceMatrix.from_numpy(a)
    az = a[:, : n // 2] + 1j * a[:, n // 2:]; xz = x[: n // 2] + 1j * x[n // 2:]
    azd = api.DeviceMatrix.from_numpy(az)
    for tail in (0, 1):
        ctx.set_option("dmma_tail", tail)
        e_nn = np.max(np.abs(ad.matmat(x).to_numpy() - a @ x)) / np.max(np.abs(a @ x))
        e_tn = np.max(np.abs(ad.conj_matmat(y).to_numpy() - a.T @ y)) / np.max(np.abs(a.T @ y))
        e_z = np.max(np.abs(azd.matmat(xz).to_numpy() - az @ xz)) / np.max(np.abs(az @ xz))
        print(f"{m}x{n}x{l} dmma_tail={tail}: NN err {e_nn:.2e}  TN err {e_tn:.2e}  c64 NN err {e_z:.2e}", flush=True)
        assert e_nn < 1e-13 and e_tn < 1e-13 and e_z < 1e-13

m, n, l = 65536, 8192, 74
ad = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=512, decade_every=16.0)
om = api.DeviceMatrix.random_gaussian((n, l), np.float64, 42)
yd = ad.matmat(om)
for rep in range(2):
    for tail in (0, 1):
        ctx.set_option("dmma_tail", tail)
        for name, fn in (("NN", lambda: ad.matmat(om)), ("TN", lambda: ad.conj_matmat(yd))):
            for _ in range(3):
                fn()
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            ctx.synchronize()
            ms = (time.perf_counter() - t0) / 20 * 1e3
            print(f"config-2 {name} dmma_tail={tail}: {ms:.3f} ms = {2.0 * m * n * l / ms / 1e9:.2f} TFLOP/s", flush=True)
