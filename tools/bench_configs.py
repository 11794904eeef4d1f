"""BASELINE configs 3 and 5 at full size (parity-test cases, not bench lines): wall time of each stage
(second, warm pass reported as well as the first, cold one) plus the size-independent properties the
domain offers.  Usage: tools/bench_configs.py [3] [5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rusty_compression_b200 import api

ctx = api.default_context()
which = [int(a) for a in sys.argv[1:]] or [3, 5]


def timed(label, fn):
    ctx.synchronize(); t0 = time.perf_counter()
    out = fn()
    ctx.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"   {label:<46s} {dt:9.2f} ms", flush=True)
    return out


if 3 in which:
    # config 3: column ID via pivoted QR on the sketch, f32, 32768 x 32768, rank to tol 1e-4, s = 64
    n = 32768
    print(f"config 3: f32 {n}x{n}, sigma_j = 10^(-j/64), adaptive tol 1e-4, sample_size 64")
    a = timed("generate A on device", lambda: api.decaying_spectrum_matrix((n, n), np.float32, 1235, r0=1024, decade_every=64.0))
    for rep in ("cold", "warm"):
        print(f"  -- {rep} pass")
        t0 = time.perf_counter()
        q, hist = timed("sample_range_adaptive(1e-4, 64)", lambda: api.sample_range_adaptive(a, 1e-4, 64, seed=42, device=True))
        qr = timed("QR::compute_from_range_estimate", lambda: api.QR.compute_from_range_estimate(q, a))
        qrc = timed("compress(ADAPTIVE(1e-4))", lambda: qr.compress(api.ADAPTIVE(1e-4)))
        cid = timed("column_id()", lambda: qrc.column_id())
        ctx.synchronize(); print(f"   {'pipeline total':<46s} {(time.perf_counter() - t0) * 1e3:9.2f} ms")
    print("   ranks/residuals:", [(r, float(f"{e:.2e}")) for r, e in hist])
    k = qrc.rank()
    # property: ||A - C Z|| / ||A|| ~ tol, checked on a random probe (A - CZ) x without forming CZ
    x = api.DeviceMatrix.random_gaussian((n, 8), np.float32, 3)
    ax = a.matmat(x).to_numpy().astype(np.float64)
    czx = cid.dot(x.to_numpy()).astype(np.float64)
    print(f"   rank {k}; probe ||(A - CZ)x|| / ||Ax|| = {np.linalg.norm(ax - czx) / np.linalg.norm(ax):.3e}")
    ind = cid.col_ind
    assert sorted(ind.tolist()) == list(range(n))
    del a, q, qr, qrc, cid

if 5 in which:
    # config 5: two-sided ID, c64, 16384 x 16384 Helmholtz kernel matrix, rank 128
    n, k, p = 16384, 128, 10
    print(f"config 5: c64 {n}x{n} Helmholtz kernel, rank {k} (+{p})")
    a = timed("generate A on device", lambda: api.helmholtz_kernel_matrix((n, n), np.complex128))
    for rep in ("cold", "warm"):
        print(f"  -- {rep} pass")
        t0 = time.perf_counter()
        q = timed("sample_range_by_rank(128, 10)", lambda: api.sample_range_by_rank(a, k, p, seed=42, device=True))
        qr = timed("QR::compute_from_range_estimate", lambda: api.QR.compute_from_range_estimate(q, a))
        qrc = timed("compress(RANK(128))", lambda: qr.compress(api.RANK(k)))
        cid = timed("column_id()", lambda: qrc.column_id())
        ts = timed("two_sided_id()", lambda: cid.two_sided_id())
        ctx.synchronize(); print(f"   {'pipeline total':<46s} {(time.perf_counter() - t0) * 1e3:9.2f} ms")
    x = api.DeviceMatrix.random_gaussian((n, 4), np.complex128, 3)
    ax = a.matmat(x).to_numpy()
    err_c = np.linalg.norm(ax - cid.dot(x.to_numpy())) / np.linalg.norm(ax)
    err_t = np.linalg.norm(ax - ts.dot(x.to_numpy())) / np.linalg.norm(ax)
    print(f"   probe errors: column ID {err_c:.3e}, two-sided ID {err_t:.3e}")
    ri, ci = ts.row_ind[:k], ts.col_ind[:k]
    sk = a.to_numpy()[np.ix_(ri, ci)]
    print(f"   skeleton check ||X - A[row_ind, col_ind]|| / ||.|| = {np.linalg.norm(ts.x - sk) / np.linalg.norm(sk):.3e}")
