#!/usr/bin/env python
"""Small torch-free workload that drives every kernel family of librc_b200.so once per scalar type, meant to
be run under compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_workload.py [f64 f32 c64 c32] [--big]

It checks nothing but finiteness and coarse reconstruction errors: the point is the sanitizer's report.
Sizes are chosen so that the tensor-pipe GEMMs (TMA + DMMA, tcgen05 kind::tf32), both TSQR modes, the
one-CTA and the cooperative pivoted QR, Cholesky-QR2, Jacobi, TRSM and the generators all launch."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rusty_compression_b200 import api  # noqa: E402

NAMES = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}


def low_rank(m, n, r, dtype, seed):
    rng = np.random.default_rng(seed)
    cplx = np.dtype(dtype).kind == "c"
    g = lambda *s: (rng.standard_normal(s) + 1j * rng.standard_normal(s)) if cplx else rng.standard_normal(s)
    sig = 10.0 ** (-np.arange(r) / 6.0)
    return ((g(m, r) * sig) @ g(r, n)).astype(dtype)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(dtype, big):
    ctx = api.default_context()
    single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
    tol = 2e-3 if single else 1e-6
    m, n, r = (4096, 1024, 48) if big else (1024, 384, 32)
    k, p = (32, 8) if big else (24, 8)
    a = low_rank(m, n, r, dtype, 1)
    op = api.DeviceMatrix.from_numpy(a)
    for qr_mode in (0, 1):                                  # Cholesky-QR2 fast path / Householder TSQR
        ctx.set_option("qr_mode", qr_mode)
        q = api.sample_range_by_rank(op, k, p, seed=3)
        q2 = api.sample_range_power_iteration(op, k, p, 1, seed=3)
        assert np.all(np.isfinite(q)) and np.all(np.isfinite(q2))
        qa, hist = api.sample_range_adaptive(op, 1e-2 if single else 1e-4, 16, seed=5, max_rank=256)
        assert np.all(np.isfinite(qa)) and len(hist) >= 1
        qr = api.QR.compute_from_range_estimate(q2, op)
        assert rel(qr.to_mat(), a) < 0.05
        cid = qr.compress(api.RANK(k)).column_id()
        tid = cid.two_sided_id()
        assert rel(cid.to_mat(), a) < 0.05 and rel(tid.to_mat(), a) < 0.05
        x = low_rank(n, 7, 7, dtype, 9)
        assert np.all(np.isfinite(cid.dot(x))) and np.all(np.isfinite(tid.dot(x)))
        svd = api.SVD.compute_from_range_estimate(q2, op)
        assert rel(svd.to_mat(), a) < 0.05
        assert np.all(np.isfinite(svd.to_qr().to_mat()))
    ctx.set_option("qr_mode", 0)
    # deterministic factorizations: tall (TSQR + one-CTA pivoting), short-wide and square (cooperative kernel)
    for shape in [(600, 50), (40, 700), (96, 96), (130, 131)]:
        b = low_rank(shape[0], shape[1], min(shape), dtype, 11)
        qr = api.QR.compute_from(b)
        assert rel(qr.to_mat(), b) < tol, (shape, rel(qr.to_mat(), b))
        lq = api.LQ.compute_from(b)
        assert rel(lq.to_mat(), b) < tol
        rid = lq.compress(api.RANK(min(20, min(shape)))).row_id()
        assert np.all(np.isfinite(rid.to_mat())) and np.all(np.isfinite(rid.two_sided_id().to_mat()))
        svd = api.SVD.compute_from(b)
        assert rel(svd.to_mat(), b) < tol * 10
    # permutations, norms, generators
    for mode in ("COL", "ROW", "COLINV", "ROWINV"):
        idx = np.random.default_rng(1).permutation(n if mode.startswith("COL") else m)
        assert np.all(np.isfinite(api.apply_permutation_matrix(a, idx, mode)))
    for mode in ("INV", "NOINV"):
        assert np.all(np.isfinite(api.apply_permutation_vector(a[0], np.random.default_rng(2).permutation(n), mode)))
    assert np.isfinite(api.max_col_norm(a)) and np.isfinite(api.rel_diff_fro(a, a + 1)) and np.isfinite(api.rel_diff_l2(a[0], a[1]))
    assert np.all(np.isfinite(api.random_orthogonal_matrix((64, 40), dtype, 4)))
    assert np.all(np.isfinite(api.random_approximate_low_rank_matrix((120, 60), 1.0, 1e-6, dtype, 4)))
    assert np.all(np.isfinite(api.decaying_spectrum_matrix((512, 256), dtype, 4, r0=64, decade_every=8.0).to_numpy()))
    assert np.all(np.isfinite(api.tall_shard_matrix(256, 512, 256, dtype, 4, 2048, r0=64).to_numpy()))
    if np.dtype(dtype).kind == "c":
        assert np.all(np.isfinite(api.helmholtz_kernel_matrix((256, 192), dtype).to_numpy()))
    ctx.synchronize()


if __name__ == "__main__":
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    big = "--big" in sys.argv
    for name in (args or list(NAMES)):
        t0 = time.time()
        run(NAMES[name], big)
        print(f"sanitize_workload: {name} ok ({time.time() - t0:.1f} s)", flush=True)
