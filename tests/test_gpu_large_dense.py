"""the non-randomized factorisations of a LARGE dense matrix -- QR::compute_from (src/qr.rs:251-253,
examples/interpolative_decomposition.rs:25), LQ::compute_from (src/qr.rs:354-362) and SVD::compute_from
(src/svd.rs:165-169) -- against ?geqp3 / ?orgqr / ?gesdd on the same matrix.  These sizes leave the one-CTA kernels:
panel QR (block Gram-Schmidt + Cholesky-QR2 / TSQR panels), the cooperative pivoted Householder kernel on the
triangle, Q formed by compact-WY blocks (GEMMs), and the cooperative whole-GPU Jacobi kernel."""
import numpy as np
import pytest

from oracle import reference_path as ref
from golden_common import adjudicate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as _api
    return _api


def relerr(x, y):
    return float(np.linalg.norm(x - y) / np.linalg.norm(y))


def dense_matrix(shape, dtype, seed, decades=6.0):
    """Full-rank dense matrix with a geometric spectrum over `decades` decades and Gaussian singular vectors (cheap to
    build at size: G1 diag(s) G2 with orthonormalised thin factors)."""
    rng = np.random.default_rng(seed)
    m, n = shape
    k = min(m, n)
    real = np.float64
    def gauss(r, c):
        g = rng.standard_normal((r, c))
        if np.dtype(dtype).kind == "c":
            g = g + 1j * rng.standard_normal((r, c))
        return g
    u, _ = np.linalg.qr(gauss(m, k))
    v, _ = np.linalg.qr(gauss(n, k))
    s = 10.0 ** (-decades * np.arange(k, dtype=real) / k)
    return ((u * s) @ np.conj(v.T)).astype(dtype)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (1536, 1024)), (np.float64, (640, 1792)), (np.complex128, (1100, 768)),
                                         (np.float32, (1280, 1024))],
                         ids=["f64-tall", "f64-wide", "c64-tall", "f32-tall"])
def test_large_dense_pivoted_qr_matches_geqp3(api, dtype, shape):
    single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
    a = dense_matrix(shape, dtype, seed=11, decades=3.0 if single else 6.0)
    q, r, ind = api.pivoted_qr(a)
    q0, r0, ind0 = ref.pivoted_qr(a)
    k = min(shape)
    assert sorted(ind.tolist()) == list(range(shape[1]))
    assert q.shape == (shape[0], k) and r.shape == (k, shape[1])
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < (2e-5 if single else 1e-12)
    assert relerr(q.dot(r), a[:, ind]) < (1e-5 if single else 1e-13)
    assert np.max(np.abs(np.tril(r[:, :k], -1))) == 0.0
    # pivots: ?geqp3's sequence wherever the gap exceeds 1e-6 (validated step by step in double on a mismatch)
    order = adjudicate(a, ind, ind0, label=f"large dense {shape} {np.dtype(dtype).name}")
    if order is not None:
        q0, r0, ind0 = ref.pivoted_qr_with_order(a, order)
    d, d0 = np.abs(np.diagonal(r)), np.abs(np.diagonal(r0))
    assert np.max(np.abs(d - d0) / d0[0]) < (2e-5 if single else 1e-12)
    if not single:
        assert np.max(np.abs(d - d0) / d0) < 1e-8          # every diagonal entry to high RELATIVE accuracy


def test_large_dense_lq_and_id(api):
    """LQ::compute_from and the deterministic column ID of examples/interpolative_decomposition.rs at size."""
    a = dense_matrix((900, 1300), np.float64, seed=12, decades=8.0)
    lq = api.LQ.compute_from(a)
    lq0 = ref.LQ.compute_from(a)
    order = adjudicate(np.conj(a.T), lq.ind, lq0.ind, label="large dense LQ")
    assert order is None or len(order) == 900
    assert relerr(lq.to_mat(), a) < 1e-12
    qr = api.QR.compute_from(a).compress(api.ADAPTIVE(1e-4))
    qr0 = ref.QR.compute_from(a).compress(ref.ADAPTIVE(1e-4))
    assert qr.rank() == qr0.rank()
    cid, cid0 = qr.column_id(), qr0.column_id()
    e, e0 = ref.rel_diff_fro(cid.to_mat(), a), ref.rel_diff_fro(cid0.to_mat(), a)
    assert abs(e - e0) <= 1e-8 * e0, (e, e0)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (1536, 1024)), (np.float64, (700, 1800)), (np.complex128, (900, 600)),
                                         (np.float32, (1024, 640))],
                         ids=["f64-tall", "f64-wide", "c64-tall", "f32-tall"])
def test_large_dense_svd_matches_gesdd(api, dtype, shape):
    single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
    a = dense_matrix(shape, dtype, seed=13, decades=3.0 if single else 8.0)
    u, s, vt = api.compute_svd(a)
    u0, s0, vt0 = ref.compute_svd(a)
    k = min(shape)
    assert u.shape == u0.shape and vt.shape == vt0.shape and s.shape == s0.shape
    assert np.all(np.diff(s) <= 0)
    assert np.max(np.abs(s - s0)) / s0[0] < (2e-5 if single else 1e-12)          # (sgesdd itself is only good to ~n eps)
    if not single:
        assert np.max(np.abs(s - s0) / s0) < 1e-8
    assert relerr((u * s).dot(vt), a) < (2e-5 if single else 1e-12)
    assert np.max(np.abs(np.conj(u.T).dot(u) - np.eye(k))) < (5e-4 if single else 1e-10)
    assert np.max(np.abs(vt.dot(np.conj(vt.T)) - np.eye(k))) < (5e-4 if single else 1e-10)
    # the SVD container on top: to_qr (pivoted QR of diag(s) vt, src/svd.rs:150-163) at size
    svd = api.SVD.compute_from(a)
    assert relerr(svd.to_qr().to_mat(), a) < (5e-5 if single else 1e-11)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32], ids=["f64", "c64", "f32"])
def test_rank_deficient_inputs_keep_q_orthonormal(api, dtype):
    """?geqp3 / ?orgqr and ?gesdd return orthonormal factors whatever the rank.  The panel route (block Gram-Schmidt
    between column panels) has to work for that: an exactly rank-deficient dense matrix, the zero matrix, duplicated
    columns, and a sketch with more columns than the operator has rank leave whole panels without a direction of
    their own (Householder fallback with re-orthogonalisation and Gaussian completion, panel_qr in host_api.cu)."""
    rng = np.random.default_rng(31)
    single = np.dtype(dtype) == np.dtype(np.float32)
    cplx = np.dtype(dtype).kind == "c"
    tol_o, tol_r = (5e-5, 1e-5) if single else (1e-11, 1e-12)

    def gauss(r, c):
        g = rng.standard_normal((r, c))
        return (g + 1j * rng.standard_normal((r, c)) if cplx else g).astype(dtype)

    def check_qr(a, label):
        q, r, ind = api.pivoted_qr(a)
        k = min(a.shape)
        assert sorted(ind.tolist()) == list(range(a.shape[1])), label
        assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < tol_o, label
        assert np.linalg.norm(q.dot(r) - a[:, ind]) <= tol_r * max(np.linalg.norm(a), 1e-30) + 1e-30, label
        d = np.abs(np.diagonal(r))
        assert np.all(d[:-1] >= d[1:] * (1 - 1e-3) - (1e-4 if single else 1e-10) * max(d[0], 1e-30)), label

    check_qr(gauss(1500, 50) @ gauss(50, 900), "exact rank 50, 1500 x 900")
    check_qr(gauss(700, 40) @ gauss(40, 1600), "exact rank 40, 700 x 1600")
    check_qr(np.zeros((1000, 700), dtype), "zero 1000 x 700")
    dup = gauss(1200, 300)
    dup[:, 150:] = dup[:, :150]
    check_qr(dup, "duplicated columns")
    # SVD of an exactly rank-deficient matrix: singular values and the leading singular vectors
    a = gauss(900, 30) @ gauss(30, 600)
    u, s, vt = api.compute_svd(a)
    s0 = np.linalg.svd(a.astype(np.complex128 if cplx else np.float64), compute_uv=False)
    # (f32: the 570 singular values behind the rank sit at the roundoff level of the working precision, eps sqrt(m) s_0)
    assert np.max(np.abs(s - s0)) / s0[0] < (1e-4 if single else 1e-12)
    assert np.max(np.abs(s[:30] - s0[:30]) / s0[:30]) < (2e-5 if single else 1e-12)
    assert np.linalg.norm((u * s).dot(vt) - a) / np.linalg.norm(a) < (2e-5 if single else 1e-12)
    assert np.max(np.abs(np.conj(u[:, :30].T).dot(u[:, :30]) - np.eye(30))) < (2e-4 if single else 1e-10)
    assert np.max(np.abs(vt[:30].dot(np.conj(vt[:30].T)) - np.eye(30))) < (2e-4 if single else 1e-10)
    # a sketch with more columns than the operator has rank: the range basis must still be orthonormal
    op = gauss(4096, 100) @ gauss(100, 1024)
    q = api.sample_range_by_rank(op, 290, 10, seed=3)
    assert q.shape == (4096, 290)
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(290))) < tol_o
    proj = q[:, :128].dot(np.conj(q[:, :128].T).dot(op))
    assert np.linalg.norm(proj - op) / np.linalg.norm(op) < (1e-4 if single else 1e-10)      # the first 100+ columns span the range
