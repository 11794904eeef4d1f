"""Edge cases through the C ABI: degenerate shapes, exactly rank-deficient and zero matrices, oversampling
beyond the matrix size, argument violations (where the crate `assert!`s the ABI returns INVALID_ARGUMENT,
surfaced as AssertionError by the Python mirror)."""
import numpy as np
import pytest

from oracle import reference_path as ref
from oracle.philox import random_gaussian

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a


def rnd(shape, dtype, seed):
    return random_gaussian(shape, dtype, seed)


def check_qr(a, qr, tol):
    q, r, ind = qr.q, qr.r, np.asarray(qr.ind)
    k = min(a.shape)
    assert q.shape == (a.shape[0], k) and r.shape == (k, a.shape[1]) and sorted(ind.tolist()) == list(range(a.shape[1]))
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < tol
    scale = max(np.max(np.abs(a)), 1e-300)
    assert np.max(np.abs(q.dot(r) - a[:, ind])) < tol * scale * max(a.shape)
    d = np.abs(np.diag(r))
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-3) - tol * scale)          # non-increasing diagonal


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1), (1, 17), (17, 1), (2, 2), (3, 40), (40, 3), (129, 2), (2, 129)])
def test_degenerate_shapes(api, dtype, shape):
    a = rnd(shape, dtype, 5)
    tol = 1e-5 if np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype).kind != "f" or dtype == np.float32 else 1e-12
    tol = 2e-5 if dtype in (np.float32, np.complex64) else 1e-12
    check_qr(a, api.QR.compute_from(a), tol)
    lq = api.LQ.compute_from(a)
    assert np.max(np.abs(lq.l.dot(lq.q) - a[np.asarray(lq.ind), :])) < tol * max(a.shape) * np.max(np.abs(a))
    svd = api.SVD.compute_from(a)
    s_ref = np.linalg.svd(a.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64), compute_uv=False)
    assert np.max(np.abs(svd.s_f64() - s_ref)) < tol * s_ref[0] * 10
    assert ref.rel_diff_fro(svd.to_mat(), a) < tol * 10


@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
def test_zero_and_rank_deficient_matrices(api, dtype):
    z = np.zeros((50, 30), dtype=dtype)
    qr = api.QR.compute_from(z)
    assert np.all(np.isfinite(qr.q)) and np.all(qr.r == 0) and sorted(np.asarray(qr.ind).tolist()) == list(range(30))
    assert np.all(api.SVD.compute_from(z).s_f64() == 0)
    with pytest.raises(api.CompressionError):                      # quirk Q3: |r_ii / r_00| is NaN, never < tol
        qr.compress(api.ADAPTIVE(1e-3))
    # exactly rank 5: duplicate / combined columns
    b = rnd((60, 5), dtype, 1).dot(rnd((5, 40), dtype, 2))
    tol = 1e-4 if dtype == np.complex64 else 1e-10
    qr = api.QR.compute_from(b)
    d = np.abs(np.diag(qr.r))
    assert d[4] > 1e-3 * d[0] and d[5] < (1e-4 if dtype == np.complex64 else 1e-12) * d[0]
    assert qr.compress(api.ADAPTIVE(tol)).rank() == 5
    cid = qr.compress(api.RANK(5)).column_id()
    assert ref.rel_diff_fro(cid.to_mat(), b) < tol
    ts = cid.two_sided_id()
    assert ref.rel_diff_fro(ts.to_mat(), b) < 10 * tol
    svd = api.SVD.compute_from(b).compress(api.ADAPTIVE(tol))
    assert svd.rank() == 5 and ref.rel_diff_fro(svd.to_mat(), b) < tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oversampling_beyond_the_matrix(api, dtype):
    """k + p larger than the number of columns / rows: the crate clamps through min(m, n) in pivoted_qr
    (src/pivoted_qr.rs:85) and compress(RANK) (src/qr.rs:169-184)."""
    a = rnd((40, 12), dtype, 3)
    omega = rnd((12, 20), dtype, 4)                                 # l = 20 > n = 12
    q = api.sample_range_by_rank(a, 15, 5, omega=omega)
    q_ref = ref.sample_range_by_rank(a, 15, 5, ref.OmegaStream(dtype, blocks=[omega]))
    assert q.shape == q_ref.shape
    tol = 1e-4 if dtype == np.float32 else 1e-10
    assert ref.range_residual(a, q) < tol and ref.range_residual(a, q_ref) < tol
    wide = rnd((6, 50), dtype, 6)                                   # m = 6 < l = 10
    om = rnd((50, 10), dtype, 7)
    qw = api.sample_range_by_rank(wide, 8, 2, omega=om)
    assert qw.shape == ref.sample_range_by_rank(wide, 8, 2, ref.OmegaStream(dtype, blocks=[om])).shape == (6, 6)
    assert np.max(np.abs(qw.T.dot(qw) - np.eye(6))) < 10 * tol


def test_argument_violations(api):
    a = rnd((30, 20), np.float64, 8)
    with pytest.raises(AssertionError):
        api.sample_range_by_rank(a, 0, 2, seed=1)                  # k must be positive
    with pytest.raises(AssertionError):
        api.sample_range_by_rank(a, 4, 2, omega=rnd((19, 6), np.float64, 1))     # Omega with the wrong row count
    with pytest.raises(AssertionError):
        api.QR.compute_from(a).compress(api.ADAPTIVE(1.5))         # assert!(tol < 1) (src/qr.rs:188)
    with pytest.raises(AssertionError):
        api.apply_permutation_matrix(a, np.arange(7), "COL")       # length mismatch (src/permutation.rs:96-99)
    with pytest.raises(AssertionError):
        api.DeviceMatrix.from_numpy(a).matmat(rnd((21, 3), np.float64, 2))
    with pytest.raises(AssertionError):
        api.rel_diff_fro(a, a[:, :5])
    with pytest.raises(api.CompressionError):
        api.SVD.compute_from(np.eye(8)).compress(api.ADAPTIVE(1e-3))   # no singular value below tol (quirk Q3)
    q, hist = api.sample_range_adaptive(a, 1e-12, 5, seed=2, max_rank=20)     # reaches full rank, then stops
    assert q.shape[1] <= 20 and hist[-1][1] < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
def test_pipelined_upload(api, dtype):
    """rc_matrix_from_host_async / rc_matrix_await: two uploads in flight, results identical to the blocking route;
    a handle freed without ever being awaited is safe."""
    import torch
    ctx = api.default_context()
    td = torch.float64 if dtype == np.float64 else torch.complex64
    hosts = [torch.empty((700, 300), dtype=td, pin_memory=True).numpy() for _ in range(3)]
    for i, h in enumerate(hosts):
        h[...] = rnd(h.shape, dtype, 20 + i)
    padded = torch.empty((700, 320), dtype=td, pin_memory=True).numpy()
    padded[:, :300] = hosts[0]
    ops = [api.DeviceMatrix.from_numpy_async(h, ctx=ctx) for h in hosts]
    view = api.DeviceMatrix.from_numpy_async(padded[:, :300], ctx=ctx)           # row stride > cols
    never = api.DeviceMatrix.from_numpy_async(hosts[1], ctx=ctx)
    never.free()
    omega = rnd((300, 12), dtype, 3)
    for op, h in zip(ops, hosts):
        op.await_upload()
        assert np.array_equal(op.to_numpy(), h)
        y = op.matmat(api.DeviceMatrix.from_numpy(omega, ctx=ctx)).to_numpy()
        y_ref = api.DeviceMatrix.from_numpy(h, ctx=ctx).matmat(api.DeviceMatrix.from_numpy(omega, ctx=ctx)).to_numpy()
        assert np.array_equal(y, y_ref)
    assert np.array_equal(view.await_upload(block_host=True).to_numpy(), hosts[0])
    with pytest.raises(ValueError):
        api.DeviceMatrix.from_numpy_async(hosts[0].T, ctx=ctx)
    # memory the caller owns (an ordinary numpy allocation), page-locked in place
    own = np.ascontiguousarray(rnd((513, 129), dtype, 31))
    ctx.pin(own)
    try:
        assert np.array_equal(api.DeviceMatrix.from_numpy_async(own, ctx=ctx).await_upload(block_host=True).to_numpy(), own)
    finally:
        ctx.unpin(own)
