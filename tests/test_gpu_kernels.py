"""Kernel-level parity through the C ABI: every CUDA building block against numpy / the oracle."""
import numpy as np
import pytest

from golden_common import adjudicate
from oracle import reference_path as ref
from oracle.philox import random_gaussian as oracle_gaussian

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64, np.complex64, np.complex128]
TOL = {np.float32: 2e-5, np.float64: 1e-13, np.complex64: 2e-5, np.complex128: 1e-13}


def rnd(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.standard_normal(shape)
    return a.astype(dtype)


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a


def relerr(x, y):
    return np.linalg.norm(x.astype(np.complex128) - y.astype(np.complex128)) / max(np.linalg.norm(y), 1e-300)


def test_library_loaded_and_counts_launches(api):
    ctx = api.default_context()
    ctx.reset_counters()
    api.DeviceMatrix.from_numpy(np.eye(4)).matmat(np.eye(4)).to_numpy()
    assert ctx.counter("kernel_launches") >= 1


@pytest.mark.parametrize("dtype", DTYPES)
def test_gaussian_matches_oracle_philox(api, dtype):
    want = oracle_gaussian((257, 33), dtype, seed=42, stream=3, row_offset=11)
    got = api.random_gaussian((257, 33), dtype, seed=42, stream=3, row_offset=11)
    assert got.dtype == want.dtype
    tol = 1e-6 if np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype).kind != "f" or dtype == np.float32 else 1e-13
    assert np.max(np.abs(got - want)) < (1e-6 if dtype in (np.float32, np.complex64) else 1e-13)


def test_permutation_known_answers(api):
    """src/permutation.rs:192-239 verbatim, through the device gather kernels."""
    mat = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    perm = np.array([2, 0, 1])
    assert np.array_equal(api.apply_permutation_matrix(mat, perm, "COL"), [[3, 1, 2], [6, 4, 5], [9, 7, 8]])
    assert np.array_equal(api.apply_permutation_matrix(mat, perm, "COLINV"), [[2, 3, 1], [5, 6, 4], [8, 9, 7]])
    assert np.array_equal(api.apply_permutation_matrix(mat, perm, "ROW"), [[7, 8, 9], [1, 2, 3], [4, 5, 6]])
    assert np.array_equal(api.apply_permutation_matrix(mat, perm, "ROWINV"), [[4, 5, 6], [7, 8, 9], [1, 2, 3]])
    vec = np.array([1.0, 2.0, 3.0])
    assert np.array_equal(api.apply_permutation_vector(vec, perm, "NOINV"), [3, 1, 2])
    assert np.array_equal(api.apply_permutation_vector(vec, perm, "INV"), [2, 3, 1])
    assert np.array_equal(api.invert_permutation_vector(perm), [1, 2, 0])
    with pytest.raises(AssertionError):
        api.apply_permutation_matrix(mat, np.array([0, 1]), "COL")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(300, 130, 17), (64, 64, 64), (1000, 257, 5), (33, 2100, 9)])
def test_matmat_and_conj_matmat(api, dtype, shape):
    m, n, l = shape
    a = rnd((m, n), dtype, 1)
    x = rnd((n, l), dtype, 2)
    y = rnd((m, l), dtype, 3)
    op = api.DeviceMatrix.from_numpy(a)
    assert relerr(op.matmat(x).to_numpy(), a.dot(x)) < TOL[dtype] * 10
    assert relerr(op.conj_matmat(y).to_numpy(), np.conj(a.T).dot(y)) < TOL[dtype] * 10


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(4096, 1024, 74), (1000, 514, 10), (777, 1030, 138), (2048, 512, 96), (130, 4096, 20)])
def test_dmma_tma_gemm_matches_generic_and_numpy(api, dtype, shape):
    """The TMA + DMMA tensor-pipe kernels (gemm_dmma.cu) against numpy and the SIMT kernel."""
    m, n, l = shape
    a = rnd((m, n), dtype, 4)
    x = rnd((n, l), dtype, 5)
    y = rnd((m, l), dtype, 6)
    ctx = api.default_context()
    op = api.DeviceMatrix.from_numpy(a)
    ctx.set_option("gemm_impl", 0)
    y_fast = op.matmat(x).to_numpy()
    z_fast = op.conj_matmat(y).to_numpy()
    ctx.set_option("gemm_impl", 1)
    y_gen = op.matmat(x).to_numpy()
    z_gen = op.conj_matmat(y).to_numpy()
    ctx.set_option("gemm_impl", 0)
    assert relerr(y_fast, a.dot(x)) < 1e-13 and relerr(y_gen, a.dot(x)) < 1e-13
    assert relerr(z_fast, np.conj(a.T).dot(y)) < 1e-13 and relerr(z_gen, np.conj(a.T).dot(y)) < 1e-13


@pytest.mark.parametrize("l", [2, 4, 10, 12, 26, 44, 74, 76, 90, 92])
def test_dmma_ragged_last_column_group_on_dfma_tail(api, l):
    """l % 8 in {2, 4} with a single column chunk: the last 8-column group is formed with DFMAs by the warps that
    own it (gemm_dmma.cu, TAILW); checked column by column against numpy and against the all-DMMA path, for
    Y = A X, Z = A^T Y (split-K) and the c64 real expansion, with row counts that leave partial 64-row tiles."""
    ctx = api.default_context()
    m, n = 1000 + l, 520
    a, x, y = rnd((m, n), np.float64, 7), rnd((n, l), np.float64, 8), rnd((m, l), np.float64, 9)
    az, xz = rnd((m, n // 2), np.complex128, 10), rnd((n // 2, l), np.complex128, 11)
    op, opz = api.DeviceMatrix.from_numpy(a), api.DeviceMatrix.from_numpy(az)
    out = {}
    try:
        for tail in (0, 1):
            ctx.set_option("dmma_tail", tail)
            out[tail] = (op.matmat(x).to_numpy(), op.conj_matmat(y).to_numpy(), opz.matmat(xz).to_numpy())
    finally:
        ctx.set_option("dmma_tail", 1)
    refs = (a.dot(x), a.T.dot(y), az.dot(xz))
    for tail in (0, 1):
        for got, want in zip(out[tail], refs):
            assert got.shape == want.shape
            scale = np.max(np.abs(want))
            assert np.max(np.abs(got - want)) < 1e-13 * scale * np.sqrt(max(m, n)), (tail, l)
    # columns outside the tail group are produced by the same DMMA sequence in both modes
    keep = (l - 1) // 8 * 8
    assert np.array_equal(out[0][0][:, :keep], out[1][0][:, :keep])


@pytest.mark.parametrize("shape", [(4096, 1024, 74), (1000, 516, 10), (777, 1028, 138), (2048, 512, 266), (130, 4096, 20),
                                   # more work items than SMs (every CTA crosses item boundaries, with and without
                                   # column chunks): the regime in which a too-early release of the raw A slot showed
                                   (4096, 128, 2048), (8192, 4096, 320), (32768, 256, 64), (20000, 96, 200)])
def test_tcgen05_tf32x3_gemm_matches_f64_reference(api, shape):
    """f32 Y = A X on tcgen05 (kind::tf32, 3-product split, TMEM accumulators; gemm_tf32.cu) against a
    float64 reference and against the SIMT kernel: f32-level accuracy, far inside the 1e-4 of north_star."""
    m, n, l = shape
    a = rnd((m, n), np.float32, 14)
    x = rnd((n, l), np.float32, 15)
    want = a.astype(np.float64).dot(x.astype(np.float64))
    ctx = api.default_context()
    op = api.DeviceMatrix.from_numpy(a)
    ctx.set_option("gemm_impl", 0)
    y_tc = op.matmat(x).to_numpy()
    ctx.set_option("gemm_impl", 1)
    y_simt = op.matmat(x).to_numpy()
    ctx.set_option("gemm_impl", 0)
    e_tc, e_simt = relerr(y_tc, want), relerr(y_simt, want)
    assert e_simt < 2e-6
    assert e_tc < 5e-6, (e_tc, e_simt)
    # Z = A^T Y: MN-major UMMA operands, split-K over the rows of A
    yy = rnd((m, l), np.float32, 16)
    want_t = a.astype(np.float64).T.dot(yy.astype(np.float64))
    z_tc = op.conj_matmat(yy).to_numpy()
    ctx.set_option("gemm_impl", 1)
    z_simt = op.conj_matmat(yy).to_numpy()
    ctx.set_option("gemm_impl", 0)
    assert relerr(z_simt, want_t) < 2e-6
    assert relerr(z_tc, want_t) < 5e-6, relerr(z_tc, want_t)


@pytest.mark.parametrize("shape", [(2048, 512, 74), (1000, 258, 20), (4096, 64, 1024), (4096, 1024, 266)])
def test_c32_runs_on_tcgen05_through_real_expansion(api, shape):
    """c32 contractions reuse the tcgen05 TF32x3 kernels through the exact real expansion."""
    m, n, l = shape
    a, x, y = rnd((m, n), np.complex64, 21), rnd((n, l), np.complex64, 22), rnd((m, l), np.complex64, 23)
    op = api.DeviceMatrix.from_numpy(a)
    want = a.astype(np.complex128).dot(x.astype(np.complex128))
    want_h = np.conj(a.astype(np.complex128).T).dot(y.astype(np.complex128))
    assert relerr(op.matmat(x).to_numpy(), want) < 5e-6
    assert relerr(op.conj_matmat(y).to_numpy(), want_h) < 5e-6


def test_cholqr2_falls_back_on_ill_conditioned_panels(api):
    """A sketch with condition number ~1e12 is rejected by the plain Cholesky-QR2 and taken by the shifted
    Cholesky-QR3 (three GEMM-shaped rounds); with that route off, and beyond its reach (cond ~1e15), it goes to
    the Householder TSQR.  Orthonormal and backward stable to roundoff on every route."""
    ctx = api.default_context()
    a = ref.random_approximate_low_rank_matrix((3000, 60), 1.0, 1e-12, np.float64, seed=9)
    ctx.reset_counters()
    q, r, ind = api.pivoted_qr(a)
    assert ctx.counter("cholqr_fallbacks") == 1 and ctx.counter("cholqr_shifted") == 1 and ctx.counter("cholqr_used") == 1
    assert np.max(np.abs(q.T.dot(q) - np.eye(60))) < 1e-13
    assert relerr(q.dot(r), a[:, ind]) < 1e-13
    q0, r0, ind0 = ref.pivoted_qr(a)
    order = adjudicate(a, ind, ind0, label="shifted cholqr3 3000x60")
    if order is not None:
        q0, r0, ind0 = ref.pivoted_qr_with_order(a, order)
    assert np.max(np.abs(np.abs(np.diagonal(r)) - np.abs(np.diagonal(r0))) / np.abs(r0[0, 0])) < 1e-13
    ctx.set_option("shifted_cholqr", 0)
    try:
        ctx.reset_counters()
        q, r, ind = api.pivoted_qr(a)
        assert ctx.counter("cholqr_fallbacks") >= 1 and ctx.counter("cholqr_used") == 0
        assert np.max(np.abs(q.T.dot(q) - np.eye(60))) < 1e-13
        assert relerr(q.dot(r), a[:, ind]) < 1e-13
    finally:
        ctx.set_option("shifted_cholqr", 1)
    hard = ref.random_approximate_low_rank_matrix((3000, 60), 1.0, 1e-15, np.float64, seed=9)
    ctx.reset_counters()
    q, r, ind = api.pivoted_qr(hard)
    assert np.max(np.abs(q.T.dot(q) - np.eye(60))) < 1e-13
    assert relerr(q.dot(r), hard[:, ind]) < 1e-13
    well = ref.random_approximate_low_rank_matrix((3000, 60), 1.0, 1e-3, np.float64, seed=9)
    ctx.reset_counters()
    q, r, ind = api.pivoted_qr(well)
    assert ctx.counter("cholqr_used") == 1 and ctx.counter("cholqr_shifted") == 0
    assert np.max(np.abs(q.T.dot(q) - np.eye(60))) < 1e-13


def test_strided_views_upload(api):
    a = rnd((40, 30), np.float64, 7)
    assert np.array_equal(api.DeviceMatrix.from_numpy(a.T).to_numpy(), a.T)
    assert np.array_equal(api.DeviceMatrix.from_numpy(a[::2, ::3]).to_numpy(), a[::2, ::3])


@pytest.mark.parametrize("dtype", DTYPES)
def test_norms_and_rel_diff(api, dtype):
    a = rnd((500, 37), dtype, 8)
    b = a + 1e-3 * rnd((500, 37), dtype, 9)
    assert abs(api.rel_diff_fro(a, b) - ref.rel_diff_fro(a, b)) < 1e-6 * ref.rel_diff_fro(a, b) + 1e-12
    assert abs(api.max_col_norm(a) - ref.max_col_norm(a)) < 1e-5 * ref.max_col_norm(a)
    assert abs(api.rel_diff_l2(a[:, 0], b[:, 0]) - ref.rel_diff_l2(a[:, 0], b[:, 0])) < 1e-6


@pytest.mark.parametrize("qr_mode", [0, 1], ids=["cholqr2-auto", "householder-tsqr"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2000, 40), (300, 74), (5000, 7), (64, 64), (150, 100), (8192, 200)])
def test_tall_pivoted_qr_matches_lapack(api, dtype, shape, qr_mode):
    """Tall route (Cholesky-QR2 fast path with fallback / Householder TSQR) + pivot-on-R against
    ?geqp3/?orgqr: same pivots, |R| equal, Q orthonormal, Q R = A P."""
    a = ref.random_approximate_low_rank_matrix(shape, 1.0, 1e-4, dtype, seed=5)
    api.default_context().set_option("qr_mode", qr_mode)
    try:
        q, r, ind = api.pivoted_qr(a)
    finally:
        api.default_context().set_option("qr_mode", 0)
    q0, r0, ind0 = ref.pivoted_qr(a)
    tol = 2e-4 if dtype in (np.float32, np.complex64) else 1e-9
    k = min(shape)
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < (1e-5 if tol > 1e-6 else 1e-13)
    assert relerr(q.dot(r), a[:, ind]) < (1e-5 if tol > 1e-6 else 1e-13)
    # pivots: the double-precision ?geqp3 sequence wherever the gap exceeds 1e-6 (all four scalars); on a tie the
    # oracle is replayed in the device's order, nothing is skipped
    order = adjudicate(a, ind, ind0, label=f"tall {shape} {np.dtype(dtype).name}")
    if order is not None:
        q0, r0, ind0 = ref.pivoted_qr_with_order(a, order)
    assert np.max(np.abs(np.abs(np.diagonal(r)) - np.abs(np.diagonal(r0))) / np.abs(r0[0, 0])) < tol


@pytest.mark.parametrize("cluster", [1, 0], ids=["cluster", "cooperative"])
@pytest.mark.parametrize("dtype,shape", [(np.float64, (3000, 200)), (np.complex128, (2000, 138)), (np.float32, (4000, 266)),
                                         (np.float64, (260, 300)), (np.complex64, (1500, 150))],
                         ids=["f64-200", "c64-138", "f32-266", "f64-260x300", "c32-150"])
def test_medium_pivoted_qr_cluster_route(api, dtype, shape, cluster):
    """The w x w triangle of a wide sketch (config 5: 138 x 138 c64, config 4: 266 x 266 in double) does not fit one CTA:
    thread-block-cluster kernel (columns in distributed shared memory, one cluster barrier per step) against the
    cooperative grid kernel and against ?geqp3."""
    a = ref.random_approximate_low_rank_matrix(shape, 1.0, 1e-3, dtype, seed=15)
    ctx = api.default_context()
    ctx.set_option("cluster_qr", cluster)
    try:
        q, r, ind = api.pivoted_qr(a)
    finally:
        ctx.set_option("cluster_qr", 1)
    single = dtype in (np.float32, np.complex64)
    k = min(shape)
    assert sorted(ind.tolist()) == list(range(shape[1]))
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < (2e-5 if single else 1e-12)
    assert relerr(q.dot(r), a[:, ind]) < (1e-5 if single else 1e-13)
    q0, r0, ind0 = ref.pivoted_qr(a)
    order = adjudicate(a, ind, ind0, label=f"medium {shape} {np.dtype(dtype).name} cluster={cluster}")
    if order is not None:
        q0, r0, ind0 = ref.pivoted_qr_with_order(a, order)
    assert np.max(np.abs(np.abs(np.diagonal(r)) - np.abs(np.diagonal(r0))) / np.abs(r0[0, 0])) < (2e-4 if single else 1e-9)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(20, 700), (64, 3000), (50, 100)])
def test_wide_pivoted_qr_matches_lapack(api, dtype, shape):
    """Cooperative pivoted Householder kernel (short-wide k x n factor) against ?geqp3."""
    a = ref.random_approximate_low_rank_matrix(shape, 1.0, 1e-3, dtype, seed=6)
    q, r, ind = api.pivoted_qr(a)
    q0, r0, ind0 = ref.pivoted_qr(a)
    single = dtype in (np.float32, np.complex64)
    k = min(shape)
    assert sorted(ind.tolist()) == list(range(shape[1]))
    assert np.max(np.abs(np.conj(q.T).dot(q) - np.eye(k))) < (1e-5 if single else 1e-13)
    assert relerr(q.dot(r), a[:, ind]) < (1e-5 if single else 1e-13)
    order = adjudicate(a, ind, ind0, label=f"wide {shape} {np.dtype(dtype).name}")
    if order is not None:
        q0, r0, ind0 = ref.pivoted_qr_with_order(a, order)
    assert np.max(np.abs(np.abs(np.diagonal(r)) - np.abs(np.diagonal(r0))) / np.abs(r0[0, 0])) < (2e-4 if single else 1e-9)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(64, 2000), (300, 40), (50, 50), (33, 7)])
def test_svd_matches_gesdd(api, dtype, shape):
    a = ref.random_approximate_low_rank_matrix(shape, 1.0, 1e-5, dtype, seed=7)
    u, s, vt = api.compute_svd(a)
    u0, s0, vt0 = ref.compute_svd(a)
    single = dtype in (np.float32, np.complex64)
    k = min(shape)
    assert u.shape == u0.shape and vt.shape == vt0.shape and s.shape == s0.shape
    assert np.all(np.diff(s) <= 0)
    assert np.max(np.abs(s - s0)) / s0[0] < (1e-5 if single else 1e-13)
    assert relerr((u * s).dot(vt), a) < (2e-5 if single else 1e-12)
    assert np.max(np.abs(np.conj(u.T).dot(u) - np.eye(k))) < (2e-4 if single else 1e-10)
    assert np.max(np.abs(vt.dot(np.conj(vt.T)) - np.eye(k))) < (2e-4 if single else 1e-10)


@pytest.mark.parametrize("dtype", [np.complex128, np.float32])
def test_device_helmholtz_generator_matches_host_mirror(api, dtype):
    """rc_helmholtz_kernel_matrix (config-5 input, generated on device) against oracle.inputs' Philox mirror,
    including a row shard."""
    from oracle.inputs import helmholtz_kernel_matrix_philox
    want = helmholtz_kernel_matrix_philox(300, 200, dtype, seed=7)
    got = api.helmholtz_kernel_matrix((300, 200), dtype, seed=7).to_numpy()
    tol = 1e-12 if np.dtype(dtype).itemsize >= 16 else 1e-5
    assert relerr(got, want) < tol
    shard = api.helmholtz_kernel_matrix((100, 200), dtype, seed=7, row_offset=150).to_numpy()
    assert relerr(shard, want[150:250]) < tol


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
@pytest.mark.parametrize("shape", [(4096, 1024, 74), (4096, 2048, 266), (2048, 512, 64), (1000, 333, 130)])
def test_bf16_opt_in_contraction(api, dtype, shape):
    """north_star (2): "tcgen05 ... for bf16 where the caller opts in" -- context option f32_precision = 1 runs the f32 / c32
    contractions (MatMat / ConjMatMat, src/types.rs:58-101) as ONE bf16 product (tcgen05 kind::f16, FP32 accumulation)
    instead of the 3-product TF32 split.  Tolerance: bf16 keeps 8 significant bits (unit roundoff 2^-9 = 2e-3) and both
    operands are rounded, so the product differs from the f64 reference by ~2^-9 relative in the Frobenius norm:
    accepted below 1e-2, and required to be ABOVE 1e-4 -- the default path is at ~1e-6, so a result that accurate would
    mean the option was ignored."""
    m, k, n = shape
    ctx = api.default_context()
    a, x, w = rnd((m, k), dtype, 1), rnd((k, n), dtype, 2), rnd((m, n), dtype, 3)
    ad = api.DeviceMatrix.from_numpy(a)
    wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    y_ref, z_ref = a.astype(wide).dot(x.astype(wide)), np.conj(a.T).astype(wide).dot(w.astype(wide))
    ctx.set_option("f32_precision", 1)
    try:
        y16, z16 = ad.matmat(x).to_numpy(), ad.conj_matmat(w).to_numpy()
    finally:
        ctx.set_option("f32_precision", 0)
    y32, z32 = ad.matmat(x).to_numpy(), ad.conj_matmat(w).to_numpy()
    for got16, got32, want, what in ((y16, y32, y_ref, "A X"), (z16, z32, z_ref, "A^H W")):
        e16, e32 = relerr(got16, want), relerr(got32, want)
        assert e32 < 1e-5, (what, e32)
        assert 1e-4 < e16 < 1e-2, (what, e16)
    with pytest.raises(AssertionError):
        ctx.set_option("f32_precision", 7)


def test_bf16_opt_in_range_finder(api):
    """The opt-in through a pipeline: fixed-rank range finder on an f32 operator with bf16 contractions still captures
    the range to the accuracy the format allows (residual within 3e-2 of the f32 path's, here ~1e-2 absolute)."""
    from oracle.inputs import decaying_spectrum_matrix
    a, _ = decaying_spectrum_matrix(4096, 1024, np.float32, seed=4, r0=128, decade_every=16.0)
    ctx = api.default_context()
    op = api.DeviceMatrix.from_numpy(a)
    q32 = api.sample_range_by_rank(op, 32, 8, seed=1)
    ctx.set_option("f32_precision", 1)
    try:
        q16 = api.sample_range_by_rank(op, 32, 8, seed=1)
    finally:
        ctx.set_option("f32_precision", 0)
    r32, r16 = ref.range_residual(a, q32), ref.range_residual(a, q16)
    assert np.max(np.abs(q16.T.astype(np.float64).dot(q16.astype(np.float64)) - np.eye(32))) < 1e-4
    assert r16 < r32 + 3e-2 and r16 < 5e-2, (r16, r32)
