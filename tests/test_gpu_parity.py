"""Parity of the full pipelines against the oracle on the same A and the same Omega
(north_star): pivots / skeleton indices bit-exact wherever the pivot gap exceeds 1e-6 -- for all four scalar
types: the device takes its pivot decisions in double, so f32 / c32 are held to the sequence ?geqp3 picks in
double precision on the same single-precision data -- singular values, range residual and ID
error within 1e-10 relative (f64/c64) and 1e-4 (f32/c32).  A pivot tie never skips anything: the oracle is
replayed in the device's validated order (golden_common.adjudicate)."""
import numpy as np
import pytest

from golden_common import adjudicate
from oracle import reference_path as ref
from oracle.inputs import decaying_spectrum_matrix, helmholtz_kernel_matrix
from oracle.philox import random_gaussian

pytestmark = pytest.mark.gpu

RTOL = {np.float32: 1e-4, np.float64: 1e-10, np.complex64: 1e-4, np.complex128: 1e-10}


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a




@pytest.fixture(params=[0, 1], ids=["cholqr2-auto", "householder-tsqr"])
def qr_mode(request, api):
    api.default_context().set_option("qr_mode", request.param)
    yield request.param
    api.default_context().set_option("qr_mode", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
@pytest.mark.parametrize("it_count", [0, 2])
def test_rsvd_parity(api, dtype, it_count, qr_mode):
    """Config-2 pipeline at oracle-sized scale: sample_range_power_iteration ->
    SVD::compute_from_range_estimate (src/random_sampling.rs:131-160, src/svd.rs:171-183)."""
    m, n, k, p = 4096, 1024, 64, 10
    a, _sig = decaying_spectrum_matrix(m, n, dtype, seed=1234, r0=256, decade_every=16.0)
    omega = random_gaussian((n, k + p), dtype, seed=42)
    q_ref = ref.sample_range_power_iteration(a, k, p, it_count, ref.OmegaStream(dtype, blocks=[omega]))
    svd_ref = ref.SVD.compute_from_range_estimate(q_ref, a)
    op = api.DeviceMatrix.from_numpy(a)
    q_dev = api.sample_range_power_iteration(op, k, p, it_count, omega=omega)
    svd_dev = api.SVD.compute_from_range_estimate(q_dev, op)
    tol = RTOL[dtype]
    assert q_dev.shape == (m, k)
    assert np.max(np.abs(np.conj(q_dev.T).dot(q_dev) - np.eye(k))) < (1e-5 if tol > 1e-6 else 1e-12)
    res_ref, res_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
    assert abs(res_dev - res_ref) <= tol * res_ref, (res_dev, res_ref)
    s_ref, s_dev = svd_ref.s.astype(np.float64), svd_dev.s_f64()
    assert np.max(np.abs(s_dev - s_ref) / s_ref) <= tol, np.max(np.abs(s_dev - s_ref) / s_ref)
    rec_ref = ref.rel_diff_fro(svd_ref.to_mat(), a)
    rec_dev = ref.rel_diff_fro(svd_dev.to_mat(), a)
    assert abs(rec_dev - rec_ref) <= tol * rec_ref + (1e-6 if tol > 1e-6 else 0)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shifted", [1, 0], ids=["shifted-cholqr3", "householder"])
def test_rsvd_parity_steep_spectrum(api, dtype, shifted):
    """The operators the crate is written for decay fast: here sigma_j = 10^(-j/6), so the 74-column sketch spans
    twelve decades (cond ~1e12).  The speculative Cholesky-QR2 pass is rejected at its final check and the sampler is
    re-run on the shifted Cholesky-QR3 (or, with that off, on the Householder TSQR); either way the result matches the
    oracle: leading singular values to 1e-10 relative, the deep ones (down to 2e-11 of the first) and the range
    residual (1.5e-11) to the roundoff of the norm of A."""
    m, n, k, p, it_count = 4096, 1024, 64, 10, 2
    a, _sig = decaying_spectrum_matrix(m, n, dtype, seed=77, r0=128, decade_every=6.0)
    omega = random_gaussian((n, k + p), dtype, seed=42)
    q_ref = ref.sample_range_power_iteration(a, k, p, it_count, ref.OmegaStream(dtype, blocks=[omega]))
    svd_ref = ref.SVD.compute_from_range_estimate(q_ref, a)
    ctx = api.default_context()
    ctx.set_option("shifted_cholqr", shifted)
    try:
        ctx.reset_counters()
        op = api.DeviceMatrix.from_numpy(a)
        q_dev = api.sample_range_power_iteration(op, k, p, it_count, omega=omega)
        used_shifted = ctx.counter("cholqr_shifted")
        svd_dev = api.SVD.compute_from_range_estimate(q_dev, op)
    finally:
        ctx.set_option("shifted_cholqr", 1)
    assert ctx.counter("cholqr_fallbacks") >= 1
    assert (used_shifted >= 1) == bool(shifted)
    assert np.max(np.abs(np.conj(q_dev.T).dot(q_dev) - np.eye(k))) < 1e-12
    res_ref, res_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
    assert abs(res_dev - res_ref) <= 1e-3 * res_ref, (res_dev, res_ref)      # residual 1.5e-11 of ||A||: 1e-5 of it is roundoff
    s_ref, s_dev = svd_ref.s.astype(np.float64), svd_dev.s_f64()
    lead = s_ref >= 1e-5 * s_ref[0]
    assert np.max(np.abs(s_dev - s_ref)[lead] / s_ref[lead]) <= 1e-10
    assert np.max(np.abs(s_dev - s_ref)) <= 1e-14 * s_ref[0]


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
def test_column_and_two_sided_id_parity(api, dtype):
    """Config-5 pipeline at oracle scale: sample_range_by_rank -> QR::compute_from_range_estimate ->
    compress(RANK) -> column_id -> two_sided_id (src/qr.rs:270-323, src/col_interp_decomp.rs:116-125)."""
    m = n = 1500
    k, p = 48, 10
    a = helmholtz_kernel_matrix(m, n, dtype)
    omega = random_gaussian((n, k + p), dtype, seed=42)
    tol = RTOL[dtype]
    op = api.DeviceMatrix.from_numpy(a)
    q_dev = api.sample_range_by_rank(op, k, p, omega=omega)
    qr_dev = api.QR.compute_from_range_estimate(q_dev, op).compress(api.RANK(k))
    cid_dev = qr_dev.column_id()
    ts_dev = cid_dev.two_sided_id()
    q_ref = ref.sample_range_by_rank(a, k, p, ref.OmegaStream(dtype, blocks=[omega]))
    res_ref, res_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
    assert abs(res_dev - res_ref) <= tol * res_ref, (res_dev, res_ref)
    # the factor the device pivoted is b = (A^H Q_dev)^H from its own product; its skeleton columns must be the
    # double-precision ?geqp3 choice on that b wherever the gap exceeds 1e-6
    b_dev = ref.conj_t(op.conj_matmat(q_dev).to_numpy())
    qr_ref = ref.QR.compute_from_range_estimate(q_ref, a)
    order = adjudicate(b_dev, cid_dev.col_ind, qr_ref.ind, upto=k, label=f"{np.dtype(dtype).name} col_ind")
    if order is not None:
        qr_ref = ref.QR.compute_from_range_estimate(q_ref, a, order=order)
    cid_ref = qr_ref.compress(ref.RANK(k)).column_id()
    err_ref, err_dev = ref.rel_diff_fro(cid_ref.to_mat(), a), ref.rel_diff_fro(cid_dev.to_mat(), a)
    assert abs(err_dev - err_ref) <= tol * err_ref, (err_dev, err_ref)
    ts_ref = cid_ref.two_sided_id()
    order2 = adjudicate(ref.conj_t(cid_dev.c), ts_dev.row_ind, ts_ref.row_ind, upto=k, label=f"{np.dtype(dtype).name} row_ind")
    if order2 is not None:
        ts_ref = cid_ref.two_sided_id(order=order2)
    e2_ref, e2_dev = ref.rel_diff_fro(ts_ref.to_mat(), a), ref.rel_diff_fro(ts_dev.to_mat(), a)
    assert abs(e2_dev - e2_ref) <= 10 * tol * e2_ref, (e2_dev, e2_ref)
    # skeleton property: X[i, j] ~ A[row_ind[i], col_ind[j]]
    sk = a[np.ix_(ts_dev.row_ind[:k], ts_dev.col_ind[:k])]
    assert ref.rel_diff_fro(ts_dev.x, sk) < 100 * err_ref + 1e-3
    # Apply (src/col_interp_decomp.rs:141, src/two_sided_interp_decomp.rs:160)
    x = random_gaussian((n, 3), dtype, seed=9)
    assert ref.rel_diff_fro(cid_dev.dot(x), cid_ref.to_mat().dot(x)) < (1e-3 if tol > 1e-6 else 1e-9)
    assert ref.rel_diff_fro(ts_dev.dot(x[:, 0]), ts_ref.to_mat().dot(x[:, 0])) < (1e-3 if tol > 1e-6 else 1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
def test_row_id_route_and_apply_parity(api, dtype):
    """The row route: LQ::compute_from -> compress(RANK) -> row_id -> two_sided_id, and Apply for RowID
    (src/qr.rs:354-403, src/row_interp_decomp.rs:120-154): row skeleton, both errors, `dot` on a matrix and a vector."""
    m, n, k = 700, 500, 40
    a = helmholtz_kernel_matrix(m, n, dtype)
    tol = RTOL[dtype]
    name = np.dtype(dtype).name
    lq_dev, lq_ref = api.LQ.compute_from(a), ref.LQ.compute_from(a)
    order = adjudicate(ref.conj_t(a), lq_dev.ind, lq_ref.ind, upto=k, label=f"{name} LQ ind")
    if order is not None:
        lq_ref = ref.LQ.compute_from(a, order=order)
    rid_dev, rid_ref = lq_dev.compress(api.RANK(k)).row_id(), lq_ref.compress(ref.RANK(k)).row_id()
    assert rid_dev.x.shape == (m, k) and rid_dev.r.shape == (k, n) and len(rid_dev.row_ind) == m
    e_ref, e_dev = ref.rel_diff_fro(rid_ref.to_mat(), a), ref.rel_diff_fro(rid_dev.to_mat(), a)
    assert abs(e_dev - e_ref) <= tol * e_ref, (e_dev, e_ref)
    # X restricted to the skeleton rows is the identity (src/qr.rs:375-402)
    assert np.max(np.abs(rid_dev.x[rid_dev.row_ind[:k]] - np.eye(k))) < (1e-5 if tol > 1e-6 else 1e-12)
    # Apply: RowID . matrix and RowID . vector (rc_row_id_apply), against the oracle's x (r rhs)
    rhs = random_gaussian((n, 4), dtype, seed=5)
    want = rid_ref.dot(rhs)
    assert ref.rel_diff_fro(rid_dev.dot(rhs), want) < (2e-4 if tol > 1e-6 else 1e-9)
    assert ref.rel_diff_fro(rid_dev.dot(rhs[:, 1]), want[:, 1]) < (2e-4 if tol > 1e-6 else 1e-9)
    assert ref.rel_diff_fro(rid_dev.dot(rhs), rid_dev.to_mat().dot(rhs)) < (2e-5 if tol > 1e-6 else 1e-12)
    ts_dev, ts_ref = rid_dev.two_sided_id(), rid_ref.two_sided_id()
    order2 = adjudicate(rid_dev.r, ts_dev.col_ind, ts_ref.col_ind, upto=k, label=f"{name} RowID::two_sided_id col_ind")
    if order2 is not None:
        ts_ref = rid_ref.two_sided_id(order=order2)
    t_ref, t_dev = ref.rel_diff_fro(ts_ref.to_mat(), a), ref.rel_diff_fro(ts_dev.to_mat(), a)
    assert abs(t_dev - t_ref) <= 10 * tol * t_ref, (t_dev, t_ref)
    assert np.array_equal(ts_dev.row_ind, rid_dev.row_ind)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
def test_containers_from_parts(api, dtype):
    """QR / LQ / SVD assembled from parts (pub fields: src/qr.rs:31-51, src/svd.rs:13-20) behave like the
    factorisations they were taken from: rc_qr_new / rc_lq_new / rc_svd_new."""
    a = helmholtz_kernel_matrix(120, 90, dtype)
    tol = 1e-5 if RTOL[dtype] > 1e-6 else 1e-12
    qr = api.QR.compute_from(a)
    qr2 = api.QR.new(qr.q, qr.r, qr.ind)
    assert ref.rel_diff_fro(qr2.to_mat(), qr.to_mat()) < tol and np.array_equal(qr2.ind, qr.ind)
    c1, c2 = qr.compress(api.RANK(20)).column_id(), qr2.compress(api.RANK(20)).column_id()
    assert ref.rel_diff_fro(c2.to_mat(), c1.to_mat()) < tol and np.array_equal(c1.col_ind, c2.col_ind)
    # a QR the caller computed elsewhere (here: the oracle's) is accepted as is
    qo = ref.QR.compute_from(a)
    cid = api.QR.new(qo.q, qo.r, qo.ind).compress(api.RANK(20)).column_id()
    assert abs(ref.rel_diff_fro(cid.to_mat(), a) - ref.rel_diff_fro(qo.compress(ref.RANK(20)).column_id().to_mat(), a)) < 1e3 * tol
    lq = api.LQ.compute_from(a)
    lq2 = api.LQ.new(lq.l, lq.q, lq.ind)
    assert ref.rel_diff_fro(lq2.to_mat(), lq.to_mat()) < tol
    assert ref.rel_diff_fro(lq2.compress(api.RANK(20)).row_id().to_mat(), lq.compress(api.RANK(20)).row_id().to_mat()) < tol
    svd = api.SVD.compute_from(a)
    svd2 = api.SVD.new(svd.u, svd.s, svd.vt)
    assert ref.rel_diff_fro(svd2.to_mat(), svd.to_mat()) < tol and svd2.rank() == svd.rank()
    assert svd2.compress(api.ADAPTIVE(1e-2)).rank() == svd.compress(api.ADAPTIVE(1e-2)).rank()
    assert ref.rel_diff_fro(svd2.to_qr().to_mat(), svd.to_mat()) < 10 * tol
    with pytest.raises(AssertionError):
        api.QR.new(qr.q, qr.r[:-1], qr.ind)              # q.cols != r.rows
    with pytest.raises(AssertionError):
        api.SVD.new(svd.u, svd.s[:-1], svd.vt)


def test_adaptive_example_parity(api):
    """BASELINE config 1: examples/adaptive_sampling.rs (500 x 200 f64, rel_tol 1e-5, sample_size 5)
    on the same A and the same stream of Omega blocks as the oracle."""
    a = ref.random_approximate_low_rank_matrix((500, 200), 1.0, 1e-10, np.float64, seed=0)
    stream = ref.OmegaStream(np.float64, seed=11)
    q_ref, hist_ref = ref.sample_range_adaptive(a, 1e-5, 5, stream)
    q_dev, hist_dev = api.sample_range_adaptive(a, 1e-5, 5, omega_blocks=stream.drawn)
    assert [r for r, _ in hist_dev] == [r for r, _ in hist_ref]
    assert q_dev.shape == q_ref.shape and 100 <= q_dev.shape[1] <= 130
    for (_, e_dev), (_, e_ref) in zip(hist_dev, hist_ref):
        assert abs(e_dev - e_ref) <= 1e-6 * e_ref
    r_ref, r_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
    assert abs(r_dev - r_ref) <= 1e-8 * r_ref
    qr_dev = api.QR.compute_from_range_estimate(q_dev, a)
    qr_ref = ref.QR.compute_from_range_estimate(q_ref, a)
    e_ref, e_dev = ref.rel_diff_fro(a, qr_ref.to_mat()), ref.rel_diff_fro(a, qr_dev.to_mat())
    assert e_dev < 5e-5 and abs(e_dev - e_ref) <= 1e-8 * e_ref
    # device-generated Omega (Philox seed, block b = stream b) reproduces the oracle's seeded stream
    q_dev2, hist_dev2 = api.sample_range_adaptive(a, 1e-5, 5, seed=11)
    assert [r for r, _ in hist_dev2] == [r for r, _ in hist_ref]
    with pytest.raises(api.CompressionError):
        api.sample_range_adaptive(a, 1e-9, 5, seed=11, max_rank=20)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_adaptive_side_stream_overlap_matches_the_serial_order(api, dtype):
    """Single precision: the adaptive sampler launches the next sketch A Omega' on an auxiliary stream beside the
    projection + pivoted QR of the current one (option side_sms, SM budgets).  Same draws in the same order: the rank
    history equals that of the serial order (side_sms = 0) and of the oracle, the residual history agrees to f32
    roundoff of the products (their split-K factor follows the grid), the range residual to 1e-4."""
    m, n = 4096, 2048
    a, _ = decaying_spectrum_matrix(m, n, dtype, seed=21, r0=256, decade_every=32.0)
    stream = ref.OmegaStream(dtype, seed=5)
    q_ref, hist_ref = ref.sample_range_adaptive(a, 1e-3, 32, stream)
    ctx = api.default_context()
    out = {}
    try:
        for side in (0, 8):
            ctx.set_option("side_sms", side)
            out[side] = api.sample_range_adaptive(a, 1e-3, 32, omega_blocks=stream.drawn)
    finally:
        ctx.set_option("side_sms", 8)
    for side, (q_dev, hist_dev) in out.items():
        assert [r for r, _ in hist_dev] == [r for r, _ in hist_ref], side
        for (_, e_dev), (_, e_ref) in zip(hist_dev, hist_ref):
            assert abs(e_dev - e_ref) <= 1e-3 * e_ref + 1e-6
        r_ref, r_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
        assert abs(r_dev - r_ref) <= 1e-4 * r_ref + 1e-7, (side, r_dev, r_ref)


def test_device_generated_inputs_and_full_example(api):
    """examples/interpolative_decomposition.rs (500 x 100, k = 20) with the library's own generator."""
    a = api.random_approximate_low_rank_matrix((500, 100), 1.0, 1e-10, np.float64, seed=3)
    s = np.linalg.svd(a, compute_uv=False)
    assert abs(s[0] - 1.0) < 1e-10 and abs(s[-1] - 1e-10) < 1e-12
    ts = api.QR.compute_from(a).compress(api.RANK(20)).column_id().two_sided_id()
    ts_ref = ref.QR.compute_from(a).compress(ref.RANK(20)).column_id().two_sided_id()
    e, e_ref = ref.rel_diff_fro(a, ts.to_mat()), ref.rel_diff_fro(a, ts_ref.to_mat())
    assert np.array_equal(ts.col_ind[:20], ts_ref.col_ind[:20])
    assert abs(e - e_ref) <= 1e-8 * e_ref


def test_wide_sketch_panels(api):
    """Sketch wider than one shared-memory TSQR panel (l = 150 complex): block Gram-Schmidt over panels."""
    dtype = np.complex128
    a = helmholtz_kernel_matrix(1200, 900, dtype)
    k, p = 140, 10
    omega = random_gaussian((900, k + p), dtype, seed=1)
    q_ref = ref.sample_range_by_rank(a, k, p, ref.OmegaStream(dtype, blocks=[omega]))
    q_dev = api.sample_range_by_rank(a, k, p, omega=omega)
    assert np.max(np.abs(np.conj(q_dev.T).dot(q_dev) - np.eye(k))) < 1e-10
    r_ref, r_dev = ref.range_residual(a, q_ref), ref.range_residual(a, q_dev)
    assert abs(r_dev - r_ref) <= 1e-6 * r_ref + 1e-14
