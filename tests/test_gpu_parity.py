"""Parity of the full pipelines against the oracle on the same A and the same Omega
(north_star): pivots / skeleton indices bit-exact wherever the pivot gap exceeds 1e-6, singular
values, range residual and ID error within 1e-10 relative (f64/c64) and 1e-4 (f32/c32)."""
import numpy as np
import pytest

from oracle import reference_path as ref
from oracle.inputs import decaying_spectrum_matrix, helmholtz_kernel_matrix
from oracle.philox import random_gaussian

pytestmark = pytest.mark.gpu

RTOL = {np.float32: 1e-4, np.float64: 1e-10, np.complex64: 1e-4, np.complex128: 1e-10}


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a


def check_indices(a_for_gaps, got, want, min_gap):
    got, want = np.asarray(got), np.asarray(want)
    if np.array_equal(got, want):
        return
    j = int(np.nonzero(got != want)[0][0])
    gaps = ref.pivot_gaps(a_for_gaps, want)
    assert j < len(gaps) and gaps[j] <= min_gap, f"pivot mismatch at step {j} with gap {gaps[j]:.3e}"
    pytest.skip(f"pivot tie at step {j} (gap {gaps[j]:.2e} <= {min_gap}): later quantities not comparable")


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


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_column_and_two_sided_id_parity(api, dtype):
    """Config-5 pipeline at oracle scale: sample_range_by_rank -> QR::compute_from_range_estimate ->
    compress(RANK) -> column_id -> two_sided_id (src/qr.rs:270-323, src/col_interp_decomp.rs:116-125)."""
    m = n = 1500
    k, p = 48, 10
    a = helmholtz_kernel_matrix(m, n, dtype)
    omega = random_gaussian((n, k + p), dtype, seed=42)
    q_ref = ref.sample_range_by_rank(a, k, p, ref.OmegaStream(dtype, blocks=[omega]))
    qr_ref = ref.QR.compute_from_range_estimate(q_ref, a).compress(ref.RANK(k))
    cid_ref = qr_ref.column_id()
    ts_ref = cid_ref.two_sided_id()
    op = api.DeviceMatrix.from_numpy(a)
    q_dev = api.sample_range_by_rank(op, k, p, omega=omega)
    qr_dev = api.QR.compute_from_range_estimate(q_dev, op).compress(api.RANK(k))
    cid_dev = qr_dev.column_id()
    ts_dev = cid_dev.two_sided_id()
    tol = RTOL[dtype]
    gap = 1e-6 if tol < 1e-6 else 1e-3
    b_ref = ref.conj_t(ref.DenseOperator(a).conj_matmat(q_ref))
    check_indices(b_ref, cid_dev.col_ind[:k], cid_ref.col_ind[:k], gap)
    err_ref, err_dev = ref.rel_diff_fro(cid_ref.to_mat(), a), ref.rel_diff_fro(cid_dev.to_mat(), a)
    assert abs(err_dev - err_ref) <= tol * err_ref, (err_dev, err_ref)
    check_indices(ref.conj_t(cid_ref.c), ts_dev.row_ind[:k], ts_ref.row_ind[:k], gap)
    e2_ref, e2_dev = ref.rel_diff_fro(ts_ref.to_mat(), a), ref.rel_diff_fro(ts_dev.to_mat(), a)
    assert abs(e2_dev - e2_ref) <= 10 * tol * e2_ref, (e2_dev, e2_ref)
    # skeleton property: X[i, j] ~ A[row_ind[i], col_ind[j]]
    sk = a[np.ix_(ts_dev.row_ind[:k], ts_dev.col_ind[:k])]
    assert ref.rel_diff_fro(ts_dev.x, sk) < 100 * err_ref + 1e-3
    # Apply (src/col_interp_decomp.rs:141, src/two_sided_interp_decomp.rs:160)
    x = random_gaussian((n, 3), dtype, seed=9)
    assert ref.rel_diff_fro(cid_dev.dot(x), cid_ref.to_mat().dot(x)) < (1e-3 if tol > 1e-6 else 1e-9)
    assert ref.rel_diff_fro(ts_dev.dot(x[:, 0]), ts_ref.to_mat().dot(x[:, 0])) < (1e-3 if tol > 1e-6 else 1e-8)


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
