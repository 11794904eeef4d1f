"""BASELINE config 1 / the reference's one published result: examples/adaptive_sampling.rs
on the 500x200 built-in low-rank matrix terminates at rank ~115 with an estimated
residual just below 1e-5 (residuals.png), and QR-from-range reproduces A to ~1e-5."""
import numpy as np

from oracle import reference_path as ref


def test_adaptive_example_matches_published_curve():
    mat = ref.random_approximate_low_rank_matrix((500, 200), 1.0, 1e-10, np.float64, seed=0)
    for route in ("gemv", "gemm"):
        stream = ref.OmegaStream(np.float64, seed=11)
        q, res = ref.sample_range_adaptive(mat, 1e-5, 5, stream, route=route)
        rank = q.shape[1]
        assert 100 <= rank <= 130 and rank % 5 == 0
        assert res[-1][0] == rank and res[-1][1] < 1e-5 and res[-2][1] >= 1e-5
        assert np.max(np.abs(q.T.dot(q) - np.eye(rank))) < 1e-10
        # exact residual tracks the estimate within ~an order of magnitude (residuals.png)
        exact = ref.range_residual(mat, q)
        assert exact < 1e-4
        qr = ref.QR.compute_from_range_estimate(q, mat, route=route)
        err = ref.rel_diff_fro(mat, qr.to_mat())
        assert err < 5e-5


def test_power_iteration_quirk_q1():
    """it_count >= 1 all give the it_count = 1 result (src/random_sampling.rs:144-154)."""
    mat = ref.random_approximate_low_rank_matrix((300, 120), 1.0, 1e-6, np.float64, seed=3)
    outs = []
    for it in (1, 2, 3):
        outs.append(ref.sample_range_power_iteration(mat, 20, 5, it, ref.OmegaStream(np.float64, seed=5)))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    q0 = ref.sample_range_power_iteration(mat, 20, 5, 0, ref.OmegaStream(np.float64, seed=5))
    qr = ref.sample_range_by_rank(mat, 20, 5, ref.OmegaStream(np.float64, seed=5))
    assert np.array_equal(q0, qr)
    assert ref.range_residual(mat, outs[0]) <= ref.range_residual(mat, q0) * 1.0001


def test_tolerance_compression_errors_when_unreachable():
    """Quirk Q3 (src/qr.rs:196-199, src/svd.rs:97-100)."""
    import pytest
    mat = np.eye(6)
    with pytest.raises(ref.CompressionError):
        ref.QR.compute_from(mat).compress(ref.ADAPTIVE(1e-3))
    with pytest.raises(ref.CompressionError):
        ref.SVD.compute_from(mat).compress(ref.ADAPTIVE(1e-3))
    with pytest.raises(AssertionError):
        ref.QR.compute_from(mat).compress(ref.ADAPTIVE(1.5))
