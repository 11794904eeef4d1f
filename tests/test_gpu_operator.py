"""Matrix-free operators through the plugin API: the crate's samplers and
compute_from_range_estimate are generic over MatMat / ConjMatMat implemented by the caller
(src/types.rs:40-101, src/random_sampling.rs:102,130,222, src/qr.rs:221-224, src/svd.rs:110-113).
Here the caller's implementation is a device callback handed to rc_operator_create."""
import numpy as np
import pytest

from oracle import reference_path as ref
from oracle.inputs import decaying_spectrum_matrix
from oracle.philox import random_gaussian

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from rusty_compression_b200 import api as a
    return a


def test_operator_wrapping_a_dense_matrix_matches_the_dense_path(api):
    """Callbacks that multiply with a hidden dense device matrix: every pipeline must give what the dense
    operator gives (same kernels underneath, so to roundoff) and match the oracle."""
    m, n, k, p = 3000, 700, 32, 8
    a, _ = decaying_spectrum_matrix(m, n, np.float64, seed=5, r0=128, decade_every=10.0)
    dense = api.DeviceMatrix.from_numpy(a)
    calls = {"matmat": 0, "conj": 0}

    def matmat(x, ncols, y, stream):
        calls["matmat"] += 1
        y.copy_from(dense.matmat(x))

    def conj_matmat(x, ncols, z, stream):
        calls["conj"] += 1
        z.copy_from(dense.conj_matmat(x))

    op = api.Operator((m, n), np.float64, matmat, conj_matmat)
    assert op.shape == (m, n)
    omega = random_gaussian((n, k + p), np.float64, seed=42)
    q_op = api.sample_range_power_iteration(op, k, p, 2, omega=omega)
    q_dense = api.sample_range_power_iteration(dense, k, p, 2, omega=omega)
    assert calls["matmat"] == 3 and calls["conj"] == 2          # 1 + 2 it products with A, it with A^H
    assert np.max(np.abs(q_op - q_dense)) < 1e-12
    s_op = api.SVD.compute_from_range_estimate(q_op, op).s_f64()
    q_ref = ref.sample_range_power_iteration(a, k, p, 2, ref.OmegaStream(np.float64, blocks=[omega]))
    s_ref = ref.SVD.compute_from_range_estimate(q_ref, a).s
    assert np.max(np.abs(s_op - s_ref) / s_ref) < 1e-10
    # column ID from the operator, adaptive sampler with the Philox stream
    qr = api.QR.compute_from_range_estimate(api.sample_range_by_rank(op, k, p, omega=omega), op)
    cid = qr.compress(api.RANK(k)).column_id()
    cid_ref = ref.QR.compute_from_range_estimate(ref.sample_range_by_rank(a, k, p, ref.OmegaStream(np.float64, blocks=[omega])), a).compress(ref.RANK(k)).column_id()
    assert np.array_equal(cid.col_ind[:k], cid_ref.col_ind[:k])
    e, e_ref = ref.rel_diff_fro(cid.to_mat(), a), ref.rel_diff_fro(cid_ref.to_mat(), a)
    assert abs(e - e_ref) <= 1e-10 * e_ref
    qa, hist = api.sample_range_adaptive(op, 1e-6, 16, seed=3)
    qa_ref, hist_ref = ref.sample_range_adaptive(a, 1e-6, 16, ref.OmegaStream(np.float64, seed=3))
    assert [r for r, _ in hist] == [r for r, _ in hist_ref]
    assert abs(ref.range_residual(a, qa) - ref.range_residual(a, qa_ref)) <= 1e-8 * ref.range_residual(a, qa_ref)


def test_never_materialised_kernel_operator(api):
    """The config-5 operator A_ij = exp(i kappa |x_i - y_j|) / |x_i - y_j| evaluated block by block inside the
    callbacks (torch as plumbing on the raw device buffers, on the library's stream): A is never stored."""
    import torch
    m, n, k, p, kappa = 2000, 1800, 40, 8, 20.0
    rng = np.random.default_rng(7)
    xs, ys = rng.random((m, 3)), rng.random((n, 3))
    ys[:, 0] += 1.5
    xt, yt = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()

    class Raw:                                   # __cuda_array_interface__ view of a (rows x cols, ld) buffer
        def __init__(self, dm):
            rows, cols = dm.shape
            self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<c16", "data": (dm.device_ptr, False),
                                             "version": 3, "strides": (dm.ld * 16, 16)}

    def kernel_block(i0, i1):
        d = torch.cdist(xt[i0:i1], yt)
        return torch.exp(1j * kappa * d) / d

    def matmat(x, ncols, y, stream):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            xv, yv = torch.as_tensor(Raw(x), device="cuda"), torch.as_tensor(Raw(y), device="cuda")
            for i0 in range(0, m, 512):
                i1 = min(m, i0 + 512)
                yv[i0:i1] = kernel_block(i0, i1) @ xv

    def conj_matmat(x, ncols, z, stream):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            xv, zv = torch.as_tensor(Raw(x), device="cuda"), torch.as_tensor(Raw(z), device="cuda")
            acc = torch.zeros((n, ncols), dtype=torch.complex128, device="cuda")
            for i0 in range(0, m, 512):
                i1 = min(m, i0 + 512)
                acc += kernel_block(i0, i1).conj().T @ xv[i0:i1]
            zv.copy_(acc)

    op = api.Operator((m, n), np.complex128, matmat, conj_matmat)
    omega = random_gaussian((n, k + p), np.complex128, seed=42)
    q = api.sample_range_by_rank(op, k, p, omega=omega)
    qr = api.QR.compute_from_range_estimate(q, op).compress(api.RANK(k))
    cid = qr.column_id()
    # the checker may form the matrix (the product path never did)
    d = np.sqrt(((xs[:, None, :] - ys[None, :, :]) ** 2).sum(axis=2))
    a = np.exp(1j * kappa * d) / d
    q_ref = ref.sample_range_by_rank(a, k, p, ref.OmegaStream(np.complex128, blocks=[omega]))
    cid_ref = ref.QR.compute_from_range_estimate(q_ref, a).compress(ref.RANK(k)).column_id()
    r, r_ref = ref.range_residual(a, q), ref.range_residual(a, q_ref)
    assert abs(r - r_ref) <= 1e-8 * r_ref, (r, r_ref)
    assert np.array_equal(cid.col_ind[:k], cid_ref.col_ind[:k])
    e, e_ref = ref.rel_diff_fro(cid.to_mat(), a), ref.rel_diff_fro(cid_ref.to_mat(), a)
    assert abs(e - e_ref) <= 1e-8 * e_ref, (e, e_ref)


def test_operator_errors(api):
    def boom(x, ncols, y, stream):
        raise RuntimeError("user callback failed")

    op = api.Operator((64, 32), np.float32, boom)
    with pytest.raises(api.LinalgError):
        api.sample_range_by_rank(op, 4, 2, seed=1)
    assert isinstance(op.last_exception, RuntimeError)
    with pytest.raises(AssertionError):     # RC_INVALID_ARGUMENT: an operator has no entries
        op.to_numpy()
    with pytest.raises(AssertionError):     # dense-only entry point
        api.QR.compute_from(op)
