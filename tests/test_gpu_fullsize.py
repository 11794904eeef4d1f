"""BASELINE.json configs at full size.

configs[1] (the bench workload): rank-64 (+10) rSVD of a 65536 x 8192 f64 matrix with 2 power
iterations -- full parity against the oracle on the same A and the same Omega (the CPU oracle needs
~10 s at this size), plus the size-independent properties (orthonormality, U diag(s) Vt reconstruction,
agreement with the known spectrum of the synthetic matrix)."""
import numpy as np
import pytest

from oracle import reference_path as ref
from oracle.philox import random_gaussian

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_config2_full_size_parity_and_properties():
    from rusty_compression_b200 import api
    m, n, k, p, it = 65536, 8192, 64, 10, 2
    ctx = api.default_context()
    a_dev = api.decaying_spectrum_matrix((m, n), np.float64, 1234, r0=512, decade_every=16.0)
    omega = random_gaussian((n, k + p), np.float64, seed=42)
    ctx.reset_counters()
    q_dev = api.sample_range_power_iteration(a_dev, k, p, it, omega=omega, device=True)
    svd_dev = api.SVD.compute_from_range_estimate(q_dev, a_dev)
    assert ctx.counter("kernel_launches") > 50
    s_dev, u, vt = svd_dev.s_f64(), svd_dev.u, svd_dev.vt
    q = q_dev.to_numpy()
    # --- size-independent properties
    assert q.shape == (m, k) and u.shape == (m, k) and vt.shape == (k, n)
    assert np.max(np.abs(q.T.dot(q) - np.eye(k))) < 1e-12
    assert np.max(np.abs(u.T.dot(u) - np.eye(k))) < 1e-12
    assert np.max(np.abs(vt.dot(vt.T) - np.eye(k))) < 1e-12
    assert np.all(np.diff(s_dev) <= 0)
    sigma = 10.0 ** (-np.arange(k) / 16.0)            # the generator's spectrum
    assert np.all(s_dev <= sigma * (1 + 1e-10))       # Ritz values never exceed the true singular values
    assert np.max(np.abs(s_dev[:32] - sigma[:32]) / sigma[:32]) < 1e-4 and np.max(np.abs(s_dev - sigma) / sigma) < 0.3
    # same Omega generated on device from the Philox seed gives the same singular values
    q_seed = api.sample_range_power_iteration(a_dev, k, p, it, seed=42, device=True)
    s_seed = api.SVD.compute_from_range_estimate(q_seed, a_dev).s_f64()
    assert np.max(np.abs(s_seed - s_dev) / s_dev) < 1e-9
    # --- parity with the oracle (LAPACK path) on the same A and the same Omega
    a = a_dev.to_numpy()
    q_ref = ref.sample_range_power_iteration(a, k, p, it, ref.OmegaStream(np.float64, blocks=[omega]))
    svd_ref = ref.SVD.compute_from_range_estimate(q_ref, a)
    err_s = np.max(np.abs(s_dev - svd_ref.s) / svd_ref.s)
    res_dev, res_ref = ref.range_residual(a, q), ref.range_residual(a, q_ref)
    print(f"config 2 full size: max rel singular-value error {err_s:.2e}; residual {res_dev:.12e} vs {res_ref:.12e}")
    assert err_s < 1e-10
    assert abs(res_dev - res_ref) <= 1e-10 * res_ref
