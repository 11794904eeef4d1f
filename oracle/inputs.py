"""Synthetic inputs for the BASELINE.json configs, at any size.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Deterministic in
``seed``; the same host buffer is handed to the oracle and uploaded to the GPU,
so both sides see bit-identical A.
"""
import numpy as np


def _orthonormal(rows, cols, rng, dtype):
    g = rng.standard_normal((rows, cols))
    if np.dtype(dtype).kind == "c":
        g = g + 1j * rng.standard_normal((rows, cols))
    q, _ = np.linalg.qr(g)
    return q


def decaying_spectrum_matrix(m, n, dtype, seed, r0=512, decade_every=16.0):
    """Configs 2 and 3: A = U diag(sigma) V^H, U (m x r0), V (n x r0) orthonormalised
    seeded Gaussians, sigma_j = 10^(-j/decade_every).  Built in double precision,
    then cast.  Returns (A, sigma)."""
    dtype = np.dtype(dtype)
    r0 = min(r0, m, n)
    rng = np.random.default_rng(seed)
    u = _orthonormal(m, r0, rng, dtype)
    v = _orthonormal(n, r0, rng, dtype)
    sigma = 10.0 ** (-np.arange(r0) / decade_every)
    a = (u * sigma[None, :]).dot(np.conj(v.T))
    return np.ascontiguousarray(a.astype(dtype)), sigma


def helmholtz_kernel_matrix(m, n, dtype, seed=7, kappa=20.0, shift=1.5):
    """Config 5: A_ij = exp(i kappa |x_i - y_j|) / |x_i - y_j| for two well-separated
    unit boxes (gap 0.5).  Real dtypes take the real part."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    x = rng.random((m, 3))
    y = rng.random((n, 3))
    y[:, 0] += shift
    a = np.empty((m, n), dtype=np.complex128 if dtype.kind == "c" else np.float64)
    step = max(1, (1 << 22) // max(n, 1))
    for s in range(0, m, step):
        t = min(m, s + step)
        d = np.sqrt(((x[s:t, None, :] - y[None, :, :]) ** 2).sum(axis=2))
        k = np.exp(1j * kappa * d) / d
        a[s:t] = k if dtype.kind == "c" else k.real
    return np.ascontiguousarray(a.astype(dtype))


def helmholtz_kernel_matrix_philox(m, n, dtype, seed=7, kappa=20.0, shift=1.5, row_offset=0):
    """Host mirror of the library's device generator rc_helmholtz_kernel_matrix: same kernel as
    ``helmholtz_kernel_matrix`` but the points come from the Philox stream (coordinate d of point i = the
    [0, 1) uniform of counter 3 i + d; stream 301 for x, 302 for y), so row shards regenerate their rows."""
    from .philox import _uniform_pair
    dtype = np.dtype(dtype)
    ex = np.arange(3 * row_offset, 3 * (row_offset + m), dtype=np.uint64)
    ey = np.arange(0, 3 * n, dtype=np.uint64)
    x = _uniform_pair(ex, seed, 301)[1].reshape(m, 3)
    y = _uniform_pair(ey, seed, 302)[1].reshape(n, 3)
    y[:, 0] += shift
    d = np.sqrt(((x[:, None, :] - y[None, :, :]) ** 2).sum(axis=2))
    k = np.exp(1j * kappa * d) / d
    return np.ascontiguousarray((k if dtype.kind == "c" else k.real).astype(dtype))


def tall_shard_matrix(row0, rows, n, dtype, seed, m_total, r0=512, decade_every=64.0):
    """Config 4: rows [row0, row0+rows) of A = m^(-1/2) G diag(sigma) V^T with
    G_ij ~ N(0,1) from Philox keyed by (global row, j) and V (n x r0) a shared
    orthonormal factor.  Any shard of any rank can regenerate its rows."""
    from .philox import random_gaussian
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    v = _orthonormal(n, r0, rng, np.float64)
    sigma = 10.0 ** (-np.arange(r0) / decade_every)
    g = random_gaussian((rows, r0), np.float64, seed, stream=7, row_offset=row0)
    a = (g * (sigma[None, :] / np.sqrt(float(m_total)))).dot(v.T)
    return np.ascontiguousarray(a.astype(dtype))
