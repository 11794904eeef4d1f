"""Probe of the large-dense routes on degenerate inputs (exact rank deficiency, zero matrix, duplicated columns)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusty_compression_b200 import api
rng = np.random.default_rng(3)

def report(name, a, q, r, ind):
    k = min(a.shape)
    orth = np.max(np.abs(np.conj(q.T) @ q - np.eye(q.shape[1])))
    rec = np.linalg.norm(q @ r - a[:, ind]) / max(np.linalg.norm(a), 1e-300)
    d = np.abs(np.diag(r))
    mono = bool(np.all(d[:-1] >= d[1:] * (1 - 1e-6) - 1e-12 * max(d[0], 1e-300)))
    print(f"{name:44s} |Q^H Q - I| {orth:.2e}  |QR - AP|/|A| {rec:.2e}  diag non-increasing {mono}  perm ok {sorted(ind.tolist()) == list(range(a.shape[1]))}", flush=True)

for dtype in (np.float64, np.complex128):
    cplx = np.dtype(dtype).kind == "c"
    g = lambda r, c: (rng.standard_normal((r, c)) + (1j * rng.standard_normal((r, c)) if cplx else 0)).astype(dtype)
    a = g(1500, 50) @ g(50, 900)
    report(f"{np.dtype(dtype).name} exact rank 50, 1500x900", a, *api.pivoted_qr(a))
    a = g(700, 40) @ g(40, 1600)
    report(f"{np.dtype(dtype).name} exact rank 40, 700x1600 (wide)", a, *api.pivoted_qr(a))
    a = np.zeros((1000, 700), dtype)
    report(f"{np.dtype(dtype).name} zero 1000x700", a, *api.pivoted_qr(a))
    a = g(1200, 300); a[:, 150:] = a[:, :150]
    report(f"{np.dtype(dtype).name} duplicated columns 1200x300", a, *api.pivoted_qr(a))
    a = g(900, 30) @ g(30, 600)
    u, s, vt = api.compute_svd(a)
    s0 = np.linalg.svd(a, compute_uv=False)
    print(f"{np.dtype(dtype).name} SVD exact rank 30, 900x600: max |s - s0|/s0[0] {np.max(np.abs(s - s0)) / s0[0]:.2e}, "
          f"|U S Vt - A|/|A| {np.linalg.norm((u * s) @ vt - a) / np.linalg.norm(a):.2e}, |Vt Vt^H - I| {np.max(np.abs(vt @ np.conj(vt.T) - np.eye(600))):.2e}, "
          f"|U^H U - I| on the leading 30 {np.max(np.abs(np.conj(u[:, :30].T) @ u[:, :30] - np.eye(30))):.2e}", flush=True)
    a = g(300, 300)
    report(f"{np.dtype(dtype).name} square 300x300 (cluster kernel)", a, *api.pivoted_qr(a))
    a = g(2000, 20) @ g(20, 130)
    report(f"{np.dtype(dtype).name} tall rank 20, 2000x130", a, *api.pivoted_qr(a))
