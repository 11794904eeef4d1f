"""Shared driver for the golden-fixture tests: runs every pipeline of tests/golden/make_golden.py on
a given implementation (the oracle, or the CUDA path through the C ABI -- both expose the reference's
names) and compares with the frozen numbers.  Index vectors must match exactly; a mismatch is only
excused at a pivot step whose relative norm gap is below the north_star threshold (1e-6 for
f64/c64; single precision cannot resolve gaps below ~1e-3 of a downdated norm)."""
import json
import os

import numpy as np

from oracle import reference_path as ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["f32", "f64", "c32", "c64"]
K, P, IT = 16, 4, 2


def load(name):
    return dict(np.load(os.path.join(GOLDEN, f"case_{name}.npz")))


def permutation_answers():
    with open(os.path.join(GOLDEN, "permutation_known_answers.json")) as f:
        return json.load(f)


def rtol(name):
    return 1e-10 if name in ("f64", "c64") else 1e-4


def same_indices(matrix_for_gaps, got, want, name, upto=None):
    """True if equal; False if the first mismatch sits on a numerical tie (later quantities are then
    not comparable); asserts otherwise."""
    got, want = np.asarray(got)[:upto], np.asarray(want)[:upto]
    if np.array_equal(got, want):
        return True
    j = int(np.nonzero(got != want)[0][0])
    gaps = ref.pivot_gaps(matrix_for_gaps, np.asarray(want))
    lim = 1e-6 if name in ("f64", "c64") else 1e-3
    assert j < len(gaps) and gaps[j] <= lim, f"{name}: index mismatch at step {j}, gap {gaps[j]:.3e} > {lim}"
    return False


def close(got, want, tol, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert err <= tol, f"{what}: relative deviation {err:.3e} > {tol:.1e}"


def run_case(impl, name, g, make_stream, adaptive, to_np=np.asarray):
    """impl: module exposing QR/LQ/SVD/RANK/ADAPTIVE/sample_range_* with the reference's semantics."""
    a, omega = g["a"], g["omega"]
    tol = rtol(name)
    # the sqrt-of-eps sensitivity of a residual near its floor: residuals here are O(0.1), so tol holds
    qr = impl.QR.compute_from(a)
    if same_indices(a, qr.ind, g["pqr_ind"], name):
        d, dg = np.abs(np.diag(to_np(qr.r))).astype(np.float64), g["pqr_absdiag"]
        sig = dg > dg[0] * (1e-3 if tol > 1e-6 else 1e-6)          # entries above the roundoff floor
        close(d[sig], dg[sig], 50 * tol, f"{name} |diag R|")
        assert int(qr.compress(impl.ADAPTIVE(1e-3)).rank()) == int(g["qr_tol_rank"])
    same_indices(ref.conj_t(a), impl.LQ.compute_from(a).ind, g["plq_ind"], name)
    svd = impl.SVD.compute_from(a)
    s = np.asarray(svd.s, dtype=np.float64)
    big = g["svd_s"] > g["svd_s"][0] * (1e-3 if tol > 1e-6 else 1e-6)     # relative accuracy of s_j is eps * s_0 / s_j
    close(s[big], g["svd_s"][big], 10 * tol, f"{name} singular values")
    assert int(svd.compress(impl.ADAPTIVE(1e-3)).rank()) == int(g["svd_tol_rank"])

    q0 = to_np(impl.sample_range_by_rank(a, K, P, **make_stream([omega])))
    close(ref.range_residual(a, q0), g["by_rank_residual"], tol, f"{name} by-rank residual")
    qp = to_np(impl.sample_range_power_iteration(a, K, P, IT, **make_stream([omega])))
    close(ref.range_residual(a, qp), g["power_residual"], tol, f"{name} power-iteration residual")
    close(np.asarray(impl.SVD.compute_from_range_estimate(qp, a).s, dtype=np.float64), g["rsvd_s"], tol, f"{name} rSVD s")

    qrr = impl.QR.compute_from_range_estimate(q0, a)
    b = ref.conj_t(ref.DenseOperator(a).conj_matmat(q0))
    if same_indices(b, qrr.ind, g["range_qr_ind"], name, upto=K):
        cid = qrr.compress(impl.RANK(K)).column_id()
        close(ref.rel_diff_fro(to_np(cid.to_mat()), a), g["cid_error"], tol, f"{name} column-ID error")
        ts = cid.two_sided_id()
        if same_indices(ref.conj_t(to_np(cid.c)), ts.row_ind, g["ts_row_ind"], name, upto=K):
            close(ref.rel_diff_fro(to_np(ts.to_mat()), a), g["ts_error"], 10 * tol, f"{name} two-sided ID error")
    lq = impl.LQ.compute_from(a)
    if np.array_equal(np.asarray(lq.ind), g["plq_ind"]):
        rid = lq.compress(impl.RANK(K)).row_id()
        close(ref.rel_diff_fro(to_np(rid.to_mat()), a), g["rid_error"], tol, f"{name} row-ID error")
        ts2 = rid.two_sided_id()
        if same_indices(to_np(rid.r), ts2.col_ind, g["ts2_col_ind"], name, upto=K):
            close(ref.rel_diff_fro(to_np(ts2.to_mat()), a), g["ts2_error"], 10 * tol, f"{name} two-sided (row route) error")

    qa, hist = adaptive(a, 1e-3, 4, g["adaptive_blocks"])
    assert [int(r) for r, _ in hist] == [int(r) for r in g["adaptive_ranks"]], f"{name}: adaptive rank history"
    close([e for _, e in hist], g["adaptive_res"], 100 * tol, f"{name} adaptive residual history")
    close(ref.range_residual(a, to_np(qa)), g["adaptive_residual"], 100 * tol, f"{name} adaptive range residual")
