"""Shared driver for the golden-fixture tests: runs every pipeline of tests/golden/make_golden.py on
a given implementation (the oracle, or the CUDA path through the C ABI -- both expose the reference's
names) and compares with the frozen numbers.

Index vectors.  The fixtures hold, for every pivoted factorisation, the ?geqp3 sequence in
the working precision (`<key>`: what the reference itself produces), the sequence ?geqp3 picks in DOUBLE
precision on the same input (`<key>_f64`; identical for f64 / c64) and the signed per-step gaps of the
latter (`<key>_gaps`).  The oracle must reproduce `<key>` exactly.  The CUDA path takes its pivot decisions
in double for every scalar type, so it is held to the double-precision sequence: a sequence that differs from
it is replayed step by step in double precision and every step must have picked a column whose trailing norm
is the maximum or within 1e-6 (relative) of it -- for all four scalars, no single-precision allowance -- else
the test FAILS naming the first divergent step and its gap.  Nothing is skipped after a tie: when the device's
(validated) order differs from the frozen working-precision sequence, the quantities that follow are compared
with the oracle replayed in the device's order (`ref.pivoted_qr_with_order`) instead of the frozen numbers."""
import json
import os

import numpy as np

from oracle import reference_path as ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["f32", "f64", "c32", "c64"]
K, P, IT = 16, 4, 2
PIVOT_TIE = 1e-6          # north_star: indices bit-exact wherever the pivot norm gap exceeds 1e-6 relative


def load(name):
    return dict(np.load(os.path.join(GOLDEN, f"case_{name}.npz")))


def permutation_answers():
    with open(os.path.join(GOLDEN, "permutation_known_answers.json")) as f:
        return json.load(f)


def rtol(name):
    return 1e-10 if name in ("f64", "c64") else 1e-4


def adjudicate(matrix, got, want, upto=None, label="", oracle_side=False):
    """`matrix`: the input the implementation factored; `got`: its pivot vector; `want`: the ?geqp3 sequence in the
    working precision (frozen or live oracle).  Returns None when got == want on the first `upto` steps (frozen /
    oracle numbers downstream apply as they are), else -- after the complete double-precision validation of
    oracle.reference_path.check_pivot_sequence, which raises on a contract violation -- the order to replay the
    oracle with.  Never skips."""
    got, want = np.asarray(got), np.asarray(want)
    k = min(matrix.shape) if upto is None else min(min(matrix.shape), int(upto))
    if np.array_equal(got[:k], want[:k]):
        return None
    assert not oracle_side, f"{label}: the oracle no longer reproduces its frozen pivot sequence"
    report = ref.check_pivot_sequence(matrix, got, upto=k, tie=PIVOT_TIE)
    j = int(np.nonzero(got[:k] != want[:k])[0][0])
    print(f"[pivots] {label}: differs from the working-precision ?geqp3 sequence at step {j}; against the "
          f"double-precision sequence: identical={report['identical']}, first divergence {report['first_divergence']}, "
          f"{len(report['ties'])} tie(s) within {PIVOT_TIE:g}; continuing with the oracle replayed in this order")
    return got


def close(got, want, tol, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert err <= tol, f"{what}: relative deviation {err:.3e} > {tol:.1e}"


def run_case(impl, name, g, make_stream, adaptive, to_np=np.asarray, conj_matmat=None, oracle_side=False):
    """impl: module exposing QR/LQ/SVD/RANK/ADAPTIVE/sample_range_* with the reference's semantics;
    conj_matmat(a, q): the implementation's own A^H q (the factor b the implementation pivots is built from it)."""
    a, omega = g["a"], g["omega"]
    tol = rtol(name)
    if conj_matmat is None:
        conj_matmat = lambda mat, q: ref.DenseOperator(mat).conj_matmat(q)       # noqa: E731
    floor = 1e-3 if tol > 1e-6 else 1e-6          # entries above the roundoff floor of the working precision
    # ---- full pivoted QR / LQ, thin SVD
    qr = impl.QR.compute_from(a)
    order = adjudicate(a, qr.ind, g["pqr_ind"], label=f"{name} QR::compute_from", oracle_side=oracle_side)
    d = np.abs(np.diag(to_np(qr.r))).astype(np.float64)
    if order is None:
        dg, want_rank = g["pqr_absdiag"], int(g["qr_tol_rank"])
    else:
        qo = ref.QR.compute_from(a, order=order)
        dg, want_rank = np.abs(np.diag(qo.r)).astype(np.float64), int(qo.compress(ref.ADAPTIVE(1e-3)).rank())
    sig = dg > dg[0] * floor
    close(d[sig], dg[sig], 50 * tol, f"{name} |diag R|")
    assert int(qr.compress(impl.ADAPTIVE(1e-3)).rank()) == want_rank
    lq = impl.LQ.compute_from(a)
    lq_order = adjudicate(ref.conj_t(a), lq.ind, g["plq_ind"], label=f"{name} LQ::compute_from", oracle_side=oracle_side)
    svd = impl.SVD.compute_from(a)
    s = np.asarray(svd.s, dtype=np.float64)
    big = g["svd_s"] > g["svd_s"][0] * floor          # relative accuracy of s_j is eps * s_0 / s_j
    close(s[big], g["svd_s"][big], 10 * tol, f"{name} singular values")
    assert int(svd.compress(impl.ADAPTIVE(1e-3)).rank()) == int(g["svd_tol_rank"])

    # ---- fixed-rank samplers and the factorisations built on them
    q0 = to_np(impl.sample_range_by_rank(a, K, P, **make_stream([omega])))
    close(ref.range_residual(a, q0), g["by_rank_residual"], tol, f"{name} by-rank residual")
    qp = to_np(impl.sample_range_power_iteration(a, K, P, IT, **make_stream([omega])))
    close(ref.range_residual(a, qp), g["power_residual"], tol, f"{name} power-iteration residual")
    close(np.asarray(impl.SVD.compute_from_range_estimate(qp, a).s, dtype=np.float64), g["rsvd_s"], tol, f"{name} rSVD s")

    qrr = impl.QR.compute_from_range_estimate(q0, a)
    b = ref.conj_t(to_np(conj_matmat(a, q0)))
    order = adjudicate(b, qrr.ind, g["range_qr_ind"], upto=K, label=f"{name} QR::compute_from_range_estimate",
                       oracle_side=oracle_side)
    cid = qrr.compress(impl.RANK(K)).column_id()
    if order is None:
        want_cid = g["cid_error"]
    else:
        cid_o = ref.QR.compute_from_range_estimate(q0, a, order=order).compress(ref.RANK(K)).column_id()
        want_cid = ref.rel_diff_fro(cid_o.to_mat(), a)
    close(ref.rel_diff_fro(to_np(cid.to_mat()), a), want_cid, tol, f"{name} column-ID error")
    ts = cid.two_sided_id()
    c_dev = to_np(cid.c)
    order2 = adjudicate(ref.conj_t(c_dev), ts.row_ind, g["ts_row_ind"], upto=K, label=f"{name} ColumnID::two_sided_id",
                        oracle_side=oracle_side)
    if order is None and order2 is None:
        want_ts = g["ts_error"]
    else:
        ts_o = ref.ColumnID(c_dev, to_np(cid.z), np.asarray(cid.col_ind)).two_sided_id(order=order2)
        want_ts = ref.rel_diff_fro(ts_o.to_mat(), a)
    close(ref.rel_diff_fro(to_np(ts.to_mat()), a), want_ts, 10 * tol, f"{name} two-sided ID error")

    # ---- row ID route
    rid = lq.compress(impl.RANK(K)).row_id()
    if lq_order is None:
        want_rid = g["rid_error"]
    else:
        want_rid = ref.rel_diff_fro(ref.LQ.compute_from(a, order=lq_order).compress(ref.RANK(K)).row_id().to_mat(), a)
    close(ref.rel_diff_fro(to_np(rid.to_mat()), a), want_rid, tol, f"{name} row-ID error")
    ts2 = rid.two_sided_id()
    r_dev = to_np(rid.r)
    order3 = adjudicate(r_dev, ts2.col_ind, g["ts2_col_ind"], upto=K, label=f"{name} RowID::two_sided_id",
                        oracle_side=oracle_side)
    if lq_order is None and order3 is None:
        want_ts2 = g["ts2_error"]
    else:
        ts2_o = ref.RowID(to_np(rid.x), r_dev, np.asarray(rid.row_ind)).two_sided_id(order=order3)
        want_ts2 = ref.rel_diff_fro(ts2_o.to_mat(), a)
    close(ref.rel_diff_fro(to_np(ts2.to_mat()), a), want_ts2, 10 * tol, f"{name} two-sided (row route) error")

    # ---- adaptive sampler
    qa, hist = adaptive(a, 1e-3, 4, g["adaptive_blocks"])
    assert [int(r) for r, _ in hist] == [int(r) for r in g["adaptive_ranks"]], f"{name}: adaptive rank history"
    close([e for _, e in hist], g["adaptive_res"], 100 * tol, f"{name} adaptive residual history")
    close(ref.range_residual(a, to_np(qa)), g["adaptive_residual"], 100 * tol, f"{name} adaptive range residual")
