#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run from the repo root):

    python tests/golden/make_golden.py

Why these exist.  The reference (a Rust crate) cannot be compiled in this image and holds no
golden vectors of its own for the numerical routines (every test draws from `thread_rng()`); the
only exact known answers it has are the permutation tests (src/permutation.rs:192-239, file
`permutation_known_answers.json`, values transcribed from the reference's asserts).  Every other
fixture here is an output of the ORACLE (oracle/reference_path.py: the crate's control flow restated
over the same LAPACK routines, scipy/OpenBLAS) on small seeded inputs that are stored alongside, so

  * tests/test_golden_oracle.py (CPU) detects drift of the oracle itself (scipy / OpenBLAS upgrades,
    edits to the restatement), and
  * tests/test_gpu_golden.py (B200, through the C ABI) checks the CUDA path against frozen numbers
    that do not depend on the oracle running on the GPU box.

Each `case_<scalar>.npz` holds: A (96 x 80, the config-5 kernel matrix at toy size), Omega blocks,
and for every pipeline of SURVEY.md 8(a) the index vectors (exact) and the scalar quantities the
parity contract names (singular values, range residual, ID errors, adaptive history).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_path as ref                     # noqa: E402
from oracle.inputs import decaying_spectrum_matrix, helmholtz_kernel_matrix   # noqa: E402
from oracle.philox import random_gaussian                    # noqa: E402

M, N, K, P, IT = 96, 80, 16, 4, 2
NAMES = {np.float32: "f32", np.float64: "f64", np.complex64: "c32", np.complex128: "c64"}


def build_case(dtype):
    a = helmholtz_kernel_matrix(M, N, dtype, seed=21)
    omega = random_gaussian((N, K + P), dtype, seed=42)
    out = {"a": a, "omega": omega}
    # --- full pivoted QR / LQ and thin SVD of A (src/pivoted_qr.rs, src/compute_svd.rs)
    qr = ref.QR.compute_from(a)
    out["pqr_ind"] = np.asarray(qr.ind, dtype=np.int64)
    out["pqr_absdiag"] = np.abs(np.diag(qr.r)).astype(np.float64)
    lq = ref.LQ.compute_from(a)
    out["plq_ind"] = np.asarray(lq.ind, dtype=np.int64)
    out["svd_s"] = ref.SVD.compute_from(a).s.astype(np.float64)
    # --- fixed-rank samplers (src/random_sampling.rs:103-160) and the factorizations built on them
    q0 = ref.sample_range_by_rank(a, K, P, ref.OmegaStream(dtype, blocks=[omega]))
    out["by_rank_residual"] = np.float64(ref.range_residual(a, q0))
    y = ref.DenseOperator(a).matmat(omega)
    out["sketch_ind"] = np.asarray(ref.QR.compute_from(y).ind, dtype=np.int64)
    qp = ref.sample_range_power_iteration(a, K, P, IT, ref.OmegaStream(dtype, blocks=[omega]))
    out["power_residual"] = np.float64(ref.range_residual(a, qp))
    out["rsvd_s"] = ref.SVD.compute_from_range_estimate(qp, a).s.astype(np.float64)
    # --- column / two-sided ID from the sketch (src/qr.rs:270-323, src/col_interp_decomp.rs:116-125)
    qrr = ref.QR.compute_from_range_estimate(q0, a)
    out["range_qr_ind"] = np.asarray(qrr.ind, dtype=np.int64)
    cid = qrr.compress(ref.RANK(K)).column_id()
    out["cid_error"] = np.float64(ref.rel_diff_fro(cid.to_mat(), a))
    ts = cid.two_sided_id()
    out["ts_row_ind"] = np.asarray(ts.row_ind, dtype=np.int64)
    out["ts_error"] = np.float64(ref.rel_diff_fro(ts.to_mat(), a))
    # --- row ID route (src/qr.rs:363-403, src/row_interp_decomp.rs:120-130)
    rid = ref.LQ.compute_from(a).compress(ref.RANK(K)).row_id()
    out["rid_error"] = np.float64(ref.rel_diff_fro(rid.to_mat(), a))
    ts2 = rid.two_sided_id()
    out["ts2_col_ind"] = np.asarray(ts2.col_ind, dtype=np.int64)
    out["ts2_error"] = np.float64(ref.rel_diff_fro(ts2.to_mat(), a))
    # --- next to every ?geqp3 sequence in the working precision (what the reference itself produces), the
    #     sequence ?geqp3 picks in DOUBLE precision on the same input (`<key>_f64`: identical for f64 / c64) and the
    #     signed per-step gaps of that sequence (`<key>_gaps`).  The CUDA path is held to the `_f64` sequence wherever
    #     the gap exceeds 1e-6.
    b = ref.conj_t(ref.DenseOperator(a).conj_matmat(q0))
    for key, mat, upto in (("pqr_ind", a, None), ("plq_ind", ref.conj_t(a), None), ("sketch_ind", y, None),
                           ("range_qr_ind", b, K), ("ts_row_ind", ref.conj_t(cid.c), K), ("ts2_col_ind", rid.r, K)):
        seq = ref.pivot_sequence_f64(mat)
        out[key + "_f64"] = np.asarray(seq, dtype=np.int64)
        out[key + "_gaps"] = ref.pivot_gaps(mat, seq, upto=upto)
    # --- tolerance compression (src/qr.rs:187-200, src/svd.rs:87-101)
    tol = 1e-3
    out["qr_tol_rank"] = np.int64(qr.compress(ref.ADAPTIVE(tol)).rank())
    out["svd_tol_rank"] = np.int64(ref.SVD.compute_from(a).compress(ref.ADAPTIVE(tol)).rank())
    # --- adaptive sampler (src/random_sampling.rs:223-274), blocks of 4 columns
    stream = ref.OmegaStream(dtype, seed=11)
    qa, hist = ref.sample_range_adaptive(a, 1e-3, 4, stream)
    out["adaptive_blocks"] = np.concatenate(stream.drawn, axis=1)
    out["adaptive_ranks"] = np.asarray([r for r, _ in hist], dtype=np.int64)
    out["adaptive_res"] = np.asarray([e for _, e in hist], dtype=np.float64)
    out["adaptive_residual"] = np.float64(ref.range_residual(a, qa))
    return out


def main():
    for dtype, name in NAMES.items():
        case = build_case(dtype)
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **case)
        print(name, {k: (v.shape if getattr(v, "ndim", 0) else float(v)) for k, v in case.items() if k not in ("a", "omega", "adaptive_blocks")})
    # config 2 at toy size: decaying spectrum, f64 (the bench workload's construction)
    a, sig = decaying_spectrum_matrix(512, 128, np.float64, seed=1234, r0=64, decade_every=4.0)
    omega = random_gaussian((128, 20), np.float64, seed=42)
    q = ref.sample_range_power_iteration(a, 16, 4, 2, ref.OmegaStream(np.float64, blocks=[omega]))
    svd = ref.SVD.compute_from_range_estimate(q, a)
    np.savez_compressed(os.path.join(HERE, "config2_toy_f64.npz"), a=a, omega=omega, sigma=sig,
                        rsvd_s=svd.s.astype(np.float64), residual=np.float64(ref.range_residual(a, q)))
    # the reference's own exact known answers (src/permutation.rs:192-239)
    perm = {
        "source": "src/permutation.rs:192-239",
        "perm": [2, 0, 1],
        "matrix": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]],
        "COL": [[3.0, 1.0, 2.0], [6.0, 4.0, 5.0], [9.0, 7.0, 8.0]],
        "COLINV": [[2.0, 3.0, 1.0], [5.0, 6.0, 4.0], [8.0, 9.0, 7.0]],
        "ROW": [[7.0, 8.0, 9.0], [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]],
        "ROWINV": [[4.0, 5.0, 6.0], [7.0, 8.0, 9.0], [1.0, 2.0, 3.0]],
        "vector": [1.0, 2.0, 3.0],
        "NOINV": [3.0, 1.0, 2.0],
        "INV": [2.0, 3.0, 1.0],
    }
    with open(os.path.join(HERE, "permutation_known_answers.json"), "w") as f:
        json.dump(perm, f, indent=1)


if __name__ == "__main__":
    main()
