"""The reference's 89 in-file unit tests, restated once, backend-agnostic.

``backend`` is any module exposing the reference's surface with numpy in/out:
``oracle.reference_path`` (CPU oracle) or ``rusty_compression_b200.api`` (the CUDA
path through the C ABI).  Shapes and thresholds are the reference's own:
src/pivoted_qr.rs:198-316, src/qr.rs:418-615, src/svd.rs:196-320,
src/col_interp_decomp.rs:160-241, src/row_interp_decomp.rs:160-235.
Inputs come from the oracle's seeded restatement of
``random_approximate_low_rank_matrix`` (src/random_matrix.rs:70-93), because the
reference draws from ``thread_rng()`` and pins nothing.
"""
import numpy as np

from oracle import reference_path as ref

SCALARS = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}
SHAPES = {"thin": (100, 50), "thick": (50, 100)}
CASES = [(s, d) for s in SCALARS for d in SHAPES]


_SEED = [0]


def low_rank(scalar, dim, sigma_min):
    return ref.random_approximate_low_rank_matrix(SHAPES[dim], 1.0, sigma_min, SCALARS[scalar],
                                                  seed=_SEED[0] + 1000 * CASES.index((scalar, dim)))


_SEED_CACHE = {}


def run_check(check, backend, scalar, dim):
    """The reference's entrywise-relative assertions are flaky by construction (they
    divide by matrix entries that can be ~0; the crate itself loosened one case,
    src/row_interp_decomp.rs:231).  So every case uses the first seed for which the
    ORACLE satisfies the reference's assertion; the backend under test then has to
    satisfy it on exactly that input."""
    key = (check.__name__, scalar, dim)
    if key not in _SEED_CACHE:
        for seed in range(50):
            _SEED[0] = seed
            try:
                check(ref, scalar, dim)
            except AssertionError:
                continue
            _SEED_CACHE[key] = seed
            break
        else:
            raise RuntimeError(f"oracle satisfies {key} for no seed")
    _SEED[0] = _SEED_CACHE[key]
    if backend is not ref:
        check(backend, scalar, dim)
    return _SEED_CACHE[key]


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


# --- src/pivoted_qr.rs:198-246
def check_pivoted_qr(backend, scalar, dim):
    mat = low_rank(scalar, dim, 1e-5)
    q, r, ind = backend.pivoted_qr(mat)
    k = min(mat.shape)
    assert q.shape == (mat.shape[0], k) and r.shape == (k, mat.shape[1]) and len(ind) == mat.shape[1]
    assert sorted(int(i) for i in ind) == list(range(mat.shape[1]))
    qtq = np.conj(q.T).dot(q)
    assert np.max(np.abs(qtq - np.eye(k))) < 1e-6
    assert np.all(np.tril(r, -1) == 0)
    prod = q.dot(r)
    for j in range(mat.shape[1]):
        assert _rel(prod[:, j], mat[:, ind[j]]) < 1e-6
    d = np.abs(np.diagonal(r))
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-5))


# --- src/pivoted_qr.rs:248-294
def check_pivoted_lq(backend, scalar, dim):
    mat = low_rank(scalar, dim, 1e-5)
    l, q, ind = backend.pivoted_lq(mat)
    k = min(mat.shape)
    assert l.shape == (mat.shape[0], k) and q.shape == (k, mat.shape[1]) and len(ind) == mat.shape[0]
    qqt = q.dot(np.conj(q.T))
    assert np.max(np.abs(qqt - np.eye(k))) < 1e-6
    prod = l.dot(q)
    for i in range(mat.shape[0]):
        assert _rel(prod[i, :], mat[ind[i], :]) < 1e-6


# --- src/qr.rs:427-457
def check_qr_compress_rank(backend, scalar, dim, rank=30):
    mat = low_rank(scalar, dim, 1e-10)
    qr = backend.QR.compute_from(mat).compress(backend.RANK(rank))
    assert qr.q.shape[1] == rank and qr.r.shape[0] == rank
    assert qr.q.shape[0] == mat.shape[0] and qr.r.shape[1] == mat.shape[1]
    assert _rel(qr.to_mat(), mat) < 1e-4


# --- src/qr.rs:459-489
def check_qr_compress_tol(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    qr = backend.QR.compute_from(mat).compress(backend.ADAPTIVE(tol))
    assert qr.q.shape[1] < min(mat.shape)
    assert _rel(qr.to_mat(), mat) < 5 * tol


# --- src/qr.rs:491-531
def check_col_id(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    qr = backend.QR.compute_from(mat).compress(backend.ADAPTIVE(tol))
    rank = qr.rank()
    cid = qr.column_id()
    assert cid.c.shape == (mat.shape[0], rank) and cid.z.shape == (rank, mat.shape[1])
    assert len(cid.col_ind) == mat.shape[1]
    assert _rel(cid.to_mat(), mat) < 5 * tol
    for i in range(rank):
        assert _rel(cid.c[:, i], mat[:, cid.col_ind[i]]) < tol


# --- src/qr.rs:532-571
def check_row_id(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    lq = backend.LQ.compute_from(mat).compress(backend.ADAPTIVE(tol))
    rank = lq.rank()
    rid = lq.row_id()
    assert rid.x.shape == (mat.shape[0], rank) and rid.r.shape == (rank, mat.shape[1])
    assert _rel(rid.to_mat(), mat) < 5 * tol
    for i in range(rank):
        assert _rel(rid.r[i, :], mat[rid.row_ind[i], :]) < tol


# --- src/svd.rs:203-227 (tolerance table :289-298)
def check_svd_to_qr(backend, scalar, dim):
    tol = 1e-5 if scalar in ("f32", "c32") else 1e-12
    mat = low_rank(scalar, dim, 1e-10)
    svd = backend.SVD.compute_from(mat)
    assert _rel(svd.to_mat(), mat) < tol
    qr = svd.to_qr()
    assert _rel(qr.to_mat(), mat) < tol


# --- src/svd.rs:229-259
def check_svd_compress_rank(backend, scalar, dim, rank=20):
    mat = low_rank(scalar, dim, 1e-10)
    svd = backend.SVD.compute_from(mat).compress(backend.RANK(rank))
    assert svd.u.shape == (mat.shape[0], rank) and svd.vt.shape == (rank, mat.shape[1])
    assert len(svd.s) == rank
    assert _rel(svd.to_mat(), mat) < 1e-4


# --- src/svd.rs:261-287
def check_svd_compress_tol(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    svd = backend.SVD.compute_from(mat).compress(backend.ADAPTIVE(tol))
    assert svd.rank() < min(mat.shape)
    assert _rel(svd.to_mat(), mat) < tol


# --- src/col_interp_decomp.rs:176-241
def check_two_sided_from_col(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    qr = backend.QR.compute_from(mat).compress(backend.ADAPTIVE(tol))
    rank = qr.rank()
    ts = qr.column_id().two_sided_id()
    assert ts.x.shape == (rank, rank)
    assert ts.c.shape == (mat.shape[0], rank) and ts.r.shape == (rank, mat.shape[1])
    assert _rel(ts.to_mat(), mat) < 5 * tol
    for i in range(rank):
        for j in range(rank):
            want = mat[ts.row_ind[i], ts.col_ind[j]]
            assert abs(ts.x[i, j] - want) / abs(want) < 10 * tol


# --- src/row_interp_decomp.rs:176-235 (c32/thick loosened to 5e-4 at :231)
def check_two_sided_from_row(backend, scalar, dim, tol=1e-4):
    mat = low_rank(scalar, dim, 1e-10)
    lq = backend.LQ.compute_from(mat).compress(backend.ADAPTIVE(tol))
    rank = lq.rank()
    ts = lq.row_id().two_sided_id()
    assert ts.x.shape == (rank, rank)
    assert _rel(ts.to_mat(), mat) < 5 * tol
    for i in range(rank):
        for j in range(rank):
            want = mat[ts.row_ind[i], ts.col_ind[j]]
            assert abs(ts.x[i, j] - want) / abs(want) < 10 * tol


ALL_CHECKS = [check_pivoted_qr, check_pivoted_lq, check_qr_compress_rank, check_qr_compress_tol,
              check_col_id, check_row_id, check_svd_to_qr, check_svd_compress_rank,
              check_svd_compress_tol, check_two_sided_from_col, check_two_sided_from_row]
