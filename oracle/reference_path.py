"""CPU restatement of rusty-compression's randomized low-rank hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA
path and the timed CPU baseline, never a fallback.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Arithmetic is delegated to the same LAPACK routines the
crate reaches through ndarray-linalg / lax / lapack: ``?geqp3``
(src/pivoted_qr.rs:139-173), ``?orgqr/?ungqr`` (src/pivoted_qr.rs:104-111),
``?gesdd`` jobz='S' (src/compute_svd.rs:19), ``?trtrs`` (src/qr.rs:298, 392),
and BLAS ``dot``.  Parity of these numerical routines is *unpinned* by the
reference (it has no golden vectors); see ``oracle/__init__.py``.

Quirks reproduced on purpose: Q1 power iteration restarts
from A*Omega every trip; Q3 tolerance compression errors when no diagonal entry
falls below tol; Q6 adaptive loop semantics; Q7 C = Q*R11 (not gathered from A);
Q8 index vectors always full length; Q10 two_sided_id on a ColumnID uses an
uncompressed pivoted LQ of C.
"""
from dataclasses import dataclass
import math

import numpy as np
from scipy.linalg import lapack as _lp


# --------------------------------------------------------------------------- errors
class RustyCompressionError(Exception):
    """src/types.rs:11-21."""


class LinalgError(RustyCompressionError):
    pass


class CompressionError(RustyCompressionError):
    pass


class LayoutError(RustyCompressionError):
    pass


class PivotedQRError(RustyCompressionError):
    pass


# ---------------------------------------------------------------- CompressionType
@dataclass(frozen=True)
class ADAPTIVE:
    """src/lib.rs:82-87 CompressionType::ADAPTIVE(f64)."""
    tol: float


@dataclass(frozen=True)
class RANK:
    """src/lib.rs:82-87 CompressionType::RANK(usize)."""
    rank: int


_PREFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d",
           np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}


def _lapack(name, dtype):
    p = _PREFIX[np.dtype(dtype)]
    if name == "orgqr" and p in "cz":
        name = "ungqr"
    return getattr(_lp, p + name)


def conj_t(a):
    """``.t().map(|x| x.conj())`` (e.g. src/qr.rs:315)."""
    return np.ascontiguousarray(np.conj(a.T))


# ------------------------------------------------------------------ permutations
def invert_permutation_vector(perm):
    """src/permutation.rs:28-38."""
    perm = np.asarray(perm)
    inv = np.zeros(len(perm), dtype=np.int64)
    inv[perm] = np.arange(len(perm), dtype=np.int64)
    return inv


def apply_permutation_matrix(mat, index_array, mode):
    """src/permutation.rs:84-144.  mode in {'COL','ROW','COLINV','ROWINV'}."""
    index_array = np.asarray(index_array)
    m, n = mat.shape
    if mode == "COL":
        assert len(index_array) == n, "Length of index array and number of columns differ."
        return np.ascontiguousarray(mat[:, index_array])
    if mode == "ROW":
        assert len(index_array) == m, "Length of index array and number of rows differ."
        return np.ascontiguousarray(mat[index_array, :])
    if mode == "COLINV":
        assert len(index_array) == n, "Length of index array and number of columns differ."
        return np.ascontiguousarray(mat[:, invert_permutation_vector(index_array)])
    if mode == "ROWINV":
        assert len(index_array) == m, "Length of index array and number of rows differ."
        return np.ascontiguousarray(mat[invert_permutation_vector(index_array), :])
    raise ValueError(mode)


def apply_permutation_vector(vec, index_array, mode):
    """src/permutation.rs:147-184.  mode in {'INV','NOINV'}."""
    index_array = np.asarray(index_array)
    assert len(index_array) == len(vec), \
        "The input vector and the index array must have the same length"
    if mode == "INV":
        return vec[invert_permutation_vector(index_array)].copy()
    if mode == "NOINV":
        return vec[index_array].copy()
    raise ValueError(mode)


# ----------------------------------------------------------------------- RelDiff
def rel_diff_fro(first, second):
    """src/types.rs:182-188: ||first - second||_F / ||second||_F."""
    return np.linalg.norm(first - second) / np.linalg.norm(second)


def rel_diff_l2(first, second):
    """src/types.rs:190-196."""
    return np.linalg.norm(first - second) / np.linalg.norm(second)


# ---------------------------------------------------------- operator plugin API
class DenseOperator:
    """MatVec/MatMat/ConjMatVec/ConjMatMat for a dense array
    (src/types.rs:40-133, 145-146).

    ``route='gemv'`` is the reference-faithful default impl: one GEMV per column
    (src/types.rs:60-70, 90-100, 119, 129-131).  ``route='gemm'`` is the same
    product as one GEMM ("best-case CPU"); results agree up to summation order.
    """

    def __init__(self, a, route="gemm"):
        self.a = a
        self.route = route

    def nrows(self):
        return self.a.shape[0]

    def ncols(self):
        return self.a.shape[1]

    def matvec(self, x):
        return self.a.dot(x)

    def conj_matvec(self, x):
        return np.conj(np.conj(x).dot(self.a))

    def matmat(self, x):
        if self.route == "gemm":
            return self.a.dot(x)
        out = np.zeros((self.nrows(), x.shape[1]), dtype=self.a.dtype)
        for j in range(x.shape[1]):
            out[:, j] = self.matvec(np.ascontiguousarray(x[:, j]))
        return out

    def conj_matmat(self, x):
        if self.route == "gemm":
            return conj_t(conj_t(x).dot(self.a))
        out = np.zeros((self.ncols(), x.shape[1]), dtype=self.a.dtype)
        for j in range(x.shape[1]):
            out[:, j] = self.conj_matvec(np.ascontiguousarray(x[:, j]))
        return out


def _as_op(op, route="gemm"):
    return op if hasattr(op, "matmat") else DenseOperator(np.asarray(op), route)


# ------------------------------------------------------------------- pivoted QR
def pivoted_qr(arr):
    """src/pivoted_qr.rs:25-31 + 81-119 + 121-183.

    Returns (q m x k', r k' x n upper-trapezoidal, ind 0-based), k' = min(m, n),
    with arr[:, ind] = q @ r.
    """
    arr = np.asarray(arr)
    m, n = arr.shape
    k = min(m, n)
    dtype = arr.dtype
    mat = np.asfortranarray(arr).copy(order="F")              # :28-29 column-major copy
    geqp3 = _lapack("geqp3", dtype)
    qr, jpvt, tau, _work, info = geqp3(mat)                   # :139-173 (jpvt = 0: all free)
    if info != 0:
        raise PivotedQRError(f"geqp3 info={info}")
    ind = (jpvt - 1).astype(np.int64)                         # :177
    r = np.triu(qr[:k, :])                                    # :100-102
    orgqr = _lapack("orgqr", dtype)
    qfull, _work, info = orgqr(np.asfortranarray(qr[:, :k]), tau)   # :104-111
    if info != 0:
        raise PivotedQRError(f"orgqr info={info}")
    q = np.ascontiguousarray(qfull[:, :k])                    # :113-114
    return q, np.ascontiguousarray(r), ind


def pivoted_lq(arr):
    """src/pivoted_qr.rs:32-41: pivoted QR of arr^H, transposed back."""
    q, r, ind = pivoted_qr(conj_t(arr))
    return conj_t(r), conj_t(q), ind          # l, q, ind


def compute_svd(arr):
    """src/compute_svd.rs:14-30: thin SVD via ?gesdd jobz='S'."""
    arr = np.asarray(arr)
    gesdd = _lapack("gesdd", arr.dtype)
    u, s, vt, info = gesdd(np.asfortranarray(arr), compute_uv=1, full_matrices=0)
    if info != 0:
        raise LinalgError(f"gesdd info={info}")
    return np.ascontiguousarray(u), s, np.ascontiguousarray(vt)


def _solve_upper(r11, rhs):
    """``solve_triangular(UPLO::Upper, Diag::NonUnit, rhs)`` (src/qr.rs:297-299)."""
    trtrs = _lapack("trtrs", r11.dtype)
    x, info = trtrs(r11, rhs, lower=0, trans=0, unitdiag=0)
    if info != 0:
        raise LinalgError(f"trtrs info={info}")
    return x


# --------------------------------------------------------------- ID containers
class ColumnID:
    """src/col_interp_decomp.rs:23-31, 44-86."""

    def __init__(self, c, z, col_ind):
        self.c, self.z, self.col_ind = c, z, np.asarray(col_ind)

    def nrows(self):
        return self.c.shape[0]

    def ncols(self):
        return self.z.shape[1]

    def rank(self):
        return self.c.shape[1]

    def to_mat(self):
        return self.c.dot(self.z)                                   # :64

    def dot(self, rhs):
        return self.c.dot(self.z.dot(rhs))                          # :141, :152

    def two_sided_id(self, order=None):
        """src/col_interp_decomp.rs:116-125 (quirk Q10: uncompressed LQ of C).  ``order``: replay with a prescribed
        pivot order (tie adjudication in the parity tests, see ``pivoted_qr_with_order``)."""
        row_id = LQ.compute_from(self.c, order=order).row_id()
        return TwoSidedID(c=row_id.x, x=row_id.r, r=self.z,
                          row_ind=row_id.row_ind, col_ind=self.col_ind)


class RowID:
    """src/row_interp_decomp.rs:25-33, 46-89."""

    def __init__(self, x, r, row_ind):
        self.x, self.r, self.row_ind = x, r, np.asarray(row_ind)

    def nrows(self):
        return self.x.shape[0]

    def ncols(self):
        return self.r.shape[1]

    def rank(self):
        return self.r.shape[0]

    def to_mat(self):
        return self.x.dot(self.r)                                   # :66

    def dot(self, rhs):
        return self.x.dot(self.r.dot(rhs))                          # :141, :152

    def two_sided_id(self, order=None):
        """src/row_interp_decomp.rs:120-130.  ``order``: see ``ColumnID.two_sided_id``."""
        col_id = QR.compute_from(self.r, order=order).column_id()
        return TwoSidedID(c=self.x, x=col_id.c, r=col_id.z,
                          row_ind=self.row_ind, col_ind=col_id.col_ind)


class TwoSidedID:
    """src/two_sided_interp_decomp.rs:19-30, 43-96 (A ~ C X R)."""

    def __init__(self, c, x, r, row_ind, col_ind):
        self.c, self.x, self.r = c, x, r
        self.row_ind, self.col_ind = np.asarray(row_ind), np.asarray(col_ind)

    def nrows(self):
        return self.c.shape[0]

    def ncols(self):
        return self.r.shape[1]

    def rank(self):
        return self.x.shape[0]

    def to_mat(self):
        return self.c.dot(self.x.dot(self.r))                       # :62-64

    def dot(self, rhs):
        return self.c.dot(self.x.dot(self.r.dot(rhs)))              # :160, :169


# ------------------------------------------------------------------ QR / LQ
def _check_tol(tol):
    assert (tol < 1.0) and (0.0 <= tol), "Require 0 <= tol < 1.0"   # src/qr.rs:99, 188


class QR:
    """src/qr.rs:31-40 + QRTraits 141-238 + impl 240-324."""

    def __init__(self, q, r, ind):
        self.q, self.r, self.ind = q, r, np.asarray(ind)

    def nrows(self):
        return self.q.shape[0]

    def ncols(self):
        return self.r.shape[1]

    def rank(self):
        return self.q.shape[1]

    def to_mat(self):
        """:160-166  q . r[:, inv(ind)]"""
        return self.q.dot(apply_permutation_matrix(self.r, self.ind, "COLINV"))

    def compress_qr_rank(self, max_rank):
        """:169-184"""
        max_rank = min(max_rank, self.q.shape[1])
        return QR(self.q[:, :max_rank].copy(), self.r[:max_rank, :].copy(), self.ind.copy())

    def compress_qr_tolerance(self, tol):
        """:187-200 (ratio formed in the scalar type, compared in f64)."""
        _check_tol(tol)
        d = np.diagonal(self.r)
        ratio = np.abs(d / self.r[0, 0]).astype(np.float64)
        pos = np.nonzero(ratio < tol)[0]
        if len(pos) == 0:
            raise CompressionError("Could not compress to desired tolerance")
        return self.compress_qr_rank(int(pos[0]))

    def compress(self, ctype):
        """:203-208"""
        if isinstance(ctype, ADAPTIVE):
            return self.compress_qr_tolerance(ctype.tol)
        return self.compress_qr_rank(ctype.rank)

    @staticmethod
    def compute_from(arr, order=None):
        """:251-253.  ``order`` (test infrastructure, not in the crate): replay with a prescribed pivot order."""
        return QR(*(pivoted_qr(arr) if order is None else pivoted_qr_with_order(arr, order)))

    @staticmethod
    def compute_from_range_estimate(rng_q, op, route="gemm", order=None):
        """:311-323   b = (A^H Q)^H ; pivoted QR of b ; q <- Q q_b"""
        op = _as_op(op, route)
        b = conj_t(op.conj_matmat(rng_q))
        qb, rb, ind = pivoted_qr(b) if order is None else pivoted_qr_with_order(b, order)
        return QR(rng_q.dot(qb), rb, ind)

    def column_id(self):
        """:270-309"""
        rank, ncols = self.rank(), self.ncols()
        dtype = self.r.dtype
        if rank == ncols:
            z = apply_permutation_matrix(np.eye(rank, dtype=dtype), self.ind, "COLINV")
            return ColumnID(self.q.dot(self.r), z, self.ind.copy())
        z = np.zeros((rank, ncols), dtype=dtype)
        z[:, :rank] = np.eye(rank, dtype=dtype)
        first_part = np.ascontiguousarray(self.r[:, :rank])
        c = self.q.dot(first_part)
        for index in range(ncols - rank):                      # one trtrs per column, :290-301
            col = np.ascontiguousarray(self.r[:, rank + index])
            z[:, rank + index] = _solve_upper(first_part, col)
        return ColumnID(c, apply_permutation_matrix(z, self.ind, "COLINV"), self.ind.copy())


class LQ:
    """src/qr.rs:42-51 + LQTraits 54-139 + impl 326-405."""

    def __init__(self, l, q, ind):
        self.l, self.q, self.ind = l, q, np.asarray(ind)

    def nrows(self):
        return self.l.shape[0]

    def ncols(self):
        return self.q.shape[1]

    def rank(self):
        return self.q.shape[0]

    def to_mat(self):
        """:73-78  l[inv(ind), :] . q"""
        return apply_permutation_matrix(self.l, self.ind, "ROWINV").dot(self.q)

    def compress_lq_rank(self, max_rank):
        """:81-96"""
        max_rank = min(max_rank, self.q.shape[0])
        return LQ(self.l[:, :max_rank].copy(), self.q[:max_rank, :].copy(), self.ind.copy())

    def compress_lq_tolerance(self, tol):
        """:99-111"""
        _check_tol(tol)
        d = np.diagonal(self.l)
        ratio = np.abs(d / self.l[0, 0]).astype(np.float64)
        pos = np.nonzero(ratio < tol)[0]
        if len(pos) == 0:
            raise CompressionError("Could not compress to desired tolerance")
        return self.compress_lq_rank(int(pos[0]))

    def compress(self, ctype):
        """:114-119"""
        if isinstance(ctype, ADAPTIVE):
            return self.compress_lq_tolerance(ctype.tol)
        return self.compress_lq_rank(ctype.rank)

    @staticmethod
    def compute_from(arr, order=None):
        """:354-362.  ``order``: see ``QR.compute_from``."""
        q, r, ind = pivoted_qr(conj_t(arr)) if order is None else pivoted_qr_with_order(conj_t(arr), order)
        return LQ(conj_t(r), conj_t(q), ind)

    def row_id(self):
        """:363-403"""
        rank, nrows = self.rank(), self.nrows()
        dtype = self.l.dtype
        if rank == nrows:
            x = apply_permutation_matrix(np.eye(rank, dtype=dtype), self.ind, "ROWINV")
            return RowID(x, self.l.dot(self.q), self.ind.copy())
        x = np.zeros((nrows, rank), dtype=dtype)
        x[:rank, :] = np.eye(rank, dtype=dtype)
        first_part = np.ascontiguousarray(self.l[:rank, :])
        r = first_part.dot(self.q)
        first_part_t = np.ascontiguousarray(first_part.T)       # plain transpose, :383
        for index in range(nrows - rank):                       # one trtrs per row, :384-395
            row = np.ascontiguousarray(self.l[rank + index, :])
            x[rank + index, :] = _solve_upper(first_part_t, row)
        return RowID(apply_permutation_matrix(x, self.ind, "ROWINV"), r, self.ind.copy())


# ------------------------------------------------------------------------- SVD
class SVD:
    """src/svd.rs:13-20 + SVDTraits 23-122 + impl 124-186."""

    def __init__(self, u, s, vt):
        self.u, self.s, self.vt = u, s, vt

    def nrows(self):
        return self.u.shape[0]

    def ncols(self):
        return self.vt.shape[1]

    def rank(self):
        return self.u.shape[1]

    def to_mat(self):
        """:42-54  u . diag(s) . vt"""
        scaled_vt = self.vt * self.s.astype(self.vt.dtype)[:, None]
        return self.u.dot(scaled_vt)

    def compress_svd_rank(self, max_rank):
        """:68-84"""
        max_rank = min(max_rank, len(self.s))
        return SVD(self.u[:, :max_rank].copy(), self.s[:max_rank].copy(),
                   self.vt[:max_rank, :].copy())

    def compress_svd_tolerance(self, tol):
        """:87-101"""
        _check_tol(tol)
        ratio = (self.s / self.s[0]).astype(np.float64)
        pos = np.nonzero(ratio < tol)[0]
        if len(pos) == 0:
            raise CompressionError("Could not compress to desired tolerance")
        return self.compress_svd_rank(int(pos[0]))

    def compress(self, ctype):
        """:60-65"""
        if isinstance(ctype, ADAPTIVE):
            return self.compress_svd_tolerance(ctype.tol)
        return self.compress_svd_rank(ctype.rank)

    def to_qr(self):
        """:150-163  pivoted QR of diag(s) vt ; q <- u q"""
        vt = self.vt * self.s.astype(self.vt.dtype)[:, None]
        qr = QR.compute_from(vt)
        qr.q = self.u.dot(qr.q)
        return qr

    @staticmethod
    def compute_from(arr):
        """:165-169"""
        return SVD(*compute_svd(arr))

    @staticmethod
    def compute_from_range_estimate(rng_q, op, route="gemm"):
        """:171-183  b = (A^H Q)^H ; SVD of b ; u <- Q u_b"""
        op = _as_op(op, route)
        b = conj_t(op.conj_matmat(rng_q))
        ub, s, vt = compute_svd(b)
        return SVD(rng_q.dot(ub), s, vt)


# ------------------------------------------------------------------- samplers
class OmegaStream:
    """Supplies the Gaussian blocks the samplers draw, in draw order.

    The reference takes ``rng: &mut R`` (src/random_sampling.rs:66-71); for parity
    runs every draw must be reproducible on both sides, so the stream is either a
    list of pre-built blocks or a seeded Philox generator (block b uses Philox
    stream id b, see ``oracle/philox.py``)."""

    def __init__(self, dtype, seed=None, blocks=None):
        self.dtype, self.seed, self.blocks, self.count = np.dtype(dtype), seed, blocks, 0
        self.drawn = []

    def draw(self, shape):
        if self.blocks is not None:
            blk = np.asarray(self.blocks[self.count])
            assert blk.shape == tuple(shape), (blk.shape, shape)
        else:
            from .philox import random_gaussian
            blk = random_gaussian(shape, self.dtype, self.seed, stream=self.count)
        self.count += 1
        self.drawn.append(blk)
        return blk


def max_col_norm(mat):
    """src/random_sampling.rs:175-199."""
    best = mat.real.dtype.type(0)
    for j in range(mat.shape[1]):
        best = max(best, np.linalg.norm(mat[:, j]))
    return best


def sample_range_by_rank(op, k, p, omega_stream, route="gemm"):
    """src/random_sampling.rs:103-118."""
    op = _as_op(op, route)
    n = op.ncols()                                            # named `m` in the crate (:109)
    omega = omega_stream.draw((n, k + p))
    basis = op.matmat(omega)
    qr = QR.compute_from(basis).compress(RANK(k))
    return qr.q.copy()


def sample_range_power_iteration(op, k, p, it_count, omega_stream, route="gemm"):
    """src/random_sampling.rs:131-160, quirk Q1 included: ``op_omega`` inside the
    loop (:150) is a fresh binding, so every trip restarts from A*Omega (:145)."""
    op = _as_op(op, route)
    n = op.ncols()
    omega = omega_stream.draw((n, k + p))
    op_omega = op.matmat(omega)
    res = op_omega.copy()
    for index in range(it_count):
        q = QR.compute_from(op_omega).q                       # always the OUTER op_omega
        w = QR.compute_from(op.conj_matmat(q)).q
        inner_op_omega = op.matmat(w)                         # shadows, dies each trip
        if index == it_count - 1:
            res = inner_op_omega
    compressed = QR.compute_from(res).compress(RANK(k))
    return compressed.q.copy()


def sample_range_adaptive(op, rel_tol, sample_size, omega_stream, route="gemm",
                          max_rank=None):
    """src/random_sampling.rs:223-274.  Returns (q, residuals).

    ``max_rank`` is a guard the crate lacks (quirk Q6: no upper bound on rank);
    exceeding it raises CompressionError, mirroring the C ABI."""
    op = _as_op(op, route)
    real_t = np.empty(0, dtype=op.a.dtype).real.dtype.type
    tol_factor = real_t(10.0 * math.sqrt(2.0 / math.pi))       # :229-232
    n = op.ncols()
    rel_tol_t = real_t(rel_tol)
    omega = omega_stream.draw((n, sample_size))
    op_omega = op.matmat(omega)
    operator_norm = max_col_norm(op_omega) * tol_factor        # :241
    max_norm = operator_norm
    q = np.zeros((op.nrows(), 0), dtype=op.a.dtype)
    b = np.zeros((0, op.ncols()), dtype=op.a.dtype)
    residuals = []
    while max_norm / operator_norm >= rel_tol_t:               # :248
        if q.shape[1] > 0:
            op_omega = op_omega - q.dot(conj_t(q).dot(op_omega))    # :250-252
        qq = QR.compute_from(op_omega).q                       # :254
        b = np.concatenate([b, conj_t(op.conj_matmat(qq))], axis=0)   # :256-260
        q = np.concatenate([q, qq], axis=1)                    # :262
        if max_rank is not None and q.shape[1] > max_rank:
            raise CompressionError("adaptive sampler exceeded max_rank")
        omega = omega_stream.draw((n, sample_size))            # :265
        op_omega = op.matmat(omega) - q.dot(b.dot(omega))      # :266
        max_norm = max_col_norm(op_omega) * tol_factor         # :269
        residuals.append((q.shape[1], float(max_norm / operator_norm)))   # :270
    return q, residuals


# ------------------------------------------------------------ test matrices
def random_orthogonal_matrix(shape, dtype, seed, stream=0):
    """src/random_matrix.rs:35-56 with a seeded Gaussian in place of thread_rng."""
    from .philox import random_gaussian
    m, n = shape
    swap = n > m
    if swap:
        m, n = n, m
    mat = random_gaussian((m, n), dtype, seed, stream=stream)
    u, _s, _vt = compute_svd(mat)
    return conj_t(u) if swap else u


def random_approximate_low_rank_matrix(shape, sigma_max, sigma_min, dtype, seed):
    """src/random_matrix.rs:70-93 (singular values ASCENDING, quirk Q4)."""
    assert sigma_min < sigma_max, "`sigma_min` must be smaller than `sigma_max`"
    assert sigma_min > 0.0, "`sigma_min` must be positive."
    dtype = np.dtype(dtype)
    min_dim = min(shape)
    u = random_orthogonal_matrix((shape[0], min_dim), dtype, seed, stream=101)
    vt = random_orthogonal_matrix((min_dim, shape[1]), dtype, seed, stream=102)
    singvals = np.geomspace(sigma_min, sigma_max, min_dim).astype(dtype)
    return np.ascontiguousarray(u.dot(np.diag(singvals).dot(vt)))


# ------------------------------------------------------------ parity checkers
def range_residual(a, q):
    """|| A - Q Q^H A ||_F / || A ||_F evaluated in the widest precision (the
    checker the parity contract names)."""
    wide = np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(q) else np.float64
    a = a.astype(wide)
    q = q.astype(wide)
    return np.linalg.norm(a - q.dot(conj_t(q).dot(a))) / np.linalg.norm(a)


def _wide_dtype(arr):
    return np.complex128 if np.iscomplexobj(arr) else np.float64


def pivot_gaps(arr, ind, upto=None):
    """Signed relative gap at every step of a pivoted QR of ``arr`` that takes its pivots in the order
    ``ind`` (double-precision Householder replay; single-precision inputs are widened exactly):

        gap[j] = (norm of the chosen column - largest norm among the other candidates) / norm of the chosen column

    evaluated on the trailing matrix of step j.  gap[j] > 0: the choice was the maximum, by that margin;
    gap[j] < 0: another column was larger by |gap[j]|.  Used to adjudicate bit-exactness of skeleton indices:
    a choice only counts as wrong where the gap exceeds 1e-6 (north_star)."""
    w = np.array(arr, dtype=_wide_dtype(arr))[:, np.asarray(ind)]
    m, n = w.shape
    k = min(m, n) if upto is None else min(m, n, int(upto))
    gaps = np.full(k, np.inf)
    tiny = np.finfo(np.float64).tiny
    for j in range(k):
        norms = np.linalg.norm(w[j:, j:], axis=0)
        if len(norms) > 1:
            others = np.max(norms[1:])
            gaps[j] = (norms[0] - others) / max(norms[0], others, tiny)
        x = w[j:, j].copy()
        nx = np.linalg.norm(x)
        if nx == 0:
            continue
        alpha = x[0]
        phase = alpha / abs(alpha) if alpha != 0 else 1.0
        x[0] += phase * nx
        v = x / np.linalg.norm(x)
        w[j:, j:] -= 2.0 * np.outer(v, np.conj(v).dot(w[j:, j:]))
    return gaps


def pivot_sequence_f64(arr):
    """the pivot sequence ?geqp3 picks when its norms are carried in DOUBLE precision on the given
    input -- dgeqp3 / zgeqp3 on the exactly widened f32 / c32 data (identical to ``pivoted_qr(arr)[2]`` for
    f64 / c64).  sgeqp3's own downdated norms are only good to ~sqrt(eps_f32) before its safeguard fires, so with
    thousands of candidate columns its sequence is not reproducible at the 1e-6 gap the contract names; this
    one is, and it is the sequence the CUDA path is held to."""
    return pivoted_qr(np.asarray(arr).astype(_wide_dtype(arr)))[2]


def pivoted_qr_with_order(arr, ind):
    """The factorisation ?geqp3 returns when it takes its pivots in the order ``ind``: unpivoted Householder QR
    (?geqrf + ?orgqr/?ungqr) of ``arr[:, ind]``, in the working precision, same post-processing as ``pivoted_qr``
    (src/pivoted_qr.rs:100-114).  Lets a parity test continue past a numerical tie: the oracle is replayed with the
    choice the device made at the tied step instead of skipping everything that follows."""
    arr = np.asarray(arr)
    ind = np.asarray(ind, dtype=np.int64)
    m, n = arr.shape
    k = min(m, n)
    mat = np.asfortranarray(arr[:, ind])
    geqrf = _lapack("geqrf", arr.dtype)
    qr, tau, _work, info = geqrf(mat)
    if info != 0:
        raise PivotedQRError(f"geqrf info={info}")
    r = np.triu(qr[:k, :])
    orgqr = _lapack("orgqr", arr.dtype)
    qfull, _work, info = orgqr(np.asfortranarray(qr[:, :k]), tau)
    if info != 0:
        raise PivotedQRError(f"orgqr info={info}")
    return np.ascontiguousarray(qfull[:, :k]), np.ascontiguousarray(r), ind.copy()


def check_pivot_sequence(arr, got, upto=None, tie=1e-6):
    """Adjudicates a pivot sequence ``got`` (first ``upto`` steps) produced for the input ``arr`` against the contract
    "bit-exact wherever the pivot norm gap exceeds ``tie`` relative".

    Every step of ``got`` is replayed in double precision and must have picked a column whose trailing norm is the
    maximum or within ``tie`` of it -- a complete check (every step, not only the first mismatch), which is what
    makes it possible to continue after a tie instead of skipping.  Returns a report

        {"identical": got == want on the first `upto` steps (want = pivot_sequence_f64(arr)),
         "first_divergence": None | (step, |gap| at that step),
         "ties": [(step, gap) for the steps where the choice was not the strict maximum],
         "min_gap": the smallest signed gap over the checked steps}

    and raises AssertionError naming the first step that violates the contract."""
    got = np.asarray(got, dtype=np.int64)
    m, n = np.asarray(arr).shape
    k = min(m, n) if upto is None else min(m, n, int(upto))
    want = pivot_sequence_f64(arr)
    identical = bool(np.array_equal(got[:k], want[:k]))
    report = {"identical": identical, "first_divergence": None, "ties": [], "min_gap": np.inf}
    if identical:
        return report
    assert sorted(got.tolist()) == list(range(n)), "pivot vector is not a permutation"
    gaps = pivot_gaps(arr, got, upto=k)
    j0 = int(np.nonzero(got[:k] != want[:k])[0][0])
    report["first_divergence"] = (j0, float(abs(gaps[j0])))
    report["ties"] = [(int(j), float(g)) for j, g in enumerate(gaps) if g < 0]
    report["min_gap"] = float(np.min(gaps))
    bad = np.nonzero(gaps < -tie)[0]
    assert len(bad) == 0, (f"pivot sequence violates the contract: first divergence from the double-precision ?geqp3 "
                           f"sequence at step {j0} (gap {abs(gaps[j0]):.3e}); step {int(bad[0])} picked a column whose norm "
                           f"is {abs(gaps[int(bad[0])]):.3e} (relative) below the maximum, limit {tie:.1e}")
    return report
