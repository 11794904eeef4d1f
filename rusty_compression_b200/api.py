"""Host-side mirror of the reference crate's public surface (src/lib.rs:90-102) over the C ABI.

Same names, argument meaning and error behaviour as the reference (and as the oracle,
``oracle/reference_path.py``), numpy in / numpy out, so the parity tests read like the
reference's own tests.  Everything numerical happens in librc_b200.so on the GPU; this module
only moves buffers and handles.  No CPU fallback: importing it without the built library or
without a CUDA device raises.

    QR.compute_from(a).compress(ADAPTIVE(1e-4)).column_id().two_sided_id().to_mat()
    q = sample_range_power_iteration(op, k, p, it_count, omega=..., seed=...)
    SVD.compute_from_range_estimate(q, op)
"""
import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import RcError, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

DTYPE_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}
PERM_MODE = {"COL": 0, "ROW": 1, "COLINV": 2, "ROWINV": 3}
VPERM_MODE = {"INV": 0, "NOINV": 1}


# ---- error classes, mirroring RustyCompressionError (src/types.rs:11-21)
class RustyCompressionError(Exception):
    pass


class LinalgError(RustyCompressionError):
    pass


class CompressionError(RustyCompressionError):
    pass


class LayoutError(RustyCompressionError):
    pass


class PivotedQRError(RustyCompressionError):
    pass


_STATUS_EXC = {1: LinalgError, 2: CompressionError, 3: LayoutError, 4: PivotedQRError}


@dataclass(frozen=True)
class ADAPTIVE:
    """CompressionType::ADAPTIVE(f64) (src/lib.rs:82-87)."""
    tol: float


@dataclass(frozen=True)
class RANK:
    """CompressionType::RANK(usize) (src/lib.rs:82-87)."""
    rank: int


class Context:
    """Owns an rc_ctx (device, stream, workspaces, communicator)."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = c_void_p()
        st = self.lib.rc_ctx_create(int(device), ctypes.byref(h))
        if st != 0:
            raise RcError(st, f"rc_ctx_create(device={device}) failed (is a B200 visible?)")
        self.h = h
        self.device = device

    def check(self, st):
        if st == 0:
            return
        msg = self.lib.rc_last_error_string(self.h).decode()
        if st == 5:
            # where the crate panics via assert!
            raise AssertionError(msg)
        exc = _STATUS_EXC.get(st)
        if exc:
            raise exc(msg)
        raise RcError(st, msg)

    def set_option(self, key, value):
        self.check(self.lib.rc_ctx_set_option(self.h, key.encode(), int(value)))

    def counter(self, key):
        v = c_int64()
        self.check(self.lib.rc_ctx_get_counter(self.h, key.encode(), ctypes.byref(v)))
        return v.value

    def reset_counters(self):
        self.lib.rc_ctx_reset_counters(self.h)

    def synchronize(self):
        self.check(self.lib.rc_ctx_synchronize(self.h))

    def pin(self, arr):
        """Page-lock the memory of a numpy array in place (rc_host_register) so that uploads from it run at the rate
        of the link and `DeviceMatrix.from_numpy_async` does not block; undo with `unpin(arr)` before it is freed."""
        self.check(self.lib.rc_host_register(self.h, c_void_p(arr.ctypes.data), arr.nbytes))
        return arr

    def unpin(self, arr):
        self.check(self.lib.rc_host_unregister(self.h, c_void_p(arr.ctypes.data)))

    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.rc_ctx_set_stream(self.h, c_void_p(cuda_stream_ptr)))

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = ctypes.create_string_buffer(unique_id, 128)
        self.check(self.lib.rc_ctx_comm_init(self.h, buf, rank, nranks))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rc_ctx_destroy(self.h)
            self.h = None


def comm_unique_id() -> bytes:
    lib = _lib.load()
    buf = ctypes.create_string_buffer(128)
    st = lib.rc_comm_get_unique_id(buf)
    if st != 0:
        raise RcError(st, "rc_comm_get_unique_id failed")
    return buf.raw


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _u64(arr):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64))
    return a, a.ctypes.data_as(ctypes.POINTER(c_uint64))


class DeviceMatrix:
    """Device-resident dense matrix.  For A itself this is the operator implementing the
    plugin API MatVec / MatMat / ConjMatVec / ConjMatMat (src/types.rs:40-101)."""

    def __init__(self, ctx, handle, owned=True):
        self.ctx, self.h, self.owned = ctx, handle, owned

    # -- construction
    @staticmethod
    def from_numpy(arr, ctx=None):
        ctx = ctx or default_context()
        arr = np.asarray(arr)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        if arr.dtype not in DTYPE_CODE:
            raise TypeError(f"unsupported dtype {arr.dtype}")
        es = arr.dtype.itemsize
        rs, cs = (arr.strides[0] // es, arr.strides[1] // es) if arr.size else (arr.shape[1], 1)
        if arr.size and (arr.strides[0] % es or arr.strides[1] % es or rs < 0 or cs < 0):
            arr = np.ascontiguousarray(arr)
            rs, cs = arr.shape[1], 1
        h = c_void_p()
        ctx.check(ctx.lib.rc_matrix_from_host(ctx.h, DTYPE_CODE[arr.dtype], c_void_p(arr.ctypes.data), arr.shape[0],
                                              arr.shape[1], rs, cs, ctypes.byref(h)))
        return DeviceMatrix(ctx, h)

    @staticmethod
    def from_numpy_async(arr, ctx=None):
        """Pipelined upload (rc_matrix_from_host_async): returns while the copy runs on the context's copy stream.
        `arr` must be a row-major 2-D array (pinned memory for a truly asynchronous transfer) that stays alive and
        unmodified until the copy is complete; call `.await_upload()` before any other use of the handle."""
        ctx = ctx or default_context()
        if arr.dtype not in DTYPE_CODE:
            raise TypeError(f"unsupported dtype {arr.dtype}")
        es = arr.dtype.itemsize
        if arr.ndim != 2 or (arr.size and (arr.strides[1] != es or arr.strides[0] % es or arr.strides[0] < arr.shape[1] * es)):
            raise ValueError("from_numpy_async takes a row-major 2-D array")
        rs = arr.strides[0] // es if arr.size else arr.shape[1]
        h = c_void_p()
        ctx.check(ctx.lib.rc_matrix_from_host_async(ctx.h, DTYPE_CODE[arr.dtype], c_void_p(arr.ctypes.data), arr.shape[0],
                                                    arr.shape[1], rs, ctypes.byref(h)))
        m = DeviceMatrix(ctx, h)
        m._upload_src = arr                      # keeps the host buffer alive while the copy is in flight
        return m

    def await_upload(self, block_host=False):
        """Order the context stream behind a pipelined upload (and, with block_host, wait for it on the host)."""
        self.ctx.check(self.ctx.lib.rc_matrix_await(self.ctx.h, self.h, 1 if block_host else 0))
        if block_host:
            self._upload_src = None
        return self

    @staticmethod
    def wrap_device(ptr, rows, cols, ld, dtype, ctx=None):
        ctx = ctx or default_context()
        h = c_void_p()
        ctx.check(ctx.lib.rc_matrix_wrap_device(ctx.h, DTYPE_CODE[np.dtype(dtype)], c_void_p(ptr), rows, cols, ld,
                                                ctypes.byref(h)))
        return DeviceMatrix(ctx, h)

    @staticmethod
    def random_gaussian(shape, dtype, seed, stream=0, row_offset=0, ctx=None):
        ctx = ctx or default_context()
        h = c_void_p()
        ctx.check(ctx.lib.rc_random_gaussian(ctx.h, DTYPE_CODE[np.dtype(dtype)], shape[0], shape[1], seed, stream,
                                             row_offset, ctypes.byref(h)))
        return DeviceMatrix(ctx, h)

    # -- properties
    @property
    def shape(self):
        return (self.ctx.lib.rc_matrix_rows(self.h), self.ctx.lib.rc_matrix_cols(self.h))

    @property
    def dtype(self):
        return CODE_DTYPE[self.ctx.lib.rc_matrix_dtype(self.h)]

    @property
    def device_ptr(self):
        return self.ctx.lib.rc_matrix_device_ptr(self.h)

    @property
    def ld(self):
        return self.ctx.lib.rc_matrix_ld(self.h)

    def nrows(self):
        return self.shape[0]

    def ncols(self):
        return self.shape[1]

    def set_shard(self, global_rows, row_offset):
        self.ctx.check(self.ctx.lib.rc_matrix_set_shard(self.h, global_rows, row_offset))
        return self

    def to_numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.ctx.check(self.ctx.lib.rc_matrix_to_host(self.ctx.h, self.h, c_void_p(out.ctypes.data)))
        return out

    # -- plugin API
    def copy_from(self, src):
        """self <- src on the device (same shape and dtype)."""
        self.ctx.check(self.ctx.lib.rc_matrix_copy(self.ctx.h, src.h, self.h))
        return self

    def matmat(self, x):
        x = _as_dev(x, self.ctx)
        h = c_void_p()
        self.ctx.check(self.ctx.lib.rc_matmat(self.ctx.h, self.h, x.h, ctypes.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def conj_matmat(self, x):
        x = _as_dev(x, self.ctx)
        h = c_void_p()
        self.ctx.check(self.ctx.lib.rc_conj_matmat(self.ctx.h, self.h, x.h, ctypes.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def matvec(self, x):
        return self.matmat(np.asarray(x).reshape(-1, 1)).to_numpy()[:, 0]

    def conj_matvec(self, x):
        return self.conj_matmat(np.asarray(x).reshape(-1, 1)).to_numpy()[:, 0]

    def free(self):
        # a handle that outlives its context is dropped, not freed: rc_ctx_destroy has released the context object the
        # handle points into, so freeing through it would be a use-after-free (the device buffer is leaked instead --
        # free the handles before Context.close())
        if self.owned and self.h and getattr(self.ctx, "h", None):
            self.ctx.lib.rc_matrix_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Operator(DeviceMatrix):
    """A matrix-free operator: the crate's plugin API (MatVec / ConjMatVec / MatMat / ConjMatMat implemented by
    the caller, src/types.rs:40-101).  `matmat(x, ncols, y, stream)` and `conj_matmat(x, ncols, z, stream)`
    receive DeviceMatrix views of the device buffers (x: cols x ncols -> y: rows x ncols; x: rows x ncols ->
    z: cols x ncols) and the cudaStream_t (as an int) to enqueue their work on.  Accepted by the samplers and by
    QR/SVD.compute_from_range_estimate in place of a dense DeviceMatrix."""

    def __init__(self, shape, dtype, matmat, conj_matmat=None, ctx=None):
        ctx = ctx or default_context()
        rows, cols = int(shape[0]), int(shape[1])
        dt = np.dtype(dtype)

        def wrap(fn, in_rows, out_rows):
            def cb(_user, x, ldx, ncols, y, ldy, stream):
                try:
                    xv = DeviceMatrix.wrap_device(x, in_rows, ncols, ldx, dt, ctx=ctx)
                    yv = DeviceMatrix.wrap_device(y, out_rows, ncols, ldy, dt, ctx=ctx)
                    fn(xv, int(ncols), yv, int(stream or 0))
                    return 0
                except Exception as e:          # never let an exception cross the C boundary
                    self.last_exception = e
                    return 1
            return _lib.MATMAT_FN(cb)

        self.last_exception = None
        self._cb_matmat = wrap(matmat, cols, rows)                      # keep the thunks alive with the handle
        self._cb_conj = wrap(conj_matmat, rows, cols) if conj_matmat else _lib.MATMAT_FN(0)
        h = c_void_p()
        ctx.check(ctx.lib.rc_operator_create(ctx.h, DTYPE_CODE[dt], rows, cols, self._cb_matmat, self._cb_conj, None,
                                             ctypes.byref(h)))
        super().__init__(ctx, h)


def _as_dev(x, ctx=None):
    if isinstance(x, DeviceMatrix):
        return x
    return DeviceMatrix.from_numpy(x, ctx)


def _borrow(ctx, ptr):
    return DeviceMatrix(ctx, c_void_p(ptr), owned=False)


def _get_ind(ctx, fn, h, n):
    out = np.empty(n, dtype=np.uint64)
    ctx.check(fn(h, out.ctypes.data_as(ctypes.POINTER(c_uint64)), n))
    return out.astype(np.int64)


class _Handle:
    _free = None

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h
        self._cache = {}

    def _mat(self, name, getter):
        if name not in self._cache:
            self._cache[name] = _borrow(self.ctx, getter(self.h)).to_numpy()
        return self._cache[name]

    def _dev(self, getter):
        return _borrow(self.ctx, getter(self.h))

    def free(self):
        if self.h and getattr(self.ctx, "h", None):          # (see DeviceMatrix.free)
            getattr(self.ctx.lib, self._free)(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _call_new(self, fn, cls, *args):
        h = c_void_p()
        self.ctx.check(fn(self.ctx.h, self.h, *args, ctypes.byref(h)))
        return cls(self.ctx, h)

    def _call_mat(self, fn, *args):
        h = c_void_p()
        self.ctx.check(fn(self.ctx.h, self.h, *args, ctypes.byref(h)))
        return DeviceMatrix(self.ctx, h).to_numpy()


class TwoSidedID(_Handle):
    """src/two_sided_interp_decomp.rs:19-30, 43-96."""
    _free = "rc_two_sided_id_free"

    c = property(lambda s: s._mat("c", s.ctx.lib.rc_two_sided_id_get_c))
    x = property(lambda s: s._mat("x", s.ctx.lib.rc_two_sided_id_get_x))
    r = property(lambda s: s._mat("r", s.ctx.lib.rc_two_sided_id_get_r))

    @property
    def row_ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_two_sided_id_get_row_ind, self.h, self.ctx.lib.rc_two_sided_id_row_ind_len(self.h))

    @property
    def col_ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_two_sided_id_get_col_ind, self.h, self.ctx.lib.rc_two_sided_id_col_ind_len(self.h))

    def nrows(self):
        return self.c.shape[0]

    def ncols(self):
        return self.r.shape[1]

    def rank(self):
        return self.x.shape[0]

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_two_sided_id_to_mat)

    def dot(self, rhs):
        rhs = np.asarray(rhs)
        vec = rhs.ndim == 1
        d = _as_dev(rhs.reshape(-1, 1) if vec else rhs, self.ctx)
        out = self._call_mat(self.ctx.lib.rc_two_sided_id_apply, d.h)
        return out[:, 0] if vec else out


class ColumnID(_Handle):
    """src/col_interp_decomp.rs:23-31, 44-86."""
    _free = "rc_column_id_free"

    c = property(lambda s: s._mat("c", s.ctx.lib.rc_column_id_get_c))
    z = property(lambda s: s._mat("z", s.ctx.lib.rc_column_id_get_z))

    @property
    def col_ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_column_id_get_col_ind, self.h, self.ctx.lib.rc_column_id_col_ind_len(self.h))

    def nrows(self):
        return self.c.shape[0]

    def ncols(self):
        return self.z.shape[1]

    def rank(self):
        return self.c.shape[1]

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_column_id_to_mat)

    def dot(self, rhs):
        rhs = np.asarray(rhs)
        vec = rhs.ndim == 1
        d = _as_dev(rhs.reshape(-1, 1) if vec else rhs, self.ctx)
        out = self._call_mat(self.ctx.lib.rc_column_id_apply, d.h)
        return out[:, 0] if vec else out

    def two_sided_id(self):
        return self._call_new(self.ctx.lib.rc_column_id_two_sided_id, TwoSidedID)


class RowID(_Handle):
    """src/row_interp_decomp.rs:25-33, 46-89."""
    _free = "rc_row_id_free"

    x = property(lambda s: s._mat("x", s.ctx.lib.rc_row_id_get_x))
    r = property(lambda s: s._mat("r", s.ctx.lib.rc_row_id_get_r))

    @property
    def row_ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_row_id_get_row_ind, self.h, self.ctx.lib.rc_row_id_row_ind_len(self.h))

    def nrows(self):
        return self.x.shape[0]

    def ncols(self):
        return self.r.shape[1]

    def rank(self):
        return self.r.shape[0]

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_row_id_to_mat)

    def dot(self, rhs):
        rhs = np.asarray(rhs)
        vec = rhs.ndim == 1
        d = _as_dev(rhs.reshape(-1, 1) if vec else rhs, self.ctx)
        out = self._call_mat(self.ctx.lib.rc_row_id_apply, d.h)
        return out[:, 0] if vec else out

    def two_sided_id(self):
        return self._call_new(self.ctx.lib.rc_row_id_two_sided_id, TwoSidedID)


class QR(_Handle):
    """src/qr.rs:31-40 + QRTraits (:141-238)."""
    _free = "rc_qr_free"

    q = property(lambda s: s._mat("q", s.ctx.lib.rc_qr_get_q))
    r = property(lambda s: s._mat("r", s.ctx.lib.rc_qr_get_r))

    @property
    def ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_qr_get_ind, self.h, self.ctx.lib.rc_qr_ncols(self.h))

    def q_device(self):
        return self._dev(self.ctx.lib.rc_qr_get_q)

    def nrows(self):
        return self.ctx.lib.rc_qr_nrows(self.h)

    def ncols(self):
        return self.ctx.lib.rc_qr_ncols(self.h)

    def rank(self):
        return self.ctx.lib.rc_qr_rank(self.h)

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_qr_to_mat)

    def compress_qr_rank(self, max_rank):
        return self._call_new(self.ctx.lib.rc_qr_compress_rank, QR, int(max_rank))

    def compress_qr_tolerance(self, tol):
        return self._call_new(self.ctx.lib.rc_qr_compress_tolerance, QR, float(tol))

    def compress(self, ctype):
        if isinstance(ctype, ADAPTIVE):
            return self.compress_qr_tolerance(ctype.tol)
        return self.compress_qr_rank(ctype.rank)

    def column_id(self):
        return self._call_new(self.ctx.lib.rc_qr_column_id, ColumnID)

    @staticmethod
    def new(q, r, ind, ctx=None):
        """`QR { q, r, ind }` from parts (pub fields, src/qr.rs:31-40)."""
        qd = _as_dev(q, ctx)
        rd = _as_dev(r, qd.ctx)
        p, pp = _u64(ind)
        h = c_void_p()
        qd.ctx.check(qd.ctx.lib.rc_qr_new(qd.ctx.h, qd.h, rd.h, pp, len(p), ctypes.byref(h)))
        return QR(qd.ctx, h)

    @staticmethod
    def compute_from(arr, ctx=None):
        d = _as_dev(arr, ctx)
        h = c_void_p()
        d.ctx.check(d.ctx.lib.rc_qr_compute_from(d.ctx.h, d.h, ctypes.byref(h)))
        return QR(d.ctx, h)

    @staticmethod
    def compute_from_range_estimate(rng_q, op, ctx=None):
        opd = _as_dev(op, ctx)
        qd = _as_dev(rng_q, opd.ctx)
        h = c_void_p()
        opd.ctx.check(opd.ctx.lib.rc_qr_compute_from_range_estimate(opd.ctx.h, qd.h, opd.h, ctypes.byref(h)))
        return QR(opd.ctx, h)


class LQ(_Handle):
    """src/qr.rs:42-51 + LQTraits (:54-139)."""
    _free = "rc_lq_free"

    l = property(lambda s: s._mat("l", s.ctx.lib.rc_lq_get_l))
    q = property(lambda s: s._mat("q", s.ctx.lib.rc_lq_get_q))

    @property
    def ind(self):
        return _get_ind(self.ctx, self.ctx.lib.rc_lq_get_ind, self.h, self.ctx.lib.rc_lq_nrows(self.h))

    def nrows(self):
        return self.ctx.lib.rc_lq_nrows(self.h)

    def ncols(self):
        return self.ctx.lib.rc_lq_ncols(self.h)

    def rank(self):
        return self.ctx.lib.rc_lq_rank(self.h)

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_lq_to_mat)

    def compress_lq_rank(self, max_rank):
        return self._call_new(self.ctx.lib.rc_lq_compress_rank, LQ, int(max_rank))

    def compress_lq_tolerance(self, tol):
        return self._call_new(self.ctx.lib.rc_lq_compress_tolerance, LQ, float(tol))

    def compress(self, ctype):
        if isinstance(ctype, ADAPTIVE):
            return self.compress_lq_tolerance(ctype.tol)
        return self.compress_lq_rank(ctype.rank)

    def row_id(self):
        return self._call_new(self.ctx.lib.rc_lq_row_id, RowID)

    @staticmethod
    def new(l, q, ind, ctx=None):
        """`LQ { l, q, ind }` from parts (pub fields, src/qr.rs:42-51)."""
        ld = _as_dev(l, ctx)
        qd = _as_dev(q, ld.ctx)
        p, pp = _u64(ind)
        h = c_void_p()
        ld.ctx.check(ld.ctx.lib.rc_lq_new(ld.ctx.h, ld.h, qd.h, pp, len(p), ctypes.byref(h)))
        return LQ(ld.ctx, h)

    @staticmethod
    def compute_from(arr, ctx=None):
        d = _as_dev(arr, ctx)
        h = c_void_p()
        d.ctx.check(d.ctx.lib.rc_lq_compute_from(d.ctx.h, d.h, ctypes.byref(h)))
        return LQ(d.ctx, h)


class SVD(_Handle):
    """src/svd.rs:13-20 + SVDTraits (:23-122)."""
    _free = "rc_svd_free"

    u = property(lambda s: s._mat("u", s.ctx.lib.rc_svd_get_u))
    vt = property(lambda s: s._mat("vt", s.ctx.lib.rc_svd_get_vt))

    @property
    def s(self):
        n = self.ctx.lib.rc_svd_rank(self.h)
        out = np.empty(n, dtype=np.float64)
        self.ctx.check(self.ctx.lib.rc_svd_get_s(self.h, out.ctypes.data_as(ctypes.POINTER(c_double)), n))
        real = np.empty(0, dtype=_borrow(self.ctx, self.ctx.lib.rc_svd_get_u(self.h)).dtype).real.dtype
        return out.astype(real)

    def s_f64(self):
        n = self.ctx.lib.rc_svd_rank(self.h)
        out = np.empty(n, dtype=np.float64)
        self.ctx.check(self.ctx.lib.rc_svd_get_s(self.h, out.ctypes.data_as(ctypes.POINTER(c_double)), n))
        return out

    def nrows(self):
        return self.u.shape[0]

    def ncols(self):
        return self.vt.shape[1]

    def rank(self):
        return self.ctx.lib.rc_svd_rank(self.h)

    def to_mat(self):
        return self._call_mat(self.ctx.lib.rc_svd_to_mat)

    def to_qr(self):
        return self._call_new(self.ctx.lib.rc_svd_to_qr, QR)

    def compress_svd_rank(self, max_rank):
        return self._call_new(self.ctx.lib.rc_svd_compress_rank, SVD, int(max_rank))

    def compress_svd_tolerance(self, tol):
        return self._call_new(self.ctx.lib.rc_svd_compress_tolerance, SVD, float(tol))

    def compress(self, ctype):
        if isinstance(ctype, ADAPTIVE):
            return self.compress_svd_tolerance(ctype.tol)
        return self.compress_svd_rank(ctype.rank)

    @staticmethod
    def new(u, s, vt, ctx=None):
        """`SVD { u, s, vt }` from parts (pub fields, src/svd.rs:13-20)."""
        ud = _as_dev(u, ctx)
        vd = _as_dev(vt, ud.ctx)
        sv = np.ascontiguousarray(np.asarray(s, dtype=np.float64))
        h = c_void_p()
        ud.ctx.check(ud.ctx.lib.rc_svd_new(ud.ctx.h, ud.h, sv.ctypes.data_as(ctypes.POINTER(c_double)), len(sv), vd.h,
                                           ctypes.byref(h)))
        return SVD(ud.ctx, h)

    @staticmethod
    def compute_from(arr, ctx=None):
        d = _as_dev(arr, ctx)
        h = c_void_p()
        d.ctx.check(d.ctx.lib.rc_svd_compute_from(d.ctx.h, d.h, ctypes.byref(h)))
        return SVD(d.ctx, h)

    @staticmethod
    def compute_from_range_estimate(rng_q, op, ctx=None):
        opd = _as_dev(op, ctx)
        qd = _as_dev(rng_q, opd.ctx)
        h = c_void_p()
        opd.ctx.check(opd.ctx.lib.rc_svd_compute_from_range_estimate(opd.ctx.h, qd.h, opd.h, ctypes.byref(h)))
        return SVD(opd.ctx, h)


# ------------------------------------------------------------------ free functions
def pivoted_qr(arr, ctx=None):
    """PivotedQR::pivoted_qr (src/pivoted_qr.rs:25-31) -> (q, r, ind)."""
    qr = QR.compute_from(arr, ctx)
    return qr.q, qr.r, qr.ind


def pivoted_lq(arr, ctx=None):
    """PivotedQR::pivoted_lq (src/pivoted_qr.rs:32-41) -> (l, q, ind)."""
    lq = LQ.compute_from(arr, ctx)
    return lq.l, lq.q, lq.ind


def compute_svd(arr, ctx=None):
    """ComputeSVD::compute_svd (src/compute_svd.rs:14-30) -> (u, s, vt)."""
    svd = SVD.compute_from(arr, ctx)
    return svd.u, svd.s, svd.vt


def invert_permutation_vector(perm):
    lib = _lib.load()
    p, pp = _u64(perm)
    out = np.empty(len(p), dtype=np.uint64)
    st = lib.rc_invert_permutation_vector(pp, len(p), out.ctypes.data_as(ctypes.POINTER(c_uint64)))
    if st != 0:
        raise AssertionError("not a permutation")
    return out.astype(np.int64)


def apply_permutation_matrix(mat, index_array, mode, ctx=None):
    d = _as_dev(mat, ctx)
    p, pp = _u64(index_array)
    h = c_void_p()
    d.ctx.check(d.ctx.lib.rc_apply_permutation_matrix(d.ctx.h, d.h, pp, len(p), PERM_MODE[mode], ctypes.byref(h)))
    return DeviceMatrix(d.ctx, h).to_numpy()


def apply_permutation_vector(vec, index_array, mode, ctx=None):
    vec = np.asarray(vec)
    d = _as_dev(vec.reshape(1, -1), ctx)
    p, pp = _u64(index_array)
    h = c_void_p()
    d.ctx.check(d.ctx.lib.rc_apply_permutation_vector(d.ctx.h, d.h, pp, len(p), VPERM_MODE[mode], ctypes.byref(h)))
    return DeviceMatrix(d.ctx, h).to_numpy()[0]


def rel_diff_fro(first, second, ctx=None):
    a = _as_dev(first, ctx)
    b = _as_dev(second, a.ctx)
    out = c_double()
    a.ctx.check(a.ctx.lib.rc_rel_diff_fro(a.ctx.h, a.h, b.h, ctypes.byref(out)))
    return out.value


def rel_diff_l2(first, second, ctx=None):
    a = _as_dev(np.asarray(first).reshape(1, -1), ctx)
    b = _as_dev(np.asarray(second).reshape(1, -1), a.ctx)
    out = c_double()
    a.ctx.check(a.ctx.lib.rc_rel_diff_l2(a.ctx.h, a.h, b.h, ctypes.byref(out)))
    return out.value


def max_col_norm(mat, ctx=None):
    d = _as_dev(mat, ctx)
    out = c_double()
    d.ctx.check(d.ctx.lib.rc_max_col_norm(d.ctx.h, d.h, ctypes.byref(out)))
    return out.value


def random_gaussian(shape, dtype, seed, stream=0, row_offset=0, ctx=None):
    return DeviceMatrix.random_gaussian(shape, dtype, seed, stream, row_offset, ctx).to_numpy()


def random_orthogonal_matrix(shape, dtype, seed, stream=0, ctx=None):
    ctx = ctx or default_context()
    h = c_void_p()
    ctx.check(ctx.lib.rc_random_orthogonal_matrix(ctx.h, DTYPE_CODE[np.dtype(dtype)], shape[0], shape[1], seed, stream,
                                                  ctypes.byref(h)))
    return DeviceMatrix(ctx, h).to_numpy()


def random_approximate_low_rank_matrix(shape, sigma_max, sigma_min, dtype, seed, ctx=None, device=False):
    ctx = ctx or default_context()
    h = c_void_p()
    ctx.check(ctx.lib.rc_random_approximate_low_rank_matrix(ctx.h, DTYPE_CODE[np.dtype(dtype)], shape[0], shape[1],
                                                            sigma_max, sigma_min, seed, ctypes.byref(h)))
    d = DeviceMatrix(ctx, h)
    return d if device else d.to_numpy()


def decaying_spectrum_matrix(shape, dtype, seed, r0=512, decade_every=16.0, row_offset=0, ctx=None):
    """Device-generated bench input; returns a DeviceMatrix."""
    ctx = ctx or default_context()
    h = c_void_p()
    ctx.check(ctx.lib.rc_decaying_spectrum_matrix(ctx.h, DTYPE_CODE[np.dtype(dtype)], shape[0], shape[1], r0,
                                                  decade_every, seed, row_offset, ctypes.byref(h)))
    return DeviceMatrix(ctx, h)


def helmholtz_kernel_matrix(shape, dtype, seed=7, kappa=20.0, shift=1.5, row_offset=0, ctx=None):
    """Device-generated config-5 input; the host mirror is oracle.inputs.helmholtz_kernel_matrix_philox."""
    ctx = ctx or default_context()
    h = c_void_p()
    ctx.check(ctx.lib.rc_helmholtz_kernel_matrix(ctx.h, DTYPE_CODE[np.dtype(dtype)], shape[0], shape[1], seed, kappa, shift,
                                                 row_offset, ctypes.byref(h)))
    return DeviceMatrix(ctx, h)


def tall_shard_matrix(row_offset, rows, cols, dtype, seed, m_total, r0=512, decade_every=64.0, ctx=None):
    """Device-generated rows [row_offset, row_offset + rows) of the config-4 input."""
    ctx = ctx or default_context()
    h = c_void_p()
    ctx.check(ctx.lib.rc_tall_shard_matrix(ctx.h, DTYPE_CODE[np.dtype(dtype)], rows, cols, r0, decade_every, seed,
                                           row_offset, m_total, ctypes.byref(h)))
    return DeviceMatrix(ctx, h)


def _omega_handle(omega, ctx):
    if omega is None:
        return None, c_void_p(None)
    d = _as_dev(omega, ctx)
    return d, d.h


def sample_range_by_rank(op, k, p, omega=None, seed=0, ctx=None, device=False):
    """SampleRange::sample_range_by_rank (src/random_sampling.rs:103-118)."""
    opd = _as_dev(op, ctx)
    keep, oh = _omega_handle(omega, opd.ctx)
    h = c_void_p()
    opd.ctx.check(opd.ctx.lib.rc_sample_range_by_rank(opd.ctx.h, opd.h, k, p, oh, seed, ctypes.byref(h)))
    q = DeviceMatrix(opd.ctx, h)
    return q if device else q.to_numpy()


def sample_range_power_iteration(op, k, p, it_count, omega=None, seed=0, ctx=None, device=False):
    """SampleRangePowerIteration::sample_range_power_iteration (src/random_sampling.rs:131-160)."""
    opd = _as_dev(op, ctx)
    keep, oh = _omega_handle(omega, opd.ctx)
    h = c_void_p()
    opd.ctx.check(opd.ctx.lib.rc_sample_range_power_iteration(opd.ctx.h, opd.h, k, p, it_count, oh, seed,
                                                              ctypes.byref(h)))
    q = DeviceMatrix(opd.ctx, h)
    return q if device else q.to_numpy()


def sample_range_adaptive(op, rel_tol, sample_size, omega_blocks=None, seed=0, max_rank=0, ctx=None, device=False):
    """AdaptiveSampling::sample_range_adaptive (src/random_sampling.rs:223-274) -> (q, residuals).
    omega_blocks: None or a list of n x sample_size arrays / one n x (B*sample_size) array (draw order)."""
    opd = _as_dev(op, ctx)
    if omega_blocks is not None and not isinstance(omega_blocks, (np.ndarray, DeviceMatrix)):
        omega_blocks = np.concatenate([np.asarray(b) for b in omega_blocks], axis=1)
    keep, oh = _omega_handle(omega_blocks, opd.ctx)
    cap = 4096
    hr = np.zeros(cap, dtype=np.uint64)
    hv = np.zeros(cap, dtype=np.float64)
    hl = c_size_t()
    h = c_void_p()
    opd.ctx.check(opd.ctx.lib.rc_sample_range_adaptive(
        opd.ctx.h, opd.h, float(rel_tol), int(sample_size), oh, seed, int(max_rank), ctypes.byref(h),
        hr.ctypes.data_as(ctypes.POINTER(c_uint64)), hv.ctypes.data_as(ctypes.POINTER(c_double)), cap, ctypes.byref(hl)))
    q = DeviceMatrix(opd.ctx, h)
    n = min(hl.value, cap)
    residuals = [(int(hr[i]), float(hv[i])) for i in range(n)]
    return (q if device else q.to_numpy()), residuals
