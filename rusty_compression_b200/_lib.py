"""ctypes binding of include/rc_api.h (the C ABI is the product; this file is the harness-side
stub a Python host needs, the analogue of the Rust `-sys` crate shown in INTEGRATION.md).

There is NO fallback: if librc_b200.so is missing or a call fails, this module raises."""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librc_b200.so")

RC_OK = 0
STATUS_NAMES = {0: "OK", 1: "LINALG_ERROR", 2: "COMPRESSION_ERROR", 3: "LAYOUT_ERROR", 4: "PIVOTED_QR_ERROR",
                5: "INVALID_ARGUMENT", 6: "CUDA_ERROR", 7: "NCCL_ERROR", 8: "OUT_OF_MEMORY"}

H = c_void_p              # every opaque handle
PH = POINTER(c_void_p)
PU64 = POINTER(c_uint64)
PD = POINTER(c_double)

# rc_matmat_fn: int (*)(void* user, const void* x, int64 ldx, int64 ncols, void* y, int64 ldy, void* cuda_stream)
MATMAT_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p)

# name -> (restype, argtypes); must list every symbol rc_api.h declares (tests/test_abi_symbols.py)
SIGNATURES = {
    "rc_version": (c_int, []),
    "rc_ctx_create": (c_int, [c_int, PH]),
    "rc_ctx_destroy": (c_int, [H]),
    "rc_ctx_set_stream": (c_int, [H, c_void_p]),
    "rc_ctx_synchronize": (c_int, [H]),
    "rc_last_error_string": (c_char_p, [H]),
    "rc_ctx_set_option": (c_int, [H, c_char_p, c_int64]),
    "rc_ctx_get_counter": (c_int, [H, c_char_p, POINTER(c_int64)]),
    "rc_ctx_reset_counters": (c_int, [H]),
    "rc_comm_get_unique_id": (c_int, [c_void_p]),
    "rc_ctx_comm_init": (c_int, [H, c_void_p, c_int, c_int]),
    "rc_ctx_comm_info": (c_int, [H, POINTER(c_int), POINTER(c_int)]),
    "rc_matrix_create": (c_int, [H, c_int, c_int64, c_int64, PH]),
    "rc_matrix_from_host": (c_int, [H, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, PH]),
    "rc_matrix_from_host_async": (c_int, [H, c_int, c_void_p, c_int64, c_int64, c_int64, PH]),
    "rc_matrix_await": (c_int, [H, H, c_int]),
    "rc_host_register": (c_int, [H, c_void_p, c_size_t]),
    "rc_host_unregister": (c_int, [H, c_void_p]),
    "rc_matrix_wrap_device": (c_int, [H, c_int, c_void_p, c_int64, c_int64, c_int64, PH]),
    "rc_column_id_col_ind_len": (c_size_t, [H]),
    "rc_row_id_row_ind_len": (c_size_t, [H]),
    "rc_two_sided_id_row_ind_len": (c_size_t, [H]),
    "rc_two_sided_id_col_ind_len": (c_size_t, [H]),
    "rc_matrix_copy": (c_int, [H, H, H]),
    "rc_operator_create": (c_int, [H, c_int, c_int64, c_int64, MATMAT_FN, MATMAT_FN, c_void_p, PH]),
    "rc_matrix_to_host": (c_int, [H, H, c_void_p]),
    "rc_matrix_to_device": (c_int, [H, H, c_void_p]),
    "rc_matrix_free": (c_int, [H]),
    "rc_matrix_rows": (c_int64, [H]),
    "rc_matrix_cols": (c_int64, [H]),
    "rc_matrix_ld": (c_int64, [H]),
    "rc_matrix_dtype": (c_int, [H]),
    "rc_matrix_device_ptr": (c_void_p, [H]),
    "rc_matrix_set_shard": (c_int, [H, c_int64, c_int64]),
    "rc_matmat": (c_int, [H, H, H, PH]),
    "rc_conj_matmat": (c_int, [H, H, H, PH]),
    "rc_random_gaussian": (c_int, [H, c_int, c_int64, c_int64, c_uint64, c_uint32, c_int64, PH]),
    "rc_random_orthogonal_matrix": (c_int, [H, c_int, c_int64, c_int64, c_uint64, c_uint32, PH]),
    "rc_random_approximate_low_rank_matrix": (c_int, [H, c_int, c_int64, c_int64, c_double, c_double, c_uint64, PH]),
    "rc_decaying_spectrum_matrix": (c_int, [H, c_int, c_int64, c_int64, c_int64, c_double, c_uint64, c_int64, PH]),
    "rc_helmholtz_kernel_matrix": (c_int, [H, c_int, c_int64, c_int64, c_uint64, c_double, c_double, c_int64, PH]),
    "rc_tall_shard_matrix": (c_int, [H, c_int, c_int64, c_int64, c_int64, c_double, c_uint64, c_int64, c_int64, PH]),
    "rc_rel_diff_fro": (c_int, [H, H, H, PD]),
    "rc_rel_diff_l2": (c_int, [H, H, H, PD]),
    "rc_max_col_norm": (c_int, [H, H, PD]),
    "rc_invert_permutation_vector": (c_int, [PU64, c_size_t, PU64]),
    "rc_apply_permutation_matrix": (c_int, [H, H, PU64, c_size_t, c_int, PH]),
    "rc_apply_permutation_vector": (c_int, [H, H, PU64, c_size_t, c_int, PH]),
    "rc_sample_range_by_rank": (c_int, [H, H, c_int64, c_int64, H, c_uint64, PH]),
    "rc_sample_range_power_iteration": (c_int, [H, H, c_int64, c_int64, c_int64, H, c_uint64, PH]),
    "rc_sample_range_adaptive": (c_int, [H, H, c_double, c_int64, H, c_uint64, c_int64, PH, PU64, PD, c_size_t,
                                         POINTER(c_size_t)]),
    "rc_qr_compute_from": (c_int, [H, H, PH]),
    "rc_qr_new": (c_int, [H, H, H, PU64, c_size_t, PH]),
    "rc_lq_new": (c_int, [H, H, H, PU64, c_size_t, PH]),
    "rc_svd_new": (c_int, [H, H, PD, c_size_t, H, PH]),
    "rc_qr_compute_from_range_estimate": (c_int, [H, H, H, PH]),
    "rc_qr_compress_rank": (c_int, [H, H, c_int64, PH]),
    "rc_qr_compress_tolerance": (c_int, [H, H, c_double, PH]),
    "rc_qr_to_mat": (c_int, [H, H, PH]),
    "rc_qr_column_id": (c_int, [H, H, PH]),
    "rc_qr_get_q": (c_void_p, [H]),
    "rc_qr_get_r": (c_void_p, [H]),
    "rc_qr_rank": (c_int64, [H]),
    "rc_qr_nrows": (c_int64, [H]),
    "rc_qr_ncols": (c_int64, [H]),
    "rc_qr_get_ind": (c_int, [H, PU64, c_size_t]),
    "rc_qr_free": (c_int, [H]),
    "rc_lq_compute_from": (c_int, [H, H, PH]),
    "rc_lq_compress_rank": (c_int, [H, H, c_int64, PH]),
    "rc_lq_compress_tolerance": (c_int, [H, H, c_double, PH]),
    "rc_lq_to_mat": (c_int, [H, H, PH]),
    "rc_lq_row_id": (c_int, [H, H, PH]),
    "rc_lq_get_l": (c_void_p, [H]),
    "rc_lq_get_q": (c_void_p, [H]),
    "rc_lq_rank": (c_int64, [H]),
    "rc_lq_nrows": (c_int64, [H]),
    "rc_lq_ncols": (c_int64, [H]),
    "rc_lq_get_ind": (c_int, [H, PU64, c_size_t]),
    "rc_lq_free": (c_int, [H]),
    "rc_svd_compute_from": (c_int, [H, H, PH]),
    "rc_svd_compute_from_range_estimate": (c_int, [H, H, H, PH]),
    "rc_svd_compress_rank": (c_int, [H, H, c_int64, PH]),
    "rc_svd_compress_tolerance": (c_int, [H, H, c_double, PH]),
    "rc_svd_to_mat": (c_int, [H, H, PH]),
    "rc_svd_to_qr": (c_int, [H, H, PH]),
    "rc_svd_get_u": (c_void_p, [H]),
    "rc_svd_get_vt": (c_void_p, [H]),
    "rc_svd_rank": (c_int64, [H]),
    "rc_svd_get_s": (c_int, [H, PD, c_size_t]),
    "rc_svd_free": (c_int, [H]),
    "rc_column_id_new": (c_int, [H, H, H, PU64, c_size_t, PH]),
    "rc_column_id_get_c": (c_void_p, [H]),
    "rc_column_id_get_z": (c_void_p, [H]),
    "rc_column_id_get_col_ind": (c_int, [H, PU64, c_size_t]),
    "rc_column_id_to_mat": (c_int, [H, H, PH]),
    "rc_column_id_apply": (c_int, [H, H, H, PH]),
    "rc_column_id_two_sided_id": (c_int, [H, H, PH]),
    "rc_column_id_free": (c_int, [H]),
    "rc_row_id_new": (c_int, [H, H, H, PU64, c_size_t, PH]),
    "rc_row_id_get_x": (c_void_p, [H]),
    "rc_row_id_get_r": (c_void_p, [H]),
    "rc_row_id_get_row_ind": (c_int, [H, PU64, c_size_t]),
    "rc_row_id_to_mat": (c_int, [H, H, PH]),
    "rc_row_id_apply": (c_int, [H, H, H, PH]),
    "rc_row_id_two_sided_id": (c_int, [H, H, PH]),
    "rc_row_id_free": (c_int, [H]),
    "rc_two_sided_id_new": (c_int, [H, H, H, H, PU64, c_size_t, PU64, c_size_t, PH]),
    "rc_two_sided_id_get_c": (c_void_p, [H]),
    "rc_two_sided_id_get_x": (c_void_p, [H]),
    "rc_two_sided_id_get_r": (c_void_p, [H]),
    "rc_two_sided_id_get_row_ind": (c_int, [H, PU64, c_size_t]),
    "rc_two_sided_id_get_col_ind": (c_int, [H, PU64, c_size_t]),
    "rc_two_sided_id_to_mat": (c_int, [H, H, PH]),
    "rc_two_sided_id_apply": (c_int, [H, H, H, PH]),
    "rc_two_sided_id_free": (c_int, [H]),
}

_lib = None


class RcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def load():
    """Load the CUDA library.  Raises ImportError loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m rusty_compression_b200.build`).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
