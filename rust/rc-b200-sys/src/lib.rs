// AUTO-GENERATED from include/rc_api.h by the repo's generator (see rust/README.md). Source only.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct rc_ctx { _private: [u8; 0] }
#[repr(C)] pub struct rc_matrix { _private: [u8; 0] }
#[repr(C)] pub struct rc_qr { _private: [u8; 0] }
#[repr(C)] pub struct rc_lq { _private: [u8; 0] }
#[repr(C)] pub struct rc_svd { _private: [u8; 0] }
#[repr(C)] pub struct rc_column_id { _private: [u8; 0] }
#[repr(C)] pub struct rc_row_id { _private: [u8; 0] }
#[repr(C)] pub struct rc_two_sided_id { _private: [u8; 0] }

pub const RC_F32: c_int = 0;
pub const RC_F64: c_int = 1;
pub const RC_C32: c_int = 2;
pub const RC_C64: c_int = 3;
pub const RC_OK: c_int = 0;
pub const RC_LINALG_ERROR: c_int = 1;
pub const RC_COMPRESSION_ERROR: c_int = 2;
pub const RC_LAYOUT_ERROR: c_int = 3;
pub const RC_PIVOTED_QR_ERROR: c_int = 4;
pub const RC_INVALID_ARGUMENT: c_int = 5;

/// `rc_matmat_fn` of include/rc_api.h: device-side product callback of a matrix-free operator.
pub type rc_matmat_fn = Option<unsafe extern "C" fn(user: *mut c_void, x: *const c_void, ldx: i64, ncols: i64, y: *mut c_void, ldy: i64, cuda_stream: *mut c_void) -> c_int>;

extern "C" {
    pub fn rc_version() -> c_int;
    pub fn rc_ctx_create(device: c_int, out: *mut *mut rc_ctx) -> c_int;
    pub fn rc_ctx_destroy(ctx: *mut rc_ctx) -> c_int;
    pub fn rc_ctx_set_stream(ctx: *mut rc_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn rc_ctx_synchronize(ctx: *mut rc_ctx) -> c_int;
    pub fn rc_last_error_string(ctx: *mut rc_ctx) -> *const c_char;
    pub fn rc_ctx_set_option(ctx: *mut rc_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn rc_ctx_get_counter(ctx: *mut rc_ctx, key: *const c_char, out: *mut i64) -> c_int;
    pub fn rc_ctx_reset_counters(ctx: *mut rc_ctx) -> c_int;
    pub fn rc_comm_get_unique_id(out_id_128_bytes: *mut c_void) -> c_int;
    pub fn rc_ctx_comm_init(ctx: *mut rc_ctx, id_128_bytes: *const c_void, rank: c_int, nranks: c_int) -> c_int;
    pub fn rc_ctx_comm_info(ctx: *mut rc_ctx, rank: *mut c_int, nranks: *mut c_int) -> c_int;
    pub fn rc_matrix_create(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_matrix_from_host(ctx: *mut rc_ctx, dtype: c_int, host: *const c_void, rows: i64, cols: i64, row_stride: i64, col_stride: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_matrix_from_host_async(ctx: *mut rc_ctx, dtype: c_int, host: *const c_void, rows: i64, cols: i64, row_stride: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_host_register(ctx: *mut rc_ctx, host: *mut c_void, bytes: usize) -> c_int;
    pub fn rc_host_unregister(ctx: *mut rc_ctx, host: *mut c_void) -> c_int;
    pub fn rc_matrix_await(ctx: *mut rc_ctx, m: *mut rc_matrix, block_host: c_int) -> c_int;
    pub fn rc_matrix_wrap_device(ctx: *mut rc_ctx, dtype: c_int, device_ptr: *mut c_void, rows: i64, cols: i64, ld: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_matrix_to_host(ctx: *mut rc_ctx, m: *const rc_matrix, host: *mut c_void) -> c_int;
    pub fn rc_matrix_to_device(ctx: *mut rc_ctx, m: *const rc_matrix, device_ptr: *mut c_void) -> c_int;
    pub fn rc_matrix_free(m: *mut rc_matrix) -> c_int;
    pub fn rc_matrix_rows(m: *const rc_matrix) -> i64;
    pub fn rc_matrix_cols(m: *const rc_matrix) -> i64;
    pub fn rc_matrix_ld(m: *const rc_matrix) -> i64;
    pub fn rc_matrix_dtype(m: *const rc_matrix) -> c_int;
    pub fn rc_matrix_device_ptr(m: *const rc_matrix) -> *mut c_void;
    pub fn rc_matrix_set_shard(m: *mut rc_matrix, global_rows: i64, row_offset: i64) -> c_int;
    pub fn rc_matmat(ctx: *mut rc_ctx, a: *const rc_matrix, x: *const rc_matrix, y: *mut *mut rc_matrix) -> c_int;
    pub fn rc_conj_matmat(ctx: *mut rc_ctx, a: *const rc_matrix, x: *const rc_matrix, z: *mut *mut rc_matrix) -> c_int;
    pub fn rc_random_gaussian(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, seed: u64, stream: u32, row_offset: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_random_orthogonal_matrix(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, seed: u64, stream: u32, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_random_approximate_low_rank_matrix(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, sigma_max: f64, sigma_min: f64, seed: u64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_decaying_spectrum_matrix(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, r0: i64, decade_every: f64, seed: u64, row_offset: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_column_id_col_ind_len(id: *const rc_column_id) -> usize;
    pub fn rc_row_id_row_ind_len(id: *const rc_row_id) -> usize;
    pub fn rc_two_sided_id_row_ind_len(id: *const rc_two_sided_id) -> usize;
    pub fn rc_two_sided_id_col_ind_len(id: *const rc_two_sided_id) -> usize;
    pub fn rc_matrix_copy(ctx: *mut rc_ctx, src: *const rc_matrix, dst: *mut rc_matrix) -> c_int;
    pub fn rc_operator_create(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, matmat: rc_matmat_fn, conj_matmat: rc_matmat_fn, user: *mut c_void, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_helmholtz_kernel_matrix(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, seed: u64, kappa: f64, shift: f64, row_offset: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_tall_shard_matrix(ctx: *mut rc_ctx, dtype: c_int, rows: i64, cols: i64, r0: i64, decade_every: f64, seed: u64, row_offset: i64, m_total: i64, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_rel_diff_fro(ctx: *mut rc_ctx, first: *const rc_matrix, second: *const rc_matrix, out: *mut f64) -> c_int;
    pub fn rc_rel_diff_l2(ctx: *mut rc_ctx, first: *const rc_matrix, second: *const rc_matrix, out: *mut f64) -> c_int;
    pub fn rc_max_col_norm(ctx: *mut rc_ctx, m: *const rc_matrix, out: *mut f64) -> c_int;
    pub fn rc_invert_permutation_vector(perm: *const u64, n: usize, inverse: *mut u64) -> c_int;
    pub fn rc_apply_permutation_matrix(ctx: *mut rc_ctx, m: *const rc_matrix, index_array: *const u64, n: usize, mode: c_int, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_apply_permutation_vector(ctx: *mut rc_ctx, v: *const rc_matrix, index_array: *const u64, n: usize, mode: c_int, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_sample_range_by_rank(ctx: *mut rc_ctx, a: *const rc_matrix, k: i64, p: i64, omega: *const rc_matrix, seed: u64, q: *mut *mut rc_matrix) -> c_int;
    pub fn rc_sample_range_power_iteration(ctx: *mut rc_ctx, a: *const rc_matrix, k: i64, p: i64, it_count: i64, omega: *const rc_matrix, seed: u64, q: *mut *mut rc_matrix) -> c_int;
    pub fn rc_sample_range_adaptive(ctx: *mut rc_ctx, a: *const rc_matrix, rel_tol: f64, sample_size: i64, omega_blocks: *const rc_matrix, seed: u64, max_rank: i64, q: *mut *mut rc_matrix, hist_rank: *mut u64, hist_res: *mut f64, hist_cap: usize, hist_len: *mut usize) -> c_int;
    pub fn rc_qr_compute_from(ctx: *mut rc_ctx, arr: *const rc_matrix, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_qr_new(ctx: *mut rc_ctx, q: *const rc_matrix, r: *const rc_matrix, ind: *const u64, n: usize, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_qr_compute_from_range_estimate(ctx: *mut rc_ctx, range: *const rc_matrix, op: *const rc_matrix, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_qr_compress_rank(ctx: *mut rc_ctx, qr: *const rc_qr, max_rank: i64, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_qr_compress_tolerance(ctx: *mut rc_ctx, qr: *const rc_qr, tol: f64, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_qr_to_mat(ctx: *mut rc_ctx, qr: *const rc_qr, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_qr_column_id(ctx: *mut rc_ctx, qr: *const rc_qr, out: *mut *mut rc_column_id) -> c_int;
    pub fn rc_qr_get_q(qr: *const rc_qr) -> *const rc_matrix;
    pub fn rc_qr_get_r(qr: *const rc_qr) -> *const rc_matrix;
    pub fn rc_qr_rank(qr: *const rc_qr) -> i64;
    pub fn rc_qr_nrows(qr: *const rc_qr) -> i64;
    pub fn rc_qr_ncols(qr: *const rc_qr) -> i64;
    pub fn rc_qr_get_ind(qr: *const rc_qr, out: *mut u64, n: usize) -> c_int;
    pub fn rc_qr_free(qr: *mut rc_qr) -> c_int;
    pub fn rc_lq_compute_from(ctx: *mut rc_ctx, arr: *const rc_matrix, out: *mut *mut rc_lq) -> c_int;
    pub fn rc_lq_new(ctx: *mut rc_ctx, l: *const rc_matrix, q: *const rc_matrix, ind: *const u64, n: usize, out: *mut *mut rc_lq) -> c_int;
    pub fn rc_lq_compress_rank(ctx: *mut rc_ctx, lq: *const rc_lq, max_rank: i64, out: *mut *mut rc_lq) -> c_int;
    pub fn rc_lq_compress_tolerance(ctx: *mut rc_ctx, lq: *const rc_lq, tol: f64, out: *mut *mut rc_lq) -> c_int;
    pub fn rc_lq_to_mat(ctx: *mut rc_ctx, lq: *const rc_lq, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_lq_row_id(ctx: *mut rc_ctx, lq: *const rc_lq, out: *mut *mut rc_row_id) -> c_int;
    pub fn rc_lq_get_l(lq: *const rc_lq) -> *const rc_matrix;
    pub fn rc_lq_get_q(lq: *const rc_lq) -> *const rc_matrix;
    pub fn rc_lq_rank(lq: *const rc_lq) -> i64;
    pub fn rc_lq_nrows(lq: *const rc_lq) -> i64;
    pub fn rc_lq_ncols(lq: *const rc_lq) -> i64;
    pub fn rc_lq_get_ind(lq: *const rc_lq, out: *mut u64, n: usize) -> c_int;
    pub fn rc_lq_free(lq: *mut rc_lq) -> c_int;
    pub fn rc_svd_compute_from(ctx: *mut rc_ctx, arr: *const rc_matrix, out: *mut *mut rc_svd) -> c_int;
    pub fn rc_svd_new(ctx: *mut rc_ctx, u: *const rc_matrix, s: *const f64, ns: usize, vt: *const rc_matrix, out: *mut *mut rc_svd) -> c_int;
    pub fn rc_svd_compute_from_range_estimate(ctx: *mut rc_ctx, range: *const rc_matrix, op: *const rc_matrix, out: *mut *mut rc_svd) -> c_int;
    pub fn rc_svd_compress_rank(ctx: *mut rc_ctx, svd: *const rc_svd, max_rank: i64, out: *mut *mut rc_svd) -> c_int;
    pub fn rc_svd_compress_tolerance(ctx: *mut rc_ctx, svd: *const rc_svd, tol: f64, out: *mut *mut rc_svd) -> c_int;
    pub fn rc_svd_to_mat(ctx: *mut rc_ctx, svd: *const rc_svd, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_svd_to_qr(ctx: *mut rc_ctx, svd: *const rc_svd, out: *mut *mut rc_qr) -> c_int;
    pub fn rc_svd_get_u(svd: *const rc_svd) -> *const rc_matrix;
    pub fn rc_svd_get_vt(svd: *const rc_svd) -> *const rc_matrix;
    pub fn rc_svd_rank(svd: *const rc_svd) -> i64;
    pub fn rc_svd_get_s(svd: *const rc_svd, out: *mut f64, n: usize) -> c_int;
    pub fn rc_svd_free(svd: *mut rc_svd) -> c_int;
    pub fn rc_column_id_new(ctx: *mut rc_ctx, c: *const rc_matrix, z: *const rc_matrix, col_ind: *const u64, n: usize, out: *mut *mut rc_column_id) -> c_int;
    pub fn rc_column_id_get_c(id: *const rc_column_id) -> *const rc_matrix;
    pub fn rc_column_id_get_z(id: *const rc_column_id) -> *const rc_matrix;
    pub fn rc_column_id_get_col_ind(id: *const rc_column_id, out: *mut u64, n: usize) -> c_int;
    pub fn rc_column_id_to_mat(ctx: *mut rc_ctx, id: *const rc_column_id, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_column_id_apply(ctx: *mut rc_ctx, id: *const rc_column_id, rhs: *const rc_matrix, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_column_id_two_sided_id(ctx: *mut rc_ctx, id: *const rc_column_id, out: *mut *mut rc_two_sided_id) -> c_int;
    pub fn rc_column_id_free(id: *mut rc_column_id) -> c_int;
    pub fn rc_row_id_new(ctx: *mut rc_ctx, x: *const rc_matrix, r: *const rc_matrix, row_ind: *const u64, n: usize, out: *mut *mut rc_row_id) -> c_int;
    pub fn rc_row_id_get_x(id: *const rc_row_id) -> *const rc_matrix;
    pub fn rc_row_id_get_r(id: *const rc_row_id) -> *const rc_matrix;
    pub fn rc_row_id_get_row_ind(id: *const rc_row_id, out: *mut u64, n: usize) -> c_int;
    pub fn rc_row_id_to_mat(ctx: *mut rc_ctx, id: *const rc_row_id, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_row_id_apply(ctx: *mut rc_ctx, id: *const rc_row_id, rhs: *const rc_matrix, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_row_id_two_sided_id(ctx: *mut rc_ctx, id: *const rc_row_id, out: *mut *mut rc_two_sided_id) -> c_int;
    pub fn rc_row_id_free(id: *mut rc_row_id) -> c_int;
    pub fn rc_two_sided_id_new(ctx: *mut rc_ctx, x: *const rc_matrix, r: *const rc_matrix, c: *const rc_matrix, col_ind: *const u64, n_col: usize, row_ind: *const u64, n_row: usize, out: *mut *mut rc_two_sided_id) -> c_int;
    pub fn rc_two_sided_id_get_c(id: *const rc_two_sided_id) -> *const rc_matrix;
    pub fn rc_two_sided_id_get_x(id: *const rc_two_sided_id) -> *const rc_matrix;
    pub fn rc_two_sided_id_get_r(id: *const rc_two_sided_id) -> *const rc_matrix;
    pub fn rc_two_sided_id_get_row_ind(id: *const rc_two_sided_id, out: *mut u64, n: usize) -> c_int;
    pub fn rc_two_sided_id_get_col_ind(id: *const rc_two_sided_id, out: *mut u64, n: usize) -> c_int;
    pub fn rc_two_sided_id_to_mat(ctx: *mut rc_ctx, id: *const rc_two_sided_id, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_two_sided_id_apply(ctx: *mut rc_ctx, id: *const rc_two_sided_id, rhs: *const rc_matrix, out: *mut *mut rc_matrix) -> c_int;
    pub fn rc_two_sided_id_free(id: *mut rc_two_sided_id) -> c_int;
}
