/* rc_api.h -- C ABI of the B200-native randomized low-rank engine.
 *
 * Drop-in boundary for the hot path of the Rust crate rusty-compression
 * (reference tree: /root/reference, citations are "file:line" into it).  Every entry point
 * names the reference interface it replaces.  The library is built by nvcc for sm_100a from
 * rusty_compression_b200/csrc (hand-written kernels only: no cuBLAS, no cuSOLVER, no CPU
 * fallback) and exports exactly the symbols declared here (tests/test_abi_symbols.py).
 *
 * Conventions
 *   - Plain C: opaque handles, pointers and sizes.  No exceptions cross the ABI; every
 *     function returns an rc_status and rc_last_error_string() describes the last failure.
 *   - Scalars: the crate monomorphises over f32/f64/c32/c64 (e.g. src/qr.rs:408-416); here
 *     the scalar is the rc_dtype tag carried by every rc_matrix.  Complex values are
 *     interleaved (re, im) pairs of the real type, like num::Complex (#[repr(C)]).
 *   - Layout at the boundary: row-major with explicit element strides on the way in (ndarray
 *     views, src/pivoted_qr.rs:25-31), dense row-major on the way out (freshly owned arrays).
 *   - Index vectors are 0-based uint64_t (Rust usize, src/qr.rs:39,50) and always have full
 *     length (quirk Q8, src/qr.rs:182,306).
 *   - Ownership: inputs are borrowed; every result is a new handle the caller frees with the
 *     matching *_free.  The library never retains host pointers after a call returns.
 *   - Threading: calls on one rc_ctx are serialised by the caller; results are valid when a
 *     function returns (host outputs) or are ordered on the context's CUDA stream (handles).
 *   - Randomness: the crate takes `rng: &mut R` (src/random_sampling.rs:66-71).  Here Omega is
 *     either ingested (`omega != NULL`, the parity mode) or generated on device by
 *     Philox4x32-10 keyed by `seed` (counter = row-major element index, Box-Muller), so every
 *     row shard regenerates the same Omega.
 */
#ifndef RC_API_H
#define RC_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: RustyCompressionError (src/types.rs:11-21) plus the classes a device
 *      library adds.  Where the crate panics via assert! (src/qr.rs:99,188; src/svd.rs:88;
 *      src/permutation.rs:96-133,161-164; src/random_matrix.rs:78-82) the ABI returns
 *      RC_INVALID_ARGUMENT instead of aborting. */
typedef enum rc_status {
    RC_OK = 0,
    RC_LINALG_ERROR = 1,      /* RustyCompressionError::LinalgError      */
    RC_COMPRESSION_ERROR = 2, /* ::CompressionError (src/qr.rs:196-199)  */
    RC_LAYOUT_ERROR = 3,      /* ::LayoutError (src/pivoted_qr.rs:88-91) */
    RC_PIVOTED_QR_ERROR = 4,  /* ::PivotedQRError (src/pivoted_qr.rs:95) */
    RC_INVALID_ARGUMENT = 5,
    RC_CUDA_ERROR = 6,
    RC_NCCL_ERROR = 7,
    RC_OUT_OF_MEMORY = 8
} rc_status;

/* ---- scalar tags: f32, f64, c32, c64 (src/types.rs:9) */
typedef enum rc_dtype { RC_F32 = 0, RC_F64 = 1, RC_C32 = 2, RC_C64 = 3 } rc_dtype;

/* ---- MatrixPermutationMode / VectorPermutationMode (src/permutation.rs:7-24) */
typedef enum rc_perm_mode {
    RC_PERM_COL = 0, RC_PERM_ROW = 1, RC_PERM_COLINV = 2, RC_PERM_ROWINV = 3
} rc_perm_mode;
typedef enum rc_vperm_mode { RC_VPERM_INV = 0, RC_VPERM_NOINV = 1 } rc_vperm_mode;

typedef struct rc_ctx rc_ctx;           /* device, stream, workspaces, communicator          */
typedef struct rc_matrix rc_matrix;     /* device-resident dense matrix / operator (MatVec..) */
typedef struct rc_qr rc_qr;             /* QR{q,r,ind}  src/qr.rs:31-40                       */
typedef struct rc_lq rc_lq;             /* LQ{l,q,ind}  src/qr.rs:42-51                       */
typedef struct rc_svd rc_svd;           /* SVD{u,s,vt}  src/svd.rs:13-20                      */
typedef struct rc_column_id rc_column_id;   /* ColumnID{c,z,col_ind}  src/col_interp_decomp.rs:23-31 */
typedef struct rc_row_id rc_row_id;         /* RowID{x,r,row_ind}     src/row_interp_decomp.rs:25-33 */
typedef struct rc_two_sided_id rc_two_sided_id; /* TwoSidedID{c,x,r,row_ind,col_ind}
                                                   src/two_sided_interp_decomp.rs:19-30 */

/* ================================================================ context */
int rc_version(void);
rc_status rc_ctx_create(int device, rc_ctx** out);
rc_status rc_ctx_destroy(rc_ctx* ctx);
/* Use an existing CUDA stream (e.g. torch's current stream), passed as cudaStream_t. */
rc_status rc_ctx_set_stream(rc_ctx* ctx, void* cuda_stream);
rc_status rc_ctx_synchronize(rc_ctx* ctx);
const char* rc_last_error_string(rc_ctx* ctx);
/* Tuning / test knobs.  Keys: "gemm_impl" (0 auto, 1 generic SIMT tiles only),
 * "true_power_iteration" (0 = reference semantics incl. quirk Q1, 1 = textbook iteration),
 * "qr_mode" (0 = Cholesky-QR2 fast path for well-conditioned tall panels with automatic
 * fallback to Householder TSQR, 1 = Householder TSQR always),
 * "f32_precision" (0, the default = f32 / c32 contractions as a 3-product TF32 split on tcgen05, f32-accurate; 1 = the
 * caller opts in to ONE bf16 product per contraction, tcgen05 kind::f16 with FP32 accumulation: ~3e-3 relative),
 * "pivot_precision" (1, the default = pivot decisions on f32 / c32 inputs are taken in double precision -- the sequence
 * ?geqp3 picks in double on the same single-precision data; 0 = working precision),
 * "reuse_range_b" (1 = compute_from_range_estimate reuses the B = Q^H A the adaptive sampler built),
 * "dmma_tail" (1 = a ragged last 8-column group of the FP64 tensor-pipe GEMM is formed with DFMAs,
 * 0 = padded DMMA; same results up to summation order in those columns), "trace" (1 = stage timer
 * on stderr; synchronises at every mark),
 * "speculate" (1 = no host synchronisation inside the power-iteration sampler: the Cholesky-QR2 status words are checked
 * once at the end -- those of the first sketch right after its factorisation -- and a rejected panel is redone, or the
 * sampler re-run, on the next route: shifted Cholesky-QR, then Householder TSQR), "shifted_cholqr" (1 = double-precision
 * sketches the plain Cholesky-QR2 rejects try two shifted Cholesky rounds in front of it before the Householder TSQR;
 * 0 = Householder at once), "overlap" (1 = independent stages on auxiliary streams), "side_sms" (SMs left to the
 * small-kernel chain that runs beside a big product on another stream, default 8, 0 = no such overlap), "fused_small_qr" (1 = small pivoted QRs in the fused one-CTA kernel), "cluster_qr" (1 = medium
 * pivoted QRs -- the factor fits 8 CTAs' shared memory -- in the thread-block-cluster kernel, 0 = cooperative grid
 * kernel), "svd_precondition" (1 = the SVD of a matrix too wide for one CTA runs Jacobi on R^H of a pivoted QR; 0 = on the
 * unpivoted triangle), "tf32_ring" (A/B variants of the tcgen05 TF32 kernel: 0 default, 1 = six TMEM split stages, 2 = high part
 * from shared memory), "workspace_cache" (1 = device blocks of 1 MiB .. 8 GiB are cached per context and reused in
 * stream order; 0 = every allocation goes to cudaMallocAsync), "release_workspaces" (any value: empty the cache and trim
 * the device memory pool now).
 * A handle created through a context must be freed before rc_ctx_destroy: the free routines reach into the context. */
rc_status rc_ctx_set_option(rc_ctx* ctx, const char* key, int64_t value);
/* Counters: "kernel_launches" (own kernels launched so far), "gemm_flops", "h2d_bytes",
 * "d2h_bytes", "cholqr_used", "cholqr_fallbacks" (rejections of a Cholesky-QR attempt), "cholqr_shifted" (panels
 * factored on the shifted route), "range_b_reused", "workspace_cache_hits", "workspace_cache_misses",
 * "workspace_cached_bytes".  rc_ctx_reset_counters zeroes them. */
rc_status rc_ctx_get_counter(rc_ctx* ctx, const char* key, int64_t* out);
rc_status rc_ctx_reset_counters(rc_ctx* ctx);

/* Row-sharded multi-GPU (no reference counterpart: the crate is single-process).  One
 * process per GPU; rank g owns rows [row_offset, row_offset + local_rows) of every m-row
 * object; n-row and small objects are replicated.  Collectives run over NCCL on the
 * context stream.  The 128-byte unique id is created on rank 0 and handed to the other
 * ranks by the host (torch.distributed broadcast in bench.py). */
rc_status rc_comm_get_unique_id(void* out_id_128_bytes);
rc_status rc_ctx_comm_init(rc_ctx* ctx, const void* id_128_bytes, int rank, int nranks);
rc_status rc_ctx_comm_info(rc_ctx* ctx, int* rank, int* nranks);

/* ================================================================ matrices / operator plugin
 * rc_matrix is the device-resident stand-in for ndarray Array2 and, for A itself, the operator
 * that implements MatVec/MatMat/ConjMatVec/ConjMatMat (src/types.rs:40-101, 103-133). */
rc_status rc_matrix_create(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols, rc_matrix** out);
/* Upload a strided host view (strides in elements), like `mat.assign(&arr)` (src/pivoted_qr.rs:29). */
rc_status rc_matrix_from_host(rc_ctx* ctx, rc_dtype dtype, const void* host, int64_t rows,
                              int64_t cols, int64_t row_stride, int64_t col_stride, rc_matrix** out);
/* Pipelined upload of a row-major host view (row stride in elements, unit column stride; pinned memory for a
 * transfer that really is asynchronous): returns as soon as the copy is queued on the context's copy stream, so
 * the transfer of the next operator overlaps the kernels working on the current one.  The handle must go through
 * rc_matrix_await before ANY other use.  rc_matrix_await orders the context stream behind the copy (it does not
 * block the host unless `block_host` != 0); the host buffer must stay valid and unmodified until the copy is
 * complete (rc_matrix_await with block_host = 1, or any later rc_ctx_synchronize / blocking call on the context).
 * No counterpart in the reference (a host library has no transfer to hide). */
rc_status rc_matrix_from_host_async(rc_ctx* ctx, rc_dtype dtype, const void* host, int64_t rows,
                                    int64_t cols, int64_t row_stride, rc_matrix** out);
rc_status rc_matrix_await(rc_ctx* ctx, rc_matrix* m, int block_host);
/* Page-lock / unlock host memory the caller owns (e.g. the allocation behind an ndarray Array2, src/types.rs:58-71
 * takes such arrays by reference): uploads from registered memory run at the full rate of the link and the
 * asynchronous upload above does not block.  Wrappers of cudaHostRegister / cudaHostUnregister. */
rc_status rc_host_register(rc_ctx* ctx, void* host, size_t bytes);
rc_status rc_host_unregister(rc_ctx* ctx, void* host);
/* Borrow device memory (row-major, leading dimension `ld` elements); never freed by the library. */
rc_status rc_matrix_wrap_device(rc_ctx* ctx, rc_dtype dtype, void* device_ptr, int64_t rows,
                                int64_t cols, int64_t ld, rc_matrix** out);
/* dst <- src, device to device, same shape and scalar type; leading dimensions may differ (wrapped buffers). */
rc_status rc_matrix_copy(rc_ctx* ctx, const rc_matrix* src, rc_matrix* dst);
/* Matrix-free operators -- the crate's plugin API (`MatVec`, `ConjMatVec`, `MatMat`, `ConjMatMat` implemented
 * by the caller, src/types.rs:40-101; the samplers and compute_from_range_estimate are generic over them,
 * src/random_sampling.rs:102,130,222, src/qr.rs:221-224, src/svd.rs:110-113).  The callbacks compute
 * Y = A X (matmat; X cols x ncols, Y rows x ncols) and Z = A^H X (conj_matmat; X rows x ncols, Z cols x ncols)
 * on DEVICE buffers, row-major with the given leading dimensions (in elements), enqueued on `cuda_stream`
 * (the cudaStream_t the context works on); they return 0 on success.  The handle is accepted wherever an
 * operator is: rc_matmat, rc_conj_matmat, the three samplers, rc_qr/rc_svd_compute_from_range_estimate.
 * Entry points that need a dense matrix return RC_INVALID_ARGUMENT for it.  Free with rc_matrix_free. */
typedef int (*rc_matmat_fn)(void* user, const void* x, int64_t ldx, int64_t ncols, void* y, int64_t ldy,
                            void* cuda_stream);
rc_status rc_operator_create(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols, rc_matmat_fn matmat,
                             rc_matmat_fn conj_matmat, void* user, rc_matrix** out);
/* Download to a dense row-major host buffer of rows*cols elements. */
rc_status rc_matrix_to_host(rc_ctx* ctx, const rc_matrix* m, void* host);
/* Device-to-device copy into a dense row-major device buffer (ld = cols). */
rc_status rc_matrix_to_device(rc_ctx* ctx, const rc_matrix* m, void* device_ptr);
rc_status rc_matrix_free(rc_matrix* m);
int64_t rc_matrix_rows(const rc_matrix* m);
int64_t rc_matrix_cols(const rc_matrix* m);
int64_t rc_matrix_ld(const rc_matrix* m);
int rc_matrix_dtype(const rc_matrix* m);
void* rc_matrix_device_ptr(const rc_matrix* m);
/* Mark an m-row matrix as the local row block of a row-sharded global matrix
 * (global_rows total, this shard starts at row_offset). */
rc_status rc_matrix_set_shard(rc_matrix* m, int64_t global_rows, int64_t row_offset);

/* MatMat::matmat (src/types.rs:58-71): Y = A X.  One GEMM instead of the crate's per-column
 * GEMV loop (quirk Q2); same result up to summation order. */
rc_status rc_matmat(rc_ctx* ctx, const rc_matrix* a, const rc_matrix* x, rc_matrix** y);
/* ConjMatMat::conj_matmat (src/types.rs:88-101, 123-133): Z = A^H X.  For a row-sharded A the
 * partial products are summed across ranks (all-reduce). */
rc_status rc_conj_matmat(rc_ctx* ctx, const rc_matrix* a, const rc_matrix* x, rc_matrix** z);

/* RandomMatrix::random_gaussian (src/random_matrix.rs:21, 96-145), seeded.  Element (i, j) of a
 * matrix with `cols` columns uses Philox counter (row_offset + i) * cols + j, stream id `stream`. */
rc_status rc_random_gaussian(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols, uint64_t seed,
                             uint32_t stream, int64_t row_offset, rc_matrix** out);
/* RandomMatrix::random_orthogonal_matrix (src/random_matrix.rs:35-56). */
rc_status rc_random_orthogonal_matrix(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols,
                                      uint64_t seed, uint32_t stream, rc_matrix** out);
/* RandomMatrix::random_approximate_low_rank_matrix (src/random_matrix.rs:70-93). */
rc_status rc_random_approximate_low_rank_matrix(rc_ctx* ctx, rc_dtype dtype, int64_t rows,
                                                int64_t cols, double sigma_max, double sigma_min,
                                                uint64_t seed, rc_matrix** out);
/* Bench/test input of SURVEY.md 8(d): A = U diag(10^(-j/decade_every)) V^H with U, V
 * orthonormalised Philox Gaussians of rank r0, built entirely on device.  `row_offset` selects
 * which rows of the global Gaussian seed U, so that row shards of a taller matrix differ while V
 * (and hence the row space) is shared by all shards. */
rc_status rc_decaying_spectrum_matrix(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols,
                                      int64_t r0, double decade_every, uint64_t seed,
                                      int64_t row_offset, rc_matrix** out);
/* Bench input of BASELINE config 4: rows [row_offset, row_offset + rows) of
 * A = m_total^(-1/2) G diag(10^(-j/decade_every)) V^H with G_ij ~ N(0,1) keyed by (global row, j) and V a
 * shared orthonormal factor, generated on device so that every rank of a row-sharded run builds exactly
 * its own 2^23/P rows (the 256 GiB matrix never exists in one place). */
rc_status rc_tall_shard_matrix(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols, int64_t r0,
                               double decade_every, uint64_t seed, int64_t row_offset,
                               int64_t m_total, rc_matrix** out);
/* Bench/test input of BASELINE config 5: A_ij = exp(i kappa |x_i - y_j|) / |x_i - y_j| for x_i
 * uniform in [0,1]^3 and y_j uniform in [0,1]^3 + (shift, 0, 0) (Philox-seeded, so row shards regenerate their
 * own rows via `row_offset`); real scalar types take the real part.  Generated on device. */
rc_status rc_helmholtz_kernel_matrix(rc_ctx* ctx, rc_dtype dtype, int64_t rows, int64_t cols, uint64_t seed,
                                     double kappa, double shift, int64_t row_offset, rc_matrix** out);

/* RelDiff::{rel_diff_fro, rel_diff_l2} (src/types.rs:162-204): ||first - second|| / ||second||. */
rc_status rc_rel_diff_fro(rc_ctx* ctx, const rc_matrix* first, const rc_matrix* second, double* out);
rc_status rc_rel_diff_l2(rc_ctx* ctx, const rc_matrix* first, const rc_matrix* second, double* out);
/* MaxColNorm::max_col_norm (src/random_sampling.rs:175-199). */
rc_status rc_max_col_norm(rc_ctx* ctx, const rc_matrix* m, double* out);

/* ================================================================ permutations
 * src/permutation.rs:28-38, 77-145, 147-184. */
rc_status rc_invert_permutation_vector(const uint64_t* perm, size_t n, uint64_t* inverse);
rc_status rc_apply_permutation_matrix(rc_ctx* ctx, const rc_matrix* m, const uint64_t* index_array,
                                      size_t n, rc_perm_mode mode, rc_matrix** out);
/* `v` is a 1 x n or n x 1 matrix. */
rc_status rc_apply_permutation_vector(rc_ctx* ctx, const rc_matrix* v, const uint64_t* index_array,
                                      size_t n, rc_vperm_mode mode, rc_matrix** out);

/* ================================================================ randomized range finders */
/* SampleRange::sample_range_by_rank (src/random_sampling.rs:58-72, 100-126).
 * omega: NULL (Philox from `seed`, stream 0) or an n x (k+p) matrix to ingest.  q: m x min(k, k'). */
rc_status rc_sample_range_by_rank(rc_ctx* ctx, const rc_matrix* a, int64_t k, int64_t p,
                                  const rc_matrix* omega, uint64_t seed, rc_matrix** q);
/* SampleRangePowerIteration::sample_range_power_iteration (src/random_sampling.rs:82-98, 128-168),
 * including quirk Q1 (every trip restarts from A*Omega, :144-154). */
rc_status rc_sample_range_power_iteration(rc_ctx* ctx, const rc_matrix* a, int64_t k, int64_t p,
                                          int64_t it_count, const rc_matrix* omega, uint64_t seed,
                                          rc_matrix** q);
/* AdaptiveSampling::sample_range_adaptive (src/random_sampling.rs:202-218, 220-282).
 * omega_blocks: NULL (Philox, block b = stream b) or an n x (B*sample_size) matrix whose column
 * blocks are the draws in order (RC_INVALID_ARGUMENT when exhausted).  The convergence history
 * Vec<(rank, rel_res)> is returned through hist_rank/hist_res (capacity hist_cap, length
 * *hist_len).  max_rank (0 = min(m, n)) bounds the loop the crate leaves unbounded (quirk Q6);
 * exceeding it returns RC_COMPRESSION_ERROR. */
rc_status rc_sample_range_adaptive(rc_ctx* ctx, const rc_matrix* a, double rel_tol,
                                   int64_t sample_size, const rc_matrix* omega_blocks, uint64_t seed,
                                   int64_t max_rank, rc_matrix** q, uint64_t* hist_rank,
                                   double* hist_res, size_t hist_cap, size_t* hist_len);

/* ================================================================ QR / LQ
 * PivotedQR::pivoted_qr (src/pivoted_qr.rs:11-31, 81-184) under QRTraits::compute_from
 * (src/qr.rs:214, 251-253): A P = Q R, q m x k', r k' x n upper trapezoidal, k' = min(m, n). */
rc_status rc_qr_compute_from(rc_ctx* ctx, const rc_matrix* arr, rc_qr** out);
/* QR{q, r, ind} built from parts: the struct has pub fields (src/qr.rs:31-40), so a caller of the crate can assemble
 * one and call column_id() / compress() / to_mat() on it.  q is m x k, r is k x n, ind has length n; the matrices
 * are copied. */
rc_status rc_qr_new(rc_ctx* ctx, const rc_matrix* q, const rc_matrix* r, const uint64_t* ind, size_t n,
                    rc_qr** out);
/* QRTraits::compute_from_range_estimate (src/qr.rs:221-224, 311-323). */
rc_status rc_qr_compute_from_range_estimate(rc_ctx* ctx, const rc_matrix* range, const rc_matrix* op,
                                            rc_qr** out);
/* compress(CompressionType::RANK / ADAPTIVE) (src/lib.rs:82-87; src/qr.rs:169-208). */
rc_status rc_qr_compress_rank(rc_ctx* ctx, const rc_qr* qr, int64_t max_rank, rc_qr** out);
rc_status rc_qr_compress_tolerance(rc_ctx* ctx, const rc_qr* qr, double tol, rc_qr** out);
rc_status rc_qr_to_mat(rc_ctx* ctx, const rc_qr* qr, rc_matrix** out);          /* src/qr.rs:160-166 */
rc_status rc_qr_column_id(rc_ctx* ctx, const rc_qr* qr, rc_column_id** out);    /* src/qr.rs:270-309 */
const rc_matrix* rc_qr_get_q(const rc_qr* qr);
const rc_matrix* rc_qr_get_r(const rc_qr* qr);
int64_t rc_qr_rank(const rc_qr* qr);
int64_t rc_qr_nrows(const rc_qr* qr);
int64_t rc_qr_ncols(const rc_qr* qr);
rc_status rc_qr_get_ind(const rc_qr* qr, uint64_t* out, size_t n);
rc_status rc_qr_free(rc_qr* qr);

/* PivotedQR::pivoted_lq (src/pivoted_qr.rs:32-41), LQTraits (src/qr.rs:54-139, 326-405). */
rc_status rc_lq_compute_from(rc_ctx* ctx, const rc_matrix* arr, rc_lq** out);
/* LQ{l, q, ind} from parts (pub fields, src/qr.rs:42-51): l is m x k, q is k x n, ind has length m. */
rc_status rc_lq_new(rc_ctx* ctx, const rc_matrix* l, const rc_matrix* q, const uint64_t* ind, size_t n,
                    rc_lq** out);
rc_status rc_lq_compress_rank(rc_ctx* ctx, const rc_lq* lq, int64_t max_rank, rc_lq** out);
rc_status rc_lq_compress_tolerance(rc_ctx* ctx, const rc_lq* lq, double tol, rc_lq** out);
rc_status rc_lq_to_mat(rc_ctx* ctx, const rc_lq* lq, rc_matrix** out);          /* src/qr.rs:73-78 */
rc_status rc_lq_row_id(rc_ctx* ctx, const rc_lq* lq, rc_row_id** out);          /* src/qr.rs:363-403 */
const rc_matrix* rc_lq_get_l(const rc_lq* lq);
const rc_matrix* rc_lq_get_q(const rc_lq* lq);
int64_t rc_lq_rank(const rc_lq* lq);
int64_t rc_lq_nrows(const rc_lq* lq);
int64_t rc_lq_ncols(const rc_lq* lq);
rc_status rc_lq_get_ind(const rc_lq* lq, uint64_t* out, size_t n);
rc_status rc_lq_free(rc_lq* lq);

/* ================================================================ SVD
 * ComputeSVD::compute_svd (src/compute_svd.rs:8-35) under SVDTraits::compute_from (src/svd.rs:103). */
rc_status rc_svd_compute_from(rc_ctx* ctx, const rc_matrix* arr, rc_svd** out);
/* SVD{u, s, vt} from parts (pub fields, src/svd.rs:13-20): u is m x k, s has k real entries (passed as double,
 * A::Real widened), vt is k x n. */
rc_status rc_svd_new(rc_ctx* ctx, const rc_matrix* u, const double* s, size_t ns, const rc_matrix* vt,
                     rc_svd** out);
/* SVDTraits::compute_from_range_estimate (src/svd.rs:110-113, 171-183). */
rc_status rc_svd_compute_from_range_estimate(rc_ctx* ctx, const rc_matrix* range, const rc_matrix* op,
                                             rc_svd** out);
rc_status rc_svd_compress_rank(rc_ctx* ctx, const rc_svd* svd, int64_t max_rank, rc_svd** out);   /* src/svd.rs:68-84 */
rc_status rc_svd_compress_tolerance(rc_ctx* ctx, const rc_svd* svd, double tol, rc_svd** out);    /* src/svd.rs:87-101 */
rc_status rc_svd_to_mat(rc_ctx* ctx, const rc_svd* svd, rc_matrix** out);      /* src/svd.rs:42-54 */
rc_status rc_svd_to_qr(rc_ctx* ctx, const rc_svd* svd, rc_qr** out);           /* src/svd.rs:150-163 */
const rc_matrix* rc_svd_get_u(const rc_svd* svd);
const rc_matrix* rc_svd_get_vt(const rc_svd* svd);
int64_t rc_svd_rank(const rc_svd* svd);
/* Singular values, descending, widened to double (A::Real in the crate). */
rc_status rc_svd_get_s(const rc_svd* svd, double* out, size_t n);
rc_status rc_svd_free(rc_svd* svd);

/* ================================================================ interpolative decompositions
 * ColumnID / ColumnIDTraits / Apply (src/col_interp_decomp.rs:23-31, 44-86, 88-156). */
rc_status rc_column_id_new(rc_ctx* ctx, const rc_matrix* c, const rc_matrix* z,
                           const uint64_t* col_ind, size_t n, rc_column_id** out);
const rc_matrix* rc_column_id_get_c(const rc_column_id* id);
const rc_matrix* rc_column_id_get_z(const rc_column_id* id);
rc_status rc_column_id_get_col_ind(const rc_column_id* id, uint64_t* out, size_t n);
rc_status rc_column_id_to_mat(rc_ctx* ctx, const rc_column_id* id, rc_matrix** out);
rc_status rc_column_id_apply(rc_ctx* ctx, const rc_column_id* id, const rc_matrix* rhs, rc_matrix** out);
rc_status rc_column_id_two_sided_id(rc_ctx* ctx, const rc_column_id* id, rc_two_sided_id** out); /* :116-125 */
rc_status rc_column_id_free(rc_column_id* id);

/* RowID / RowIDTraits / Apply (src/row_interp_decomp.rs:25-33, 46-89, 91-156). */
rc_status rc_row_id_new(rc_ctx* ctx, const rc_matrix* x, const rc_matrix* r,
                        const uint64_t* row_ind, size_t n, rc_row_id** out);
const rc_matrix* rc_row_id_get_x(const rc_row_id* id);
const rc_matrix* rc_row_id_get_r(const rc_row_id* id);
rc_status rc_row_id_get_row_ind(const rc_row_id* id, uint64_t* out, size_t n);
rc_status rc_row_id_to_mat(rc_ctx* ctx, const rc_row_id* id, rc_matrix** out);
rc_status rc_row_id_apply(rc_ctx* ctx, const rc_row_id* id, const rc_matrix* rhs, rc_matrix** out);
rc_status rc_row_id_two_sided_id(rc_ctx* ctx, const rc_row_id* id, rc_two_sided_id** out);       /* :120-130 */
rc_status rc_row_id_free(rc_row_id* id);

/* TwoSidedID / TwoSidedIDTraits / Apply (src/two_sided_interp_decomp.rs:19-30, 43-96, 98-173).
 * Argument order of `new` follows the crate: (x, r, c, col_ind, row_ind) (:89-95). */
rc_status rc_two_sided_id_new(rc_ctx* ctx, const rc_matrix* x, const rc_matrix* r, const rc_matrix* c,
                              const uint64_t* col_ind, size_t n_col, const uint64_t* row_ind,
                              size_t n_row, rc_two_sided_id** out);
const rc_matrix* rc_two_sided_id_get_c(const rc_two_sided_id* id);
const rc_matrix* rc_two_sided_id_get_x(const rc_two_sided_id* id);
const rc_matrix* rc_two_sided_id_get_r(const rc_two_sided_id* id);
rc_status rc_two_sided_id_get_row_ind(const rc_two_sided_id* id, uint64_t* out, size_t n);
rc_status rc_two_sided_id_get_col_ind(const rc_two_sided_id* id, uint64_t* out, size_t n);
rc_status rc_two_sided_id_to_mat(rc_ctx* ctx, const rc_two_sided_id* id, rc_matrix** out);
rc_status rc_two_sided_id_apply(rc_ctx* ctx, const rc_two_sided_id* id, const rc_matrix* rhs,
                                rc_matrix** out);
rc_status rc_two_sided_id_free(rc_two_sided_id* id);
/* Lengths of the index vectors (always the full n or m, quirk Q8: src/qr.rs:182, 306).  With row-sharded
 * factors the local row count of c / x is not the length of row_ind. */
size_t rc_column_id_col_ind_len(const rc_column_id* id);
size_t rc_row_id_row_ind_len(const rc_row_id* id);
size_t rc_two_sided_id_row_ind_len(const rc_two_sided_id* id);
size_t rc_two_sided_id_col_ind_len(const rc_two_sided_id* id);

#ifdef __cplusplus
}
#endif
#endif /* RC_API_H */
