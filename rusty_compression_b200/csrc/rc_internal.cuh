// Internal declarations shared by the kernels and the host orchestration.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <stdexcept>
#include <map>
#include <unordered_map>

#include "../../include/rc_api.h"
#include "rc_scalar.cuh"

struct RcError {
    rc_status status;
    std::string msg;
};

#define RC_THROW(st, ...)                                                     \
    do {                                                                      \
        char _buf[512];                                                       \
        snprintf(_buf, sizeof(_buf), __VA_ARGS__);                            \
        throw RcError{st, std::string(_buf)};                                 \
    } while (0)

#define RC_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            rc_status _st = (_e == cudaErrorMemoryAllocation) ? RC_OUT_OF_MEMORY : RC_CUDA_ERROR; \
            RC_THROW(_st, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                     cudaGetErrorString(_e));                                              \
        }                                                                                  \
    } while (0)

#define RC_REQUIRE(cond, ...)                                    \
    do {                                                         \
        if (!(cond)) RC_THROW(RC_INVALID_ARGUMENT, __VA_ARGS__); \
    } while (0)

struct rc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int sm_count = 148;
    size_t smem_optin = 0;
    // options
    int gemm_impl = 0;            // 0 auto, 1 generic only
    int f32_precision = 0;        // option "f32_precision": 0 = 3xTF32 split (f32-accurate, default), 1 = bf16 single product (opt-in)
    int tf32_ring = 0;            // option "tf32_ring": 1 = deeper TMEM split ring in the tcgen05 TF32 kernel (A/B tuning)
    int dmma_tail = 1;            // 1: ragged last column group of the DMMA GEMM on the DFMA tail path (gemm_dmma.cu)
    int true_power_iteration = 0;
    int trace = 0;                // option "trace": print wall time between rc_trace() marks (stream-synchronising)
    double trace_t0 = 0.0;
    int reuse_range_b = 1;        // reuse B = Q^H A of the adaptive sampler in compute_from_range_estimate
    int pivot_f64 = 1;            // 1: pivot decisions on f32 / c32 inputs are taken in double (pivqr.cu, pivoted_qr_impl)
    int qr_mode = 0;              // 0 auto (Cholesky-QR2 fast path with Householder-TSQR fallback), 1 TSQR only
    int64_t cholqr_used = 0, cholqr_fallbacks = 0, range_b_reused = 0;
    // Speculative execution (host_api.cu, DeferScope): inside a deferred region the Cholesky-QR2 panels do not read
    // their status words back (no host synchronisation in the middle of a pipeline); the words are collected here and
    // checked once at the end, and a failed check re-runs the region with the Householder TSQR forced.
    int svd_precondition = 1;     // option "svd_precondition": wide SVDs run Jacobi on R^H of a pivoted QR (host_api.cu: svd_tall)
    int cluster_qr = 1;           // option "cluster_qr": medium pivoted QRs in the thread-block-cluster kernel (pivqr.cu)
    int fused_small_qr = 1;       // option "fused_small_qr": small pivoted QRs in the fused one-CTA kernel (pivqr.cu)
    int speculate = 1;            // option "speculate"
    int overlap = 1;              // option "overlap": independent stages on auxiliary streams
    int side_sms = 8;             // option "side_sms": SMs left to the small-kernel chain that runs beside a big product (SmBudget)
    int defer_depth = 0;
    bool force_householder = false;
    bool force_shifted = false;   // re-run of a speculative region on shifted Cholesky-QR3 (host_api.cu: cholqr2)
    int shifted_cholqr = 1;       // option "shifted_cholqr": ill-conditioned double-precision sketches try a shifted first round
                                  // before the Householder TSQR
    int64_t cholqr_shifted = 0;   // counter "cholqr_shifted"
    struct DeferredCheck { double* status; bool single; bool shifted; };
    std::vector<DeferredCheck> deferred;
    // auxiliary streams for independent stages (the two power-iteration trips of the reference are independent of
    // each other, quirk Q1), each with its own scratch for the dynamic GEMM tile scheduler
    cudaStream_t aux_stream[2] = {nullptr, nullptr};
    int* aux_tile_counter[2] = {nullptr, nullptr};
    cudaEvent_t aux_event[4] = {nullptr, nullptr, nullptr, nullptr};
    // upload stream of rc_matrix_from_host_async (created on first use): the next operator's host-to-device copy runs
    // here while the context stream computes on the current one
    cudaStream_t copy_stream = nullptr;
    // counters
    int64_t launches = 0, gemm_flops = 0, h2d_bytes = 0, d2h_bytes = 0;
    int* tile_counter = nullptr;   // device scratch for the dynamic GEMM tile scheduler
    // communicator (row-sharded multi-GPU)
    void* comm = nullptr;
    int rank = 0, nranks = 1;
    // workspace cache (rc_dev_alloc / rc_dev_free below): freed device blocks of >= 1 MiB, per stream, by size
    struct CachedBlock { void* p; cudaStream_t freed_on; cudaEvent_t freed_at; };
    std::unordered_map<void*, size_t> live_blocks;                                   // cacheable blocks handed out
    std::multimap<size_t, CachedBlock> block_cache;                                  // size -> free blocks
    std::vector<cudaEvent_t> event_pool;
    size_t cached_bytes = 0;
    int64_t cache_hits = 0, cache_misses = 0;
    int block_cache_on = 1;       // option "workspace_cache"
};

struct rc_matrix {
    rc_ctx* ctx = nullptr;
    int dtype = RC_F64;
    int64_t rows = 0, cols = 0, ld = 0;
    void* data = nullptr;
    bool owns = false;
    // row sharding: this handle holds rows [row_offset, row_offset + rows) of a
    // global_rows x cols matrix (global_rows == 0: not sharded)
    int64_t global_rows = 0, row_offset = 0;
    // identity of this buffer (unique per allocation) and, for a range estimate Q produced by the
    // adaptive sampler, the factor B = Q^H A it already computed for the operator `companion_op_id`
    // (QR/SVD::compute_from_range_estimate reuse it instead of another pass over A)
    uint64_t id = 0;
    // rc_matrix_from_host_async: recorded behind the upload on the context's copy stream; rc_matrix_await consumes it
    cudaEvent_t upload_done = nullptr;
    rc_matrix* companion = nullptr;
    uint64_t companion_op_id = 0;
    // matrix-free operator (rc_operator_create): no data, products go through the caller's callbacks
    int (*op_matmat)(void*, const void*, int64_t, int64_t, void*, int64_t, void*) = nullptr;
    int (*op_conj_matmat)(void*, const void*, int64_t, int64_t, void*, int64_t, void*) = nullptr;
    void* op_user = nullptr;
};

// Every entry point runs with the context's device current and restores the caller's device on exit: a host
// process with several contexts (one per GPU) or a framework that switches devices between calls (torch) must
// not see kernels, stream operations or cudaFuncSetAttribute land on the wrong device.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

inline size_t rc_dtype_size(int dt) {
    static const size_t s[4] = {4, 8, 8, 16};
    return s[dt];
}
inline int rc_real_dtype(int dt) { return dt & 1; }

// ---------------------------------------------------------------------------------------------------------------
// Device workspaces.  Every pipeline allocates and frees dozens of buffers per call, many of them GBs (the tall shards
// of config 4: Y, Q, projections, split operand copies).  Straight cudaMallocAsync / cudaFreeAsync left that to the
// driver's stream-ordered pool, whose reuse of large, differently sized, short-lived blocks only settled after five or
// six identical passes (measured at 2^22 x 8192 f32, 137 GB resident: 6.5 s, 1.7 s, 2.1 s, 0.49 s, 0.65 s, 0.25 s per
// pass; 40-400 ms at 2^20 rows) -- the kernels themselves took 0.20 s.  So blocks of 1 MiB ... 8 GiB are cached here: a
// freed block keeps an event recorded on the stream it was freed on and is handed out again to the next request it fits
// without wasting more than a quarter of it; a request from ANOTHER stream takes it only once that event has completed
// (CUDA's own rule: memory freed on a stream may be reused by work ordered after the free).  Every pass after the first
// is allocation-free.  Small blocks and cache misses go to cudaMallocAsync; when that runs out of memory the cache is
// released and the request retried.
void rc_cache_release(rc_ctx* c);        // hand every cached block back to the driver (host_api.cu)
inline void* rc_dev_alloc(rc_ctx* c, size_t bytes) {
    if (bytes == 0) return nullptr;
    constexpr size_t kMin = (size_t)1 << 20, kMax = (size_t)8 << 30;     // (operators of tens of GB are not worth caching)
    void* p = nullptr;
    const bool cacheable = c->block_cache_on && bytes >= kMin && bytes <= kMax;
    if (cacheable) {
        // best fit, preferring a block freed on this stream; a block freed on ANOTHER stream is taken only when the work
        // that preceded its free has completed (cudaEventQuery) -- waiting for the event instead would order this stream
        // behind whatever the other stream is running, e.g. the chain of small kernels behind the big product that runs
        // beside it (SmBudget, host_api.cu), which is exactly the overlap the two streams exist for
        auto it = c->block_cache.end();
        {
            int looked = 0;
            for (auto j = c->block_cache.lower_bound(bytes); j != c->block_cache.end() && j->first <= bytes + bytes / 4 && looked < 16; ++j, ++looked) {
                if (j->second.freed_on == c->stream) { it = j; break; }
                if (it == c->block_cache.end()) {
                    const cudaError_t q = cudaEventQuery(j->second.freed_at);
                    if (q == cudaSuccess) it = j;
                    else if (q == cudaErrorNotReady) (void)cudaGetLastError();   // (recorded as the last error; anything else stays)
                }
            }
        }
        if (it != c->block_cache.end()) {
            const rc_ctx::CachedBlock blk = it->second;
            if (blk.freed_on != c->stream) cudaStreamWaitEvent(c->stream, blk.freed_at, 0);
            c->event_pool.push_back(blk.freed_at);
            c->live_blocks[blk.p] = it->first;
            c->cached_bytes -= it->first;
            c->block_cache.erase(it);
            c->cache_hits++;
            return blk.p;
        }
        c->cache_misses++;
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, c->stream);
    if (e == cudaErrorMemoryAllocation && c->cached_bytes > 0) {
        cudaGetLastError();
        rc_cache_release(c);
        cudaStreamSynchronize(c->stream);           // (frees on the auxiliary streams become visible to this one)
        for (int i = 0; i < 2; ++i) if (c->aux_stream[i]) cudaStreamSynchronize(c->aux_stream[i]);
        e = cudaMallocAsync(&p, bytes, c->stream);
    }
    if (e != cudaSuccess) {
        rc_status st = (e == cudaErrorMemoryAllocation) ? RC_OUT_OF_MEMORY : RC_CUDA_ERROR;
        RC_THROW(st, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    if (cacheable) c->live_blocks[p] = bytes;
    return p;
}
inline void rc_dev_free(rc_ctx* c, void* p) {
    if (!p) return;
    auto it = c->live_blocks.find(p);
    if (it == c->live_blocks.end()) { cudaFreeAsync(p, c->stream); return; }
    const size_t bytes = it->second;
    c->live_blocks.erase(it);
    if (!c->block_cache_on) { cudaFreeAsync(p, c->stream); return; }
    cudaEvent_t ev = nullptr;
    if (!c->event_pool.empty()) { ev = c->event_pool.back(); c->event_pool.pop_back(); }
    else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); cudaFreeAsync(p, c->stream); return; }
    cudaEventRecord(ev, c->stream);
    c->block_cache.emplace(bytes, rc_ctx::CachedBlock{p, c->stream, ev});
    c->cached_bytes += bytes;
}

// Stream-ordered device buffer.
template <class T>
struct DevBuf {
    T* p = nullptr;
    rc_ctx* c = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(rc_ctx* ctx, size_t count) { alloc(ctx, count); }
    void alloc(rc_ctx* ctx, size_t count) {
        release();
        c = ctx;
        n = count;
        if (count) p = static_cast<T*>(rc_dev_alloc(ctx, count * sizeof(T)));
    }
    void release() {
        if (p) rc_dev_free(c, p);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    T* take() { T* r = p; p = nullptr; return r; }
};

#define RC_COUNT_LAUNCH(ctx) ((ctx)->launches++)
#define RC_CHECK_LAUNCH(ctx)                  \
    do {                                      \
        RC_COUNT_LAUNCH(ctx);                 \
        RC_CUDA(cudaGetLastError());          \
    } while (0)

enum RcOp { RC_OP_N = 0, RC_OP_T = 1, RC_OP_H = 2 };

// ------------------------------------------------------------------ kernels_basic.cu
template <class T> void k_fill(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld, T v);
template <class T> void k_eye(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld);
// dst (rows x cols, ldd) = src (rows x cols, lds)
template <class T> void k_copy(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols);
// dst (cols x rows) = op(src (rows x cols)), op = T or H
template <class T> void k_transpose(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, bool conj);
template <class T> void k_conj_inplace(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld);
// strided (element strides) -> dense row-major
template <class T> void k_strided_to_dense(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t rs, int64_t cs, int64_t rows, int64_t cols);
// out[:, j] = in[:, idx[j]]  (cols) ; out[i, :] = in[idx[i], :] (rows)
template <class T> void k_gather_cols(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx);
template <class T> void k_gather_rows(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx);
// zero strictly-lower part of a rows x cols matrix
template <class T> void k_triu(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld);
// rows scaled by real s[i]
template <class T> void k_scale_rows(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld, const RealOf<T>* s);
// dst = a - b
template <class T> void k_sub(rc_ctx*, T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols);
template <class T> void k_add(rc_ctx*, T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols);
// Philox Gaussian fill (row-major, element index = (row_offset + i) * cols + j)
template <class T> void k_gaussian(rc_ctx*, T* p, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, uint32_t stream, int64_t row_offset);
// dst[:, j] = factor * src[:, j] for the flagged columns j
template <class T> void k_replace_flagged_columns(rc_ctx*, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols,
                                                  const int* flags, double factor);
template <class T> void k_helmholtz(rc_ctx*, T* a, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, double kappa, double shift, int64_t row_offset);
// squared column norms (double), one per column: out[j] = sum_i |a_ij|^2
template <class T> void k_col_norms2(rc_ctx*, const T* a, int64_t lda, int64_t rows, int64_t cols, double* out);
// sum of |a_ij|^2 ; and sum of |a_ij - b_ij|^2 (double, device scalars)
template <class T> void k_fro2(rc_ctx*, const T* a, int64_t lda, int64_t rows, int64_t cols, double* out);
template <class T> void k_diff_fro2(rc_ctx*, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols, double* out);
// dst = (D) src, elementwise (f32 <-> f64, c32 <-> c64)
template <class D, class S> void k_cast(rc_ctx*, D* dst, int64_t ldd, const S* src, int64_t lds, int64_t rows, int64_t cols);
// complex real-expansion helpers for the DMMA path (c64 only)
void k_expand_rhs_c64(rc_ctx*, double* dst, int64_t ldd, const c64* x, int64_t ldx, int64_t rows, int64_t cols);
// real <-> T conversions of device vectors (singular values)
template <class T> void k_convert_real(rc_ctx*, RealOf<T>* dst, const double* src, int64_t n);

// ------------------------------------------------------------------ gemm_generic.cu / gemm_dmma.cu
// C (M x N, ldc) = alpha * op(A) * op(B) + beta * C ; all row-major.
template <class T>
void gemm(rc_ctx*, RcOp opa, RcOp opb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda,
          const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta);
template <class T>
void gemm_generic(rc_ctx*, RcOp opa, RcOp opb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda,
                  const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta);
// f64 tensor-pipe (DMMA) + TMA kernels; return false if the shape/alignment is not supported.
bool gemm_dmma_f64(rc_ctx*, bool a_transposed, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double* C, int64_t ldc);
// f32 Y = A X on tcgen05 (kind::tf32, 3-product split, TMEM accumulators); false if unsupported.
bool gemm_tf32x3_f32(rc_ctx*, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                     const float* X, int64_t ldx, float* Y, int64_t ldy);
bool gemm_tf32x3_f32_tn(rc_ctx*, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                        const float* Y, int64_t ldy, float* Z, int64_t ldz);
bool gemm_tf32x3_c32(rc_ctx*, bool a_conj_transposed, int64_t M, int64_t N, int64_t K, const c32* A, int64_t lda,
                     const c32* B, int64_t ldb, c32* C, int64_t ldc);
bool gemm_dmma_c64(rc_ctx*, bool a_conj_transposed, int64_t M, int64_t N, int64_t K, const c64* A, int64_t lda,
                   const c64* B, int64_t ldb, c64* C, int64_t ldc);

// ------------------------------------------------------------------ tsqr.cu
// Unpivoted Householder TSQR of a tall matrix Y (m x w, row-major ld), destroying Y (it is
// overwritten by the leaf Householder vectors).  Keeps the reflector tree so Q can be applied.
template <class T>
struct TsqrFactor {
    rc_ctx* ctx = nullptr;
    int64_t m = 0, w = 0;
    struct Level {
        T* v = nullptr;        // (rows x w) reflectors, row-major, ld = ldv
        int64_t ldv = 0;
        int64_t rows = 0;      // rows of this level's input
        int64_t block = 0;     // rows per block
        int64_t nblocks = 0;
        T* tau = nullptr;      // nblocks x w
        T* tpan = nullptr;     // nblocks x npanels x (8 x 8) compact-WY T factors
        bool owns_v = false;
        bool tri = false;      // input is a stack of upper triangles (tree levels)
    };
    std::vector<Level> levels;
    T* r = nullptr;            // final w x w upper-triangular R (row-major, ld = w)
    ~TsqrFactor();
};
template <class T> void tsqr_factor(rc_ctx*, T* y, int64_t ld, int64_t m, int64_t w, TsqrFactor<T>& f);
// out (m x nc, ldo) = Q * [ctop (w x nc, ldc) ; 0]
template <class T> void tsqr_apply_q(rc_ctx*, const TsqrFactor<T>& f, const T* ctop, int64_t ldc, int64_t nc, T* out, int64_t ldo);
int64_t tsqr_max_width(rc_ctx*, int dtype);

// ------------------------------------------------------------------ pivqr.cu
// Column-pivoted Householder QR (LAPACK ?geqp3 / ?laqp2 pivot rule) of a p x n matrix held
// COLUMN-major in wc (ld = ldw >= p), in place.  Outputs: r (kk x n row-major, ldr), ind (n),
// and the reflectors (vbuf p x kk column-major ld = p, tau kk) for forming Q.  kk = min(p, n).
template <class T>
void pivqr_factor(rc_ctx*, T* wc, int64_t ldw, int64_t p, int64_t n, T* r, int64_t ldr, int* ind,
                  T* vbuf, T* tau);
// Fused small-factor route: pivoted QR of M (p x n) with R (kk x n row-major) and Q[:, :ncq] (p x ncq row-major) out of
// one kernel.  `in`: mode 0 = M row-major (ldin), 1 = M column-major, 2 = S (n x p row-major) with M = S^H.
// Returns false when the shape does not fit (caller uses pivqr_factor + pivqr_form_q).
template <class T>
bool pivqr_fused(rc_ctx*, const T* in, int64_t ldin, int mode, int64_t p, int64_t n, int64_t ncq,
                 T* r, int64_t ldr, int* ind, T* q, int64_t ldq);
// q (p x nc row-major, ldq) = H_0 ... H_{kk-1} * I[:, :nc]
template <class T>
void pivqr_form_q(rc_ctx*, const T* vbuf, const T* tau, int64_t p, int64_t kk, int64_t nc, T* q, int64_t ldq);

// ------------------------------------------------------------------ jacobi.cu
// One-sided Jacobi SVD of a small rows x n matrix G (row-major, ld; rows >= n): G = U diag(s) W^H.
// On exit u (rows x n row-major), s (n, descending, double), w (n x n row-major).
// info_dev (optional, 2 ints on device): {1 if the sweep limit was reached without convergence, sweeps used}.
template <class T>
void jacobi_svd(rc_ctx*, const T* g, int64_t ldg, int64_t rows, int64_t n, T* u, int64_t ldu, double* s, T* w, int64_t ldw,
                int* info_dev = nullptr);

// ------------------------------------------------------------------ trsm.cu
// X (k x nrhs, ldx) = U^{-1} B (k x nrhs, ldb), U k x k upper triangular (row-major ldu);
// if u_transposed, U[i][j] is read from u[j*ldu + i] (i.e. the stored matrix is lower).
template <class T>
void trsm_upper(rc_ctx*, const T* u, int64_t ldu, bool u_transposed, int64_t k, const T* b, int64_t ldb,
                int64_t nrhs, T* x, int64_t ldx);

// ------------------------------------------------------------------ chol.cu
// Cholesky factor G = R^H R and R^{-1} of a small Hermitian positive-definite matrix (one CTA).
// status_dev: 4 doubles {breakdown flag, min diag R, max diag R, max |G - I|}.
template <class T>
bool chol_inv(rc_ctx*, const T* g, int64_t ldg, int64_t w, T* r, T* rinv, int64_t ldo, double* status_dev);
int64_t chol_max_width(rc_ctx*, int dtype);
// G += factor * max_j Re(G_jj) * I on the device (no host round trip)
template <class T> void chol_shift(rc_ctx*, T* g, int64_t ldg, int64_t w, double factor);
// Same contract for w <= 2 * chol_max_width: one level of 2 x 2 blocking around the one-CTA kernel (small GEMMs for the
// off-diagonal blocks).
template <class T>
bool chol_inv_blocked(rc_ctx*, const T* g, int64_t ldg, int64_t w, T* r, T* rinv, int64_t ldo, double* status_dev);

// ------------------------------------------------------------------ comm.cu
void comm_get_unique_id(void* out128);
void comm_init(rc_ctx*, const void* id128, int rank, int nranks);
void comm_destroy(rc_ctx*);
void comm_allreduce_sum(rc_ctx*, void* buf, size_t count, int dtype);   // in place
void comm_allreduce_max_f64(rc_ctx*, double* buf, size_t count);
void comm_allgather(rc_ctx*, const void* send, void* recv, size_t bytes_per_rank);
