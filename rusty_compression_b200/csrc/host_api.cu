// C ABI (include/rc_api.h) and the host-side orchestration of the reference's algorithms over
// the CUDA kernels.  Each routine cites the reference lines whose behaviour it reproduces
// (paths relative to /root/reference).  There is no CPU fallback anywhere in this file: every
// numerical step is a kernel launch on the context stream.
#include <type_traits>
#include "host_linalg.cuh"

// ============================================================================ handles
struct rc_qr { rc_matrix* q = nullptr; rc_matrix* r = nullptr; std::vector<uint64_t> ind; };
struct rc_lq { rc_matrix* l = nullptr; rc_matrix* q = nullptr; std::vector<uint64_t> ind; };
struct rc_svd { rc_matrix* u = nullptr; rc_matrix* vt = nullptr; std::vector<double> s; };
struct rc_column_id { rc_matrix* c = nullptr; rc_matrix* z = nullptr; std::vector<uint64_t> col_ind; };
struct rc_row_id { rc_matrix* x = nullptr; rc_matrix* r = nullptr; std::vector<uint64_t> row_ind; };
struct rc_two_sided_id {
    rc_matrix* c = nullptr; rc_matrix* x = nullptr; rc_matrix* r = nullptr;
    std::vector<uint64_t> row_ind, col_ind;
};

namespace {

template <class F>
rc_status guard(rc_ctx* ctx, F&& f) {
    try {
        if (ctx && ctx->stream) {
            DeviceGuard dg(ctx->device);
            f();
        } else {
            f();
        }
        return RC_OK;
    } catch (const RcError& e) {
        if (ctx) ctx->err = e.msg;
        return e.status;
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "host allocation failed";
        return RC_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what();
        return RC_INVALID_ARGUMENT;
    } catch (...) {
        if (ctx) ctx->err = "unknown error";
        return RC_INVALID_ARGUMENT;
    }
}

#define RC_DISPATCH(dt, ...)                                                      \
    switch (dt) {                                                                 \
        case RC_F32: { using T = float; __VA_ARGS__; } break;                     \
        case RC_F64: { using T = double; __VA_ARGS__; } break;                    \
        case RC_C32: { using T = c32; __VA_ARGS__; } break;                       \
        case RC_C64: { using T = c64; __VA_ARGS__; } break;                       \
        default: RC_THROW(RC_INVALID_ARGUMENT, "bad dtype %d", (int)(dt));        \
    }

void check_same(const rc_matrix* a, const rc_matrix* b) {
    RC_REQUIRE(a && b, "null matrix handle");
    RC_REQUIRE(a->dtype == b->dtype, "scalar types differ");
}

double dev_scalar(rc_ctx* c, const double* d) {
    double h = 0.0;
    RC_CUDA(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    return h;
}

}  // namespace

// Stage timer for tuning (option "trace" = 1): synchronises the stream at every mark, so it perturbs the run.
#include <chrono>
static void rc_trace(rc_ctx* c, const char* label) {
    if (!c->trace) return;
    cudaStreamSynchronize(c->stream);
    double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (label) fprintf(stderr, "[rc trace] %-34s %9.3f ms\n", label, now - c->trace_t0);
    c->trace_t0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The generators of the large test matrices leave multi-GB temporaries in the stream-ordered pool (kept there by the
// release threshold of rc_ctx_create).  Next to a matrix that fills most of the device (config 4 at 2 GPUs: 137 GB of
// 180) those cached blocks push every later workspace allocation into the allocator's out-of-memory path (trim, retry:
// hundreds of ms per call, measured 590 ms for one Cholesky-QR2 panel), so a generator hands them back when it is done.
void rc_cache_release(rc_ctx* c) {
    for (auto& kv : c->block_cache) {
        cudaFreeAsync(kv.second.p, kv.second.freed_on);
        c->event_pool.push_back(kv.second.freed_at);
    }
    c->block_cache.clear();
    c->cached_bytes = 0;
}
static void trim_pool(rc_ctx* c) {
    cudaMemPool_t pool;
    rc_cache_release(c);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return;
    for (int i = 0; i < 2; ++i) if (c->aux_stream[i]) cudaStreamSynchronize(c->aux_stream[i]);
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

// ============================================================================ GEMM dispatch
// Replaces ndarray `.dot` (reference N5).  f64/c64 contractions with a plain or (conj-)transposed
// left operand go to the TMA-fed DMMA kernels; everything else to the generic SIMT tiles.
template <class T>
void gemm(rc_ctx* c, RcOp opa, RcOp opb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda,
          const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta) {
    if (M == 0 || N == 0) return;
    if (K == 0) {
        RC_REQUIRE(rc_real(beta) == RealOf<T>(0) && rc_imag(beta) == RealOf<T>(0), "gemm: K == 0 with beta != 0");
        k_fill<T>(c, C, M, N, ldc, rc_zero<T>());
        return;
    }
    bool plain = (rc_real(alpha) == RealOf<T>(1) && rc_imag(alpha) == RealOf<T>(0) &&
                  rc_real(beta) == RealOf<T>(0) && rc_imag(beta) == RealOf<T>(0));
    bool big = (double)M * (double)N * (double)K >= 4.0e6;
    if (plain && big && c->gemm_impl == 0 && opb == RC_OP_N) {
        if constexpr (std::is_same<T, double>::value) {
            if (gemm_dmma_f64(c, opa != RC_OP_N, M, N, K, A, lda, B, ldb, C, ldc)) return;
        } else if constexpr (std::is_same<T, c64>::value) {
            if (opa == RC_OP_N || opa == RC_OP_H)
                if (gemm_dmma_c64(c, opa == RC_OP_H, M, N, K, A, lda, B, ldb, C, ldc)) return;
        } else if constexpr (std::is_same<T, float>::value) {
            // (fewer than ~1/4 of the SMs' worth of 128-row tiles and a long K: the split-K SIMT tiles win)
            if (opa == RC_OP_N && M >= 128 && (M >= 128 * 37 || K < 4096)) {
                if (gemm_tf32x3_f32(c, M, N, K, A, lda, B, ldb, C, ldc)) return;
            } else if (opa != RC_OP_N && M >= 32 && K >= 256) {
                if (gemm_tf32x3_f32_tn(c, M, N, K, A, lda, B, ldb, C, ldc)) return;
            }
        } else if constexpr (std::is_same<T, c32>::value) {
            if (opa == RC_OP_N && M >= 128) {
                if (gemm_tf32x3_c32(c, false, M, N, K, A, lda, B, ldb, C, ldc)) return;
            } else if (opa == RC_OP_H && M >= 64 && K >= 256) {
                if (gemm_tf32x3_c32(c, true, M, N, K, A, lda, B, ldb, C, ldc)) return;
            }
        }
    }
    gemm_generic<T>(c, opa, opb, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
    c->gemm_flops += (ScalarTraits<T>::is_complex ? 8 : 2) * M * N * K;
}
template void gemm<float>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float, float);
template void gemm<double>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, double, double);
template void gemm<c32>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const c32*, int64_t, const c32*, int64_t, c32*, int64_t, c32, c32);
template void gemm<c64>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const c64*, int64_t, const c64*, int64_t, c64*, int64_t, c64, c64);

// ============================================================================ pivoted QR
namespace {

// TSQR of a (possibly row-sharded) tall panel, with the extra all-gather level across ranks.
template <class T>
struct DistTsqr {
    rc_ctx* c;
    bool sharded;
    int64_t w;
    TsqrFactor<T> local, top;
    DevBuf<T> stack;
    const T* r() const { return sharded ? top.r : local.r; }
    void factor(rc_ctx* ctx, T* y, int64_t ld, int64_t m, int64_t width, bool is_sharded) {
        c = ctx; w = width; sharded = is_sharded && ctx->nranks > 1;
        tsqr_factor<T>(c, y, ld, m, w, local);
        if (sharded) {
            // all-gather of the w x w R factors; every rank factors the stack redundantly
            stack.alloc(c, (size_t)c->nranks * w * w);
            comm_allgather(c, local.r, stack.p, (size_t)w * w * sizeof(T));
            tsqr_factor<T>(c, stack.p, w, (int64_t)c->nranks * w, w, top);
        }
    }
    // out (m x nc) = Q [ctop ; 0]
    void apply(const T* ctop, int64_t ldc, int64_t nc, T* out, int64_t ldo) {
        if (!sharded) { tsqr_apply_q<T>(c, local, ctop, ldc, nc, out, ldo); return; }
        DevBuf<T> mid(c, (size_t)c->nranks * w * nc);
        tsqr_apply_q<T>(c, top, ctop, ldc, nc, mid.p, nc);
        tsqr_apply_q<T>(c, local, mid.p + (size_t)c->rank * w * nc, nc, nc, out, ldo);
    }
};

void download_ind(rc_ctx* c, const int* dind, int64_t n, std::vector<uint64_t>& ind) {
    std::vector<int> h((size_t)n);
    RC_CUDA(cudaMemcpyAsync(h.data(), dind, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    ind.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (h[i] < 0 || h[i] >= n) RC_THROW(RC_PIVOTED_QR_ERROR, "pivoted QR produced an invalid permutation");
        ind[i] = (uint64_t)h[i];
    }
}


// Acceptance test of a Cholesky-QR2 panel from its status words (chol.cu; 4 per round: breakdown flag, min and max of
// diag(R), max |G - I|): round 1 without breakdown, round 2 without breakdown and with a Gram matrix of q1 within 0.25
// of I, and one of
//   (a) diag(R1) spanning less than 1e6 (2e2 in single precision), or
//   (b) double precision: q1 orthonormal to 1e-3 already (and diag(R2) within a factor 2).  G2 - I is the backward
//       error of round 1 in the scaling of the columns of Y, so this accepts GRADED sketches whatever their diagonal
//       spans -- Cholesky is invariant under column scaling, and the sketches of a power iteration (Z = A^H q,
//       Y = A w with q, w ordered by a pivoted QR) have columns graded like the spectrum: measured on a 74-column
//       sketch over twelve decades, diag(R1) spans 2e12 and max |G2 - I| = 1e-13.
bool cholqr2_acceptable(const double* h, bool single, bool shifted = false) {
    const double max_ratio = single ? 2.0e2 : 1.0e6;
    const bool ok0 = !shifted || (h[8] == 0.0 && h[9] > 0.0 && h[12] == 0.0 && h[13] > 0.0);   // the two shifted rounds
    const bool ok2 = h[4] == 0.0 && h[5] > 0.0 && h[7] <= 0.25;
    const bool ok1 = h[0] == 0.0 && h[1] > 0.0 &&
                     (h[2] / h[1] <= max_ratio || (!single && ok2 && h[7] <= 1.0e-3 && h[6] / h[5] <= 2.0));
    return ok0 && ok1 && ok2;
}

// Deferred (speculative) region: see rc_ctx::defer_depth.  finish_deferred() reads all collected status words with
// ONE host synchronisation and says whether every speculative panel was acceptable.
struct DeferScope {
    rc_ctx* c;
    explicit DeferScope(rc_ctx* ctx) : c(ctx) { c->defer_depth++; }
    ~DeferScope() { c->defer_depth--; }
};
bool finish_deferred(rc_ctx* c) {
    bool ok = true;
    const size_t n = c->deferred.size();
    if (n == 0) { RC_CUDA(cudaStreamSynchronize(c->stream)); return true; }
    std::vector<double> h(16 * n);
    for (size_t i = 0; i < n; ++i)
        RC_CUDA(cudaMemcpyAsync(h.data() + 16 * i, c->deferred[i].status, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) {
        if (!cholqr2_acceptable(h.data() + 16 * i, c->deferred[i].single, c->deferred[i].shifted)) {
            ok = false; c->cholqr_fallbacks++; c->cholqr_used--;
            if (c->deferred[i].shifted) c->cholqr_shifted--;
        }
        rc_dev_free(c, c->deferred[i].status);
    }
    c->deferred.clear();
    return ok;
}
void drop_deferred(rc_ctx* c) {
    for (auto& d : c->deferred) rc_dev_free(c, d.status);
    c->deferred.clear();
}
// Early check inside a deferred region (the status words collected so far may have been written on the auxiliary
// streams): true when every speculative panel so far was acceptable.
bool check_deferred_now(rc_ctx* c) {
    for (int i = 0; i < 2; ++i) if (c->aux_stream[i]) RC_CUDA(cudaStreamSynchronize(c->aux_stream[i]));
    return finish_deferred(c);
}
// Non-speculative redo of a rejected panel on the next route down (shifted Cholesky-QR where it applies, with the
// Householder TSQR behind it; the Householder TSQR at once in single precision).
struct RedoScope {
    rc_ctx* c;
    int saved_depth;
    bool saved_shifted, saved_householder;
    RedoScope(rc_ctx* ctx, int dtype) : c(ctx), saved_depth(ctx->defer_depth), saved_shifted(ctx->force_shifted),
                                        saved_householder(ctx->force_householder) {
        const bool single = (dtype == RC_F32 || dtype == RC_C32);
        c->defer_depth = 0;
        if (c->shifted_cholqr && !single) c->force_shifted = true; else c->force_householder = true;
    }
    ~RedoScope() { c->defer_depth = saved_depth; c->force_shifted = saved_shifted; c->force_householder = saved_householder; }
};

// Work on an auxiliary stream (with its own tile-scheduler scratch) for the lifetime of the scope.  Everything the
// library enqueues -- kernels, stream-ordered allocations and frees, collectives -- follows c->stream.
struct StreamScope {
    rc_ctx* c;
    cudaStream_t saved_stream;
    int* saved_counter;
    StreamScope(rc_ctx* ctx, int idx) : c(ctx), saved_stream(ctx->stream), saved_counter(ctx->tile_counter) {
        if (!c->aux_stream[idx]) {
            RC_CUDA(cudaStreamCreateWithFlags(&c->aux_stream[idx], cudaStreamNonBlocking));
            RC_CUDA(cudaMalloc((void**)&c->aux_tile_counter[idx], 256));
        }
        c->stream = c->aux_stream[idx];
        c->tile_counter = c->aux_tile_counter[idx];
    }
    ~StreamScope() { c->stream = saved_stream; c->tile_counter = saved_counter; }
};
cudaEvent_t aux_event(rc_ctx* c, int i) {
    if (!c->aux_event[i]) RC_CUDA(cudaEventCreateWithFlags(&c->aux_event[i], cudaEventDisableTiming));
    return c->aux_event[i];
}
// Share of the machine for everything launched inside the scope.  Every persistent kernel sizes its grid from
// rc_ctx::sm_count (one CTA per SM for the tcgen05 kernels, two for the DMMA kernels), so two scopes whose budgets add
// up to the device run side by side on two streams with every CTA of both resident at once -- no CTA of one ever
// waits behind the persistent CTAs of the other, which is what serialises two full-size grids.
struct SmBudget {
    rc_ctx* c;
    int saved;
    SmBudget(rc_ctx* ctx, int sms) : c(ctx), saved(ctx->sm_count) { c->sm_count = std::max(1, std::min(sms, saved)); }
    ~SmBudget() { c->sm_count = saved; }
};

// Cholesky-QR2 fast path for a tall, numerically full-rank panel (Y = Q R with Q = q1 * rinv2):
// two rounds of Gram matrix (one TN GEMM + all-reduce across row shards) -> small Cholesky -> Y R^{-1}.
// All the work is GEMM-shaped, so it runs at tensor-pipe speed instead of the latency-bound
// reflector chain of the Householder TSQR.  It is only taken when the panel is well conditioned
// (diag(R) ratio below 1/sqrt(eps)-ish, second Gram matrix close to I); otherwise -- or when the
// Cholesky breaks down -- the caller falls back to the unconditionally stable Householder TSQR
// with Y untouched.  Backward error and orthogonality are O(eps) in the accepted regime
// (Yamamoto et al. 2015), the same class as Householder, which is what pivot parity needs.
// panel_scale (optional, host, in/out): the largest diagonal entry of R over the panels of a wider factorisation so far
// (total width full_w).  With it the status words are read back at once (no speculation) and a panel whose smallest
// diagonal entry is at most 10 eps sqrt(full_w) times that scale is REJECTED however well conditioned it is relative to
// itself: a panel that has no direction of its own left after the projection is rounding noise (measured ~50 eps of the
// scale at full_w = 900 in f64 and at full_w = 266 in f32, i.e. ~3 eps sqrt(full_w): the error of Q (Q^H y), not of the
// length-m dot products), and a Cholesky-QR2 of noise is orthogonal to the previous panels only to eps x (norm before
// the projection / norm of the noise) -- the loss compounds from panel to panel (measured on an exact rank-50
// 1500 x 900 matrix: 3e-15, 2e-14, 9e-12, 1e-6, 0.6).  The window must not be wider than that: the last panel of
// config 4's f32 sketch holds legitimate directions at 7e-5 of the scale.
// shifted = true: shifted Cholesky-QR (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa 2020, "sCholQR3", with the
// shifted round applied twice).  A round on G + s I cannot break down and divides cond(Y) by ~1/sqrt(s_rel); two of
// them bring any sketch with cond(Y) up to ~1e16 below 1e6, which the two plain rounds then finish to O(u)
// orthogonality: R = R2 R1 R0b R0a, residual O(u ||Y||) like the Householder route, all of it GEMM-shaped (four rounds
// = 0.9 ms for 65 536 x 74 in double against 1.8 ms for the TSQR tree).  The shift is s = 100 u (sqrt(m) + w) w max_j G_jj:
// the paper's bound 11 (m w + w (w + 1)) u ||Y||^2 assumes worst-case (linear in m) error growth in the Gram matrix and,
// at m = 65 536, leaves cond ~ 6e7 after one round of a twelve-decade sketch (measured: rejected); the accumulation
// error of a length-m dot product grows like sqrt(m), which this shift still covers ~100 times over, and a Cholesky
// that breaks down regardless is caught by its status word like any other (-> Householder TSQR).
template <class T>
bool cholqr_rounds(rc_ctx* c, const T* y, int64_t ldy, int64_t m, int64_t w, bool sharded, int dtype, bool shifted,
                   DevBuf<T>& q1, DevBuf<T>& rinv2, DevBuf<T>& rfac, int64_t& lds, double* panel_scale, int64_t full_w) {
    lds = rc_pad_ld(dtype, w);
    DevBuf<T> g(c, (size_t)w * lds), r1(c, (size_t)w * lds), rinv1(c, (size_t)w * lds), r2(c, (size_t)w * lds);
    DevBuf<double> status(c, 16);
    double h[16];
    const bool single = (dtype == RC_F32 || dtype == RC_C32);
    auto gram_chol = [&](const T* x, int64_t ldx, T* rr, T* ri, double* st, double shift_factor) -> bool {
        if (sharded && lds != w) RC_CUDA(cudaMemsetAsync(g.p, 0, sizeof(T) * w * lds, c->stream));   // padding is summed too
        gemm<T>(c, RC_OP_H, RC_OP_N, w, w, m, x, ldx, x, ldx, g.p, lds, rc_one<T>(), rc_zero<T>());
        if (sharded) comm_allreduce_sum(c, g.p, (size_t)w * lds, dtype);      // one all-reduce of the Gram matrix
        if (shift_factor > 0.0) chol_shift<T>(c, g.p, lds, w, shift_factor);
        return chol_inv_blocked<T>(c, g.p, lds, w, rr, ri, lds, st);
    };
    DevBuf<T> ya, yb, r0;
    if (shifted) {
        const double u = single ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
        const double m_glob = sharded ? (double)m * c->nranks : (double)m;     // (upper bound of the global row count)
        const double s_rel = 100.0 * u * (std::sqrt(m_glob) + (double)w) * (double)w;
        DevBuf<T> r0a(c, (size_t)w * lds), r0b(c, (size_t)w * lds), rinv0(c, (size_t)w * lds);
        if (!gram_chol(y, ldy, r0a.p, rinv0.p, status.p + 8, s_rel)) return false;
        ya.alloc(c, (size_t)m * lds);
        gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, y, ldy, rinv0.p, lds, ya.p, lds, rc_one<T>(), rc_zero<T>());
        if (!gram_chol(ya.p, lds, r0b.p, rinv0.p, status.p + 12, s_rel)) return false;
        yb.alloc(c, (size_t)m * lds);
        gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, ya.p, lds, rinv0.p, lds, yb.p, lds, rc_one<T>(), rc_zero<T>());
        ya.release();
        r0.alloc(c, (size_t)w * lds);
        gemm<T>(c, RC_OP_N, RC_OP_N, w, w, w, r0b.p, lds, r0a.p, lds, r0.p, lds, rc_one<T>(), rc_zero<T>());
        y = yb.p; ldy = lds;
    } else {
        RC_CUDA(cudaMemsetAsync(status.p + 8, 0, 8 * sizeof(double), c->stream));
    }
    // Both rounds are enqueued back to back and their status words are read with ONE host synchronisation at
    // the end (a sync after each Cholesky left the GPU idle for a launch round trip twice per panel).  If round 1
    // broke down, round 2 ran on garbage: its loops are data independent, nothing it wrote is used, and Y is
    // untouched for the Householder fallback.
    if (!gram_chol(y, ldy, r1.p, rinv1.p, status.p, 0.0)) return false;
    q1.alloc(c, (size_t)m * lds);
    gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, y, ldy, rinv1.p, lds, q1.p, lds, rc_one<T>(), rc_zero<T>());
    yb.release();
    rinv2.alloc(c, (size_t)w * lds);
    if (!gram_chol(q1.p, lds, r2.p, rinv2.p, status.p + 4, 0.0)) return false;
    rfac.alloc(c, (size_t)w * lds);
    gemm<T>(c, RC_OP_N, RC_OP_N, w, w, w, r2.p, lds, r1.p, lds, rfac.p, lds, rc_one<T>(), rc_zero<T>());
    if (shifted) {
        DevBuf<T> r21(c, (size_t)w * lds);
        k_copy<T>(c, r21.p, lds, rfac.p, lds, w, w);
        gemm<T>(c, RC_OP_N, RC_OP_N, w, w, w, r21.p, lds, r0.p, lds, rfac.p, lds, rc_one<T>(), rc_zero<T>());
    }
    if (c->defer_depth > 0 && !panel_scale) {
        // speculative: carry on as if both rounds were accepted; the status words are checked at the end of the
        // deferred region (finish_deferred), which re-runs it on the next route when a panel was not acceptable
        c->deferred.push_back({status.take(), single, shifted});
        c->cholqr_used++;
        if (shifted) c->cholqr_shifted++;
        return true;
    }
    RC_CUDA(cudaMemcpyAsync(h, status.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    if (!cholqr2_acceptable(h, single, shifted)) return false;
    if (panel_scale) {
        const double eps = single ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
        if (!(h[1] > 10.0 * eps * std::sqrt((double)std::max<int64_t>(full_w, w)) * std::max(*panel_scale, h[2]))) return false;
        *panel_scale = std::max(*panel_scale, h[2]);
    }
    rc_trace(c, shifted ? "  shifted cholqr: 4 x (gram + chol), 3 x apply" : "  cholqr2: 2 x (gram + chol), q1 = Y rinv1");
    c->cholqr_used++;
    if (shifted) c->cholqr_shifted++;
    return true;
}

template <class T>
bool cholqr2(rc_ctx* c, const T* y, int64_t ldy, int64_t m, int64_t w, bool sharded, int dtype,
             DevBuf<T>& q1, DevBuf<T>& rinv2, DevBuf<T>& rfac, int64_t& lds, double* panel_scale = nullptr, int64_t full_w = 0) {
    if (c->qr_mode == 1 || c->force_householder) return false;
    // (one CTA holds Gram matrices up to chol_max_width; twice that with one level of 2 x 2 blocking, chol_inv_blocked: the
    // 138-column c64 sketch of config 5 is ONE Cholesky-QR2 instead of two panels with block Gram-Schmidt between them.
    // Double precision only: in single precision the acceptance window is diag(R1) within 2e2, which a wide sketch of
    // a decaying spectrum does not meet as a whole -- config 4's 266 columns span 4 decades, measured: the rejected
    // full-width attempt cost 8 ms -- while its 92-column panels do)
    const bool single_prec = (dtype == RC_F32 || dtype == RC_C32);
    const int64_t wlim = (single_prec ? 1 : 2) * chol_max_width(c, dtype);
    if ((!sharded && m < 4 * w) || w > wlim || w < 2) return false;
    // the shifted route: double precision, whole sketches only (the panels of a wider factorisation have their own
    // rank-deficiency handling in panel_qr)
    const bool may_shift = c->shifted_cholqr && !single_prec && !panel_scale;
    if (c->force_shifted && may_shift)
        return cholqr_rounds<T>(c, y, ldy, m, w, sharded, dtype, true, q1, rinv2, rfac, lds, panel_scale, full_w);
    if (cholqr_rounds<T>(c, y, ldy, m, w, sharded, dtype, false, q1, rinv2, rfac, lds, panel_scale, full_w)) return true;
    if (c->defer_depth > 0 && !panel_scale) return false;      // (not reached: a speculative attempt reports success)
    c->cholqr_fallbacks++;
    if (may_shift) {
        q1.release(); rinv2.release(); rfac.release();
        if (cholqr_rounds<T>(c, y, ldy, m, w, sharded, dtype, true, q1, rinv2, rfac, lds, panel_scale, full_w)) return true;
        c->cholqr_fallbacks++;
    }
    return false;
}

// Pivoted QR of the small w x w factor of a tall panel (row-major `rf`): R (into r), pivots (dind) and the first ncq
// columns of its orthogonal factor (q1, row-major).  One fused kernel when the factor fits it (pivqr.cu), else
// transpose -> cooperative / one-CTA factorisation -> form Q.
template <class T>
void small_pivqr(rc_ctx* c, const T* rf, int64_t ldrf, int64_t w, int64_t ncq, rc_matrix* r, int* dind, T* q1, int64_t ldq1,
                 DevBuf<T>& wc, DevBuf<T>& vbuf, DevBuf<T>& tau) {
    if (pivqr_fused<T>(c, rf, ldrf, 0, w, w, ncq, P<T>(r), r->ld, dind, q1, ldq1)) return;
    k_transpose<T>(c, wc.p, w, rf, ldrf, w, w, false);          // column-major copy = transpose of the row-major factor
    pivqr_factor<T>(c, wc.p, w, w, w, P<T>(r), r->ld, dind, vbuf.p, tau.p);
    pivqr_form_q<T>(c, vbuf.p, tau.p, w, w, ncq, q1, ldq1);
}

// Unpivoted QR of a tall panel that is wider than one shared-memory block: Y = Qfull R0 by column panels.
// Panels of at most min(wmax, Cholesky width) columns; each is orthogonalised against the previous ones (block
// classical Gram-Schmidt, twice) and then factored on its own: Cholesky-QR2 when it is well conditioned (all
// GEMM-shaped, tensor pipe), Householder TSQR otherwise.  qfull: m x w (ld = ldq), r0: w x w row-major (ld = w).
template <class T>
void panel_qr(rc_ctx* c, T* y, int64_t ldy, int64_t m, int64_t w, bool sharded, int dtype,
              DevBuf<T>& qfull, int64_t& ldq, DevBuf<T>& r0) {
    const int64_t wmax = tsqr_max_width(c, dtype);
    int64_t wpan = std::min(wmax, c->qr_mode == 1 ? wmax : chol_max_width(c, dtype));
    const bool single = (dtype == RC_F32 || dtype == RC_C32);
    (void)single;
    // f32 on tcgen05: every Gram-type product of a panel is tiled 128 (output rows) x one chunk of <= 96 columns, and each
    // (tile, chunk) item streams all m rows: a 133-column panel pays 2 x 2 items per product, a panel of <= 96 columns
    // pays one (config 4, l = 266: three panels of 92 / 92 / 82 instead of two of 133 -- 12 instead of 24 items per
    // sketch, K = 92 instead of 133 in the small-K products)
    if (dtype == RC_F32 && c->gemm_impl == 0 && c->qr_mode != 1 && m >= 4096) wpan = std::min<int64_t>(wpan, 96);
    int64_t npan = (w + wpan - 1) / wpan;
    int64_t wp = (w + npan - 1) / npan;
    {   // panel starts on 16-byte boundaries (TMA operand alignment), as long as the panel still fits
        const int64_t e = std::max<int64_t>(1, (int64_t)(16 / rc_dtype_size(dtype)));
        const int64_t up = (wp + e - 1) / e * e;
        if (up <= wpan) wp = up;
    }

    ldq = rc_pad_ld(dtype, w);                        // 16-byte row pitch: the tensor-pipe GEMMs can take it
    qfull.alloc(c, (size_t)m * ldq);
    r0.alloc(c, (size_t)w * w);
    k_fill<T>(c, r0.p, w, w, w, rc_zero<T>());
    double scale = 0.0;                               // largest |r_jj| over the panels so far (null-direction threshold)
    for (int64_t c0 = 0; c0 < w; c0 += wp) {
        int64_t cw = std::min(wp, w - c0);
        T* yp = y + c0;
        if (c0 > 0) {
            const int64_t ldt = rc_pad_ld(dtype, cw);
            DevBuf<T> t(c, (size_t)c0 * ldt), proj(c, (size_t)m * ldt);
            for (int pass = 0; pass < 2; ++pass) {
                if (sharded && ldt != cw) RC_CUDA(cudaMemsetAsync(t.p, 0, sizeof(T) * c0 * ldt, c->stream));
                gemm<T>(c, RC_OP_H, RC_OP_N, c0, cw, m, qfull.p, ldq, yp, ldy, t.p, ldt, rc_one<T>(), rc_zero<T>());
                if (sharded) comm_allreduce_sum(c, t.p, (size_t)c0 * ldt, dtype);
                // (the subtraction stays a separate bandwidth kernel: folded into the epilogue of the tcgen05 GEMM -- one thread
                // per output row, so its loads are as uncoalesced as its stores -- the short-K product took 0.9 ms instead of
                // 0.3 + 0.45 ms at m = 2^20, l = 92)
                gemm<T>(c, RC_OP_N, RC_OP_N, m, cw, c0, qfull.p, ldq, t.p, ldt, proj.p, ldt, rc_one<T>(), rc_zero<T>());
                k_sub<T>(c, yp, ldy, yp, ldy, proj.p, ldt, m, cw);                    // Y_p -= Q (Q^H Y_p)
                k_add<T>(c, r0.p + c0, w, r0.p + c0, w, t.p, ldt, c0, cw);             // R0[0:c0, c0:c0+cw] += t
            }
            rc_trace(c, "pqr_tall: panel projection x2");
        }
        DevBuf<T> pq1, prinv2, prfac;
        int64_t lds = 0;
        if (cholqr2<T>(c, yp, ldy, m, cw, sharded, dtype, pq1, prinv2, prfac, lds, &scale, w)) {
            rc_trace(c, "pqr_tall: panel cholqr2");
            gemm<T>(c, RC_OP_N, RC_OP_N, m, cw, cw, pq1.p, lds, prinv2.p, lds, qfull.p + c0, ldq, rc_one<T>(), rc_zero<T>());
            k_copy<T>(c, r0.p + c0 * w + c0, w, prfac.p, lds, cw, cw);
            rc_trace(c, "pqr_tall: panel Q = q1 rinv2");
        } else {
            // Householder fallback: the projected panel is ill-conditioned or rank-deficient (a low-rank operator sampled
            // with more columns than its rank, an exactly rank-deficient dense matrix, the zero matrix).  Its TSQR is
            // Y_p = Q_p R_pp with Q_p orthonormal, but the directions that belong to (numerically) zero rows of R_pp are
            // whatever the reflectors leave -- unit vectors for a zero panel -- and Q_p is orthogonal to the previous
            // panels only to eps * cond(Y_p).  ?geqp3 returns an orthonormal Q whatever the rank, so:
            //   1. null directions (|r_jj| <= eps sqrt(w) x the largest diagonal entry so far) are replaced by
            //      Gaussian vectors (they multiply rows of R_pp that are zero to that level),
            //   2. Q_p is orthogonalised against the previous panels (twice) and factored again, Q_p = Q_prev T + Q_p' R',
            //   3. R0 takes the correction: R0[0:c0, panel] += T R_pp, R0[panel, panel] = R' R_pp.
            DistTsqr<T> ts;
            ts.factor(c, yp, ldy, m, cw, sharded);
            DevBuf<T> eye(c, (size_t)cw * cw), rpp(c, (size_t)cw * cw);
            k_copy<T>(c, rpp.p, cw, ts.r(), cw, cw, cw);
            k_eye<T>(c, eye.p, cw, cw, cw);
            const int64_t ldz = rc_pad_ld(dtype, cw);
            DevBuf<T> z(c, (size_t)m * ldz);
            ts.apply(eye.p, cw, cw, z.p, ldz);
            {
                const double eps = single ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
                const int64_t m_glob = sharded ? (int64_t)c->nranks * m : m;
                std::vector<T> hr((size_t)cw * cw);
                RC_CUDA(cudaMemcpyAsync(hr.data(), rpp.p, sizeof(T) * cw * cw, cudaMemcpyDeviceToHost, c->stream));
                RC_CUDA(cudaStreamSynchronize(c->stream));
                for (int64_t j = 0; j < cw; ++j) { const double d = rc_abs(hr[(size_t)j * cw + j]); if (d == d) scale = std::max(scale, d); }
                std::vector<int> hflags((size_t)cw);
                int64_t nnull = 0;
                // (eps sqrt(w): below the rounding noise of a projected panel, ~3 eps sqrt(w) -- what has to go are the
                // STRUCTURED leftovers of exactly zero rows (unit vectors); a direction at the noise level is random
                // already and the re-orthogonalisation below deals with it.  The replaced directions cost |r_jj| in
                // the reconstruction, so a wide window shows: 1e3 eps sqrt(m) gave 1e-3 relative in f32.)
                const double thr = eps * std::sqrt((double)w) * scale;
                for (int64_t j = 0; j < cw; ++j) { hflags[j] = !(rc_abs(hr[(size_t)j * cw + j]) > thr) ? 1 : 0; nnull += hflags[j]; }   // (NaN counts as null)
                if (nnull > 0) {
                    DevBuf<int> flags(c, (size_t)cw);
                    RC_CUDA(cudaMemcpyAsync(flags.p, hflags.data(), sizeof(int) * cw, cudaMemcpyHostToDevice, c->stream));
                    DevBuf<T> noise(c, (size_t)m * ldz);
                    k_gaussian<T>(c, noise.p, m, cw, ldz, 0x9e3779b97f4a7c15ull + (uint64_t)c0, 900u + (uint32_t)c->rank, 0);
                    k_replace_flagged_columns<T>(c, z.p, ldz, noise.p, ldz, m, cw, flags.p, 1.0 / std::sqrt((double)m_glob));
                    RC_CUDA(cudaStreamSynchronize(c->stream));            // (hflags is read by the copy above)
                }
            }
            DevBuf<T> tacc;
            if (c0 > 0) {
                const int64_t ldt = rc_pad_ld(dtype, cw);
                tacc.alloc(c, (size_t)c0 * ldt);
                DevBuf<T> t(c, (size_t)c0 * ldt), proj(c, (size_t)m * ldt);
                k_fill<T>(c, tacc.p, c0, cw, ldt, rc_zero<T>());
                for (int pass = 0; pass < 2; ++pass) {
                    if (sharded && ldt != cw) RC_CUDA(cudaMemsetAsync(t.p, 0, sizeof(T) * c0 * ldt, c->stream));
                    gemm<T>(c, RC_OP_H, RC_OP_N, c0, cw, m, qfull.p, ldq, z.p, ldz, t.p, ldt, rc_one<T>(), rc_zero<T>());
                    if (sharded) comm_allreduce_sum(c, t.p, (size_t)c0 * ldt, dtype);
                    gemm<T>(c, RC_OP_N, RC_OP_N, m, cw, c0, qfull.p, ldq, t.p, ldt, proj.p, ldt, rc_one<T>(), rc_zero<T>());
                    k_sub<T>(c, z.p, ldz, z.p, ldz, proj.p, ldt, m, cw);
                    k_add<T>(c, tacc.p, ldt, tacc.p, ldt, t.p, ldt, c0, cw);
                }
                // R0[0:c0, panel] += T R_pp
                DevBuf<T> trp(c, (size_t)c0 * cw);
                gemm<T>(c, RC_OP_N, RC_OP_N, c0, cw, cw, tacc.p, ldt, rpp.p, cw, trp.p, cw, rc_one<T>(), rc_zero<T>());
                k_add<T>(c, r0.p + c0, w, r0.p + c0, w, trp.p, cw, c0, cw);
            }
            DistTsqr<T> ts2;
            ts2.factor(c, z.p, ldz, m, cw, sharded);
            ts2.apply(eye.p, cw, cw, qfull.p + c0, ldq);
            // R0[panel, panel] = R' R_pp
            DevBuf<T> rr(c, (size_t)cw * cw);
            gemm<T>(c, RC_OP_N, RC_OP_N, cw, cw, cw, ts2.r(), cw, rpp.p, cw, rr.p, cw, rc_one<T>(), rc_zero<T>());
            k_triu<T>(c, rr.p, cw, cw, cw);
            k_copy<T>(c, r0.p + c0 * w + c0, w, rr.p, cw, cw, cw);
            rc_trace(c, "pqr_tall: panel Householder fallback (re-orthogonalised)");
        }
    }
}

// Tall-skinny route: Y = Q0 R0 by (distributed) Householder TSQR, pivoting on R0, Q = Q0 Q1.
// Panels wider than the shared-memory limit are orthogonalised block by block against the
// previous panels (two projection passes) before their own TSQR.
template <class T>
void pqr_tall(rc_ctx* c, T* y, int64_t ldy, int64_t m, int64_t w, int64_t ncq, bool sharded, int dtype, QrParts& out) {
    const int64_t wmax = tsqr_max_width(c, dtype);
    DevBuf<T> wc(c, (size_t)w * w), vbuf(c, (size_t)w * w), tau(c, (size_t)w);
    DevBuf<int> dind(c, (size_t)w);
    MatPtr r(mat_new(c, dtype, w, w));
    MatPtr q;                                   // (allocated where it is formed: m x ncq is GBs for the tall shards)
    DevBuf<T> q1(c, (size_t)w * ncq);

    rc_trace(c, nullptr);
    DevBuf<T> cq1, crinv2, crfac;
    int64_t lds = 0;
    if (cholqr2<T>(c, y, ldy, m, w, sharded, dtype, cq1, crinv2, crfac, lds)) {
        rc_trace(c, "pqr_tall: cholqr2");
        // pivot on R = R2 R1, then Q = q1 (R2^{-1} Q1piv)
        small_pivqr<T>(c, crfac.p, lds, w, ncq, r.get(), dind.p, q1.p, ncq, wc, vbuf, tau);
        DevBuf<T> tq(c, (size_t)w * rc_pad_ld(dtype, ncq));
        const int64_t ldtq = rc_pad_ld(dtype, ncq);
        gemm<T>(c, RC_OP_N, RC_OP_N, w, ncq, w, crinv2.p, lds, q1.p, ncq, tq.p, ldtq, rc_one<T>(), rc_zero<T>());
        q.reset(mat_new(c, dtype, m, ncq));
        gemm<T>(c, RC_OP_N, RC_OP_N, m, ncq, w, cq1.p, lds, tq.p, ldtq, P<T>(q.get()), q->ld, rc_one<T>(), rc_zero<T>());
    } else if (w <= wmax) {
        DistTsqr<T> ts;
        ts.factor(c, y, ldy, m, w, sharded);
        small_pivqr<T>(c, ts.r(), w, w, ncq, r.get(), dind.p, q1.p, ncq, wc, vbuf, tau);
        q.reset(mat_new(c, dtype, m, ncq));
        ts.apply(q1.p, ncq, ncq, P<T>(q.get()), q->ld);
    } else {
        DevBuf<T> qfull, r0p;
        int64_t ldq = 0;
        panel_qr<T>(c, y, ldy, m, w, sharded, dtype, qfull, ldq, r0p);
        const int64_t ldq1 = rc_pad_ld(dtype, ncq);
        DevBuf<T> q1p(c, (size_t)w * ldq1);
        small_pivqr<T>(c, r0p.p, w, w, ncq, r.get(), dind.p, q1p.p, ldq1, wc, vbuf, tau);
        rc_trace(c, "pqr_tall: pivoted QR of R");
        q.reset(mat_new(c, dtype, m, ncq));
        gemm<T>(c, RC_OP_N, RC_OP_N, m, ncq, w, qfull.p, ldq, q1p.p, ldq1, P<T>(q.get()), q->ld, rc_one<T>(), rc_zero<T>());
        rc_trace(c, "pqr_tall: form Q1, Q = Qfull Q1");
    }
    if (out.want_ind) download_ind(c, dind.p, w, out.ind);
    out.q.reset(q.release());
    out.r.reset(r.release());
}

}  // namespace

// PivotedQR::pivoted_qr (src/pivoted_qr.rs:25-31, 81-119): arr[:, ind] = q r.
// input_is_conj_transposed: `arr` holds S (n x p row-major) and the matrix factored is M = S^H
// (p x n) -- the access pattern of pivoted_lq / LQ::compute_from (src/pivoted_qr.rs:32-41,
// src/qr.rs:354-362) and of b = (A^H Q)^H (src/qr.rs:315), without materialising the transpose.
template <class T>
void pivoted_qr_impl(rc_ctx* c, const rc_matrix* arr, bool input_is_conj_transposed, int64_t ncq,
                     bool may_destroy, QrParts& out) {
    const int dtype = arr->dtype;
    const int64_t p = input_is_conj_transposed ? arr->cols : arr->rows;   // rows of M
    const int64_t n = input_is_conj_transposed ? arr->rows : arr->cols;   // cols of M
    RC_REQUIRE(p > 0 && n > 0, "pivoted_qr: empty matrix");
    const int64_t kk = std::min(p, n);
    if (ncq < 0 || ncq > kk) ncq = kk;
    const bool sharded = !input_is_conj_transposed && mat_sharded(arr);
    // Single-precision inputs of moderate size are factored entirely in double (input widened exactly, Q and R
    // rounded back once): the pivot sequence is then ?geqp3's in double precision on the same f32 / c32 data, for
    // tall inputs too (whose reduction to R would otherwise carry f32 roundoff into the deep pivot norms).  Larger
    // single-precision inputs keep the unpivoted reduction in working precision and only pivot in double (pivqr.cu).
    using W = typename AccOf<T>::type;
    if constexpr (!std::is_same<T, W>::value) {
        // (general dense single-precision matrices -- both dimensions large -- also go through double as long as the
        // widened copy stays modest: the pivots deep into the spectrum sit at 1e-4..1e-7 of the leading norm, below
        // the roundoff an f32 reduction to R would leave in them)
        const bool dense_general = std::min(p, n) > 512 && p * n <= ((int64_t)1 << 27);
        if (c->pivot_f64 && !sharded && (p * n <= ((int64_t)1 << 20) || dense_general)) {
            MatPtr wide(mat_new(c, dtype | 1, arr->rows, arr->cols));
            k_cast<W, T>(c, P<W>(wide.get()), wide->ld, P<T>(arr), arr->ld, arr->rows, arr->cols);
            QrParts ow;
            ow.want_ind = out.want_ind;
            pivoted_qr_impl<W>(c, wide.get(), input_is_conj_transposed, ncq, true, ow);
            MatPtr q(mat_new(c, dtype, ow.q->rows, ow.q->cols)), r(mat_new(c, dtype, ow.r->rows, ow.r->cols));
            k_cast<T, W>(c, P<T>(q.get()), q->ld, P<W>(ow.q.get()), ow.q->ld, q->rows, q->cols);
            k_cast<T, W>(c, P<T>(r.get()), r->ld, P<W>(ow.r.get()), ow.r->ld, r->rows, r->cols);
            out.q.reset(q.release());
            out.r.reset(r.release());
            out.ind = std::move(ow.ind);
            return;
        }
    }
    const int64_t p_global = sharded ? arr->global_rows : p;

    if (p_global >= n && (p >= n || sharded)) {
        // ---- tall: TSQR route (destroys its input, so work on a copy unless allowed)
        MatPtr work;
        T* y; int64_t ldy;
        if (input_is_conj_transposed) {
            work.reset(mat_conj_transpose<T>(c, arr));
            y = P<T>(work.get()); ldy = work->ld;
        } else if (may_destroy) {
            y = P<T>(arr); ldy = arr->ld;
        } else {
            work.reset(mat_clone<T>(c, arr));
            y = P<T>(work.get()); ldy = work->ld;
        }
        pqr_tall<T>(c, y, ldy, p, n, ncq, sharded, dtype, out);
        if (sharded) inherit_shard(out.q.get(), arr);
        return;
    }
    RC_REQUIRE(!sharded, "pivoted_qr: a row-sharded matrix must be tall (rows >= cols)");
    // ---- small general matrices: the fused one-CTA kernel reads the input as it is (row-major or conj-transposed)
    if (p * n <= 65536) {
        MatPtr r(mat_new(c, dtype, kk, n)), q(mat_new(c, dtype, p, ncq));
        DevBuf<int> dind(c, (size_t)n);
        if (pivqr_fused<T>(c, P<T>(arr), arr->ld, input_is_conj_transposed ? 2 : 0, p, n, ncq, P<T>(r.get()), r->ld, dind.p,
                           P<T>(q.get()), q->ld)) {
            if (out.want_ind) download_ind(c, dind.p, n, out.ind);
            out.q.reset(q.release());
            out.r.reset(r.release());
            return;
        }
    }
    // ---- general / short-wide: cooperative pivoted Householder on a column-major copy
    DevBuf<T> wc(c, (size_t)p * n);
    if (input_is_conj_transposed) {
        // column j of M = conj(row j of S): conj copy of S is already column-major M
        k_copy<T>(c, wc.p, p, P<T>(arr), arr->ld, n, p);
        k_conj_inplace<T>(c, wc.p, n, p, p);
    } else {
        k_transpose<T>(c, wc.p, p, P<T>(arr), arr->ld, p, n, false);
    }
    rc_trace(c, "  pivoted_qr: column-major copy");
    DevBuf<T> vbuf(c, (size_t)p * kk), tau(c, (size_t)kk);
    DevBuf<int> dind(c, (size_t)n);
    MatPtr r(mat_new(c, dtype, kk, n));
    pivqr_factor<T>(c, wc.p, p, p, n, P<T>(r.get()), r->ld, dind.p, vbuf.p, tau.p);
    rc_trace(c, "  pivoted_qr: factor + gather R");
    MatPtr q(mat_new(c, dtype, p, ncq));
    pivqr_form_q<T>(c, vbuf.p, tau.p, p, kk, ncq, P<T>(q.get()), q->ld);
    rc_trace(c, "  pivoted_qr: form Q");
    if (out.want_ind) download_ind(c, dind.p, n, out.ind);
    out.q.reset(q.release());
    out.r.reset(r.release());
}

// ============================================================================ SVD
// ComputeSVD::compute_svd (src/compute_svd.rs:14-30): thin SVD, s descending.
// Tall (or conj-transposed short-wide) input is reduced by TSQR to a w x w triangle, which a
// one-sided Jacobi kernel diagonalises: arr = (Q0 Ur) diag(s) W^H.
template <class T>
void svd_tall(rc_ctx* c, T* y, int64_t ldy, int64_t m, int64_t w, int dtype, bool sharded,
              MatPtr& u, std::vector<double>& s, MatPtr& wmat) {
    DevBuf<double> ds(c, (size_t)w);
    DevBuf<int> info(c, 2);
    MatPtr um(mat_new(c, dtype, m, w)), wm(mat_new(c, dtype, w, w));
    const int64_t wmax = tsqr_max_width(c, dtype);
    DevBuf<T> cq1, crinv2, crfac;
    int64_t lds = 0;
    if (m >= w && cholqr2<T>(c, y, ldy, m, w, sharded, dtype, cq1, crinv2, crfac, lds)) {
        const int64_t ldu = rc_pad_ld(dtype, w);
        DevBuf<T> ur(c, (size_t)w * ldu), tq(c, (size_t)w * ldu);
        jacobi_svd<T>(c, crfac.p, lds, w, w, ur.p, ldu, ds.p, P<T>(wm.get()), wm->ld, info.p);
        gemm<T>(c, RC_OP_N, RC_OP_N, w, w, w, crinv2.p, lds, ur.p, ldu, tq.p, ldu, rc_one<T>(), rc_zero<T>());
        gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, cq1.p, lds, tq.p, ldu, P<T>(um.get()), um->ld, rc_one<T>(), rc_zero<T>());
    } else if (w <= wmax && m >= w) {
        DistTsqr<T> ts;
        ts.factor(c, y, ldy, m, w, sharded);
        DevBuf<T> ur(c, (size_t)w * w);
        jacobi_svd<T>(c, ts.r(), w, w, w, ur.p, w, ds.p, P<T>(wm.get()), wm->ld, info.p);
        ts.apply(ur.p, w, w, P<T>(um.get()), um->ld);
    } else if (!sharded && c->svd_precondition) {
        // Wide panels (a general dense matrix, SVD::compute_from of src/svd.rs:165-169), preconditioned: column-pivoted
        // QR first (the GEMM-shaped large-dense route of pivoted_qr_impl), then one-sided Jacobi on R^H instead of on
        // an unpivoted triangle.  The rows of a rank-revealing R are graded and nearly orthogonal, so the iteration
        // converges in a few sweeps instead of ten or more (the preconditioning of LAPACK's ?gejsv; every sweep of the
        // whole-GPU kernel is n - 1 grid-wide rounds over an n x n matrix in L2):
        //   A P = Q R,  R^H = Uj S Wj^H  =>  A = (Q Wj) S (P Uj)^H.
        rc_matrix view;
        view.ctx = c; view.dtype = dtype; view.rows = m; view.cols = w; view.ld = ldy; view.data = y; view.owns = false;
        QrParts qr;
        qr.want_ind = true;
        pivoted_qr_impl<T>(c, &view, false, w, true, qr);
        MatPtr g(mat_conj_transpose<T>(c, qr.r.get()));                      // R^H, w x w lower triangular
        const int64_t ldu = rc_pad_ld(dtype, w);
        DevBuf<T> uj(c, (size_t)w * ldu), wj(c, (size_t)w * ldu);
        jacobi_svd<T>(c, P<T>(g.get()), g->ld, w, w, uj.p, ldu, ds.p, wj.p, ldu, info.p);
        gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, P<T>(qr.q.get()), qr.q->ld, wj.p, ldu, P<T>(um.get()), um->ld, rc_one<T>(), rc_zero<T>());
        // W = P Uj: row ind[i] of W is row i of Uj
        std::vector<int> inv((size_t)w);
        for (int64_t i = 0; i < w; ++i) inv[(size_t)qr.ind[(size_t)i]] = (int)i;
        DevBuf<int> dinv(c, (size_t)w);
        RC_CUDA(cudaMemcpyAsync(dinv.p, inv.data(), sizeof(int) * w, cudaMemcpyHostToDevice, c->stream));
        k_gather_rows<T>(c, P<T>(wm.get()), wm->ld, uj.p, ldu, w, w, dinv.p);
        RC_CUDA(cudaStreamSynchronize(c->stream));                            // (inv is read by the copy above)
    } else {
        // Wide panels, row-sharded input (or option "svd_precondition" = 0): unpivoted QR by column panels (all
        // GEMM-shaped), Jacobi on the w x w triangle -- one CTA if it fits shared memory, else the cooperative
        // whole-GPU kernel -- and U = Qfull Ur.
        DevBuf<T> qfull, r0;
        int64_t ldq = 0;
        panel_qr<T>(c, y, ldy, m, w, sharded, dtype, qfull, ldq, r0);
        const int64_t ldu = rc_pad_ld(dtype, w);
        DevBuf<T> ur(c, (size_t)w * ldu);
        jacobi_svd<T>(c, r0.p, w, w, w, ur.p, ldu, ds.p, P<T>(wm.get()), wm->ld, info.p);
        gemm<T>(c, RC_OP_N, RC_OP_N, m, w, w, qfull.p, ldq, ur.p, ldu, P<T>(um.get()), um->ld, rc_one<T>(), rc_zero<T>());
    }
    s.resize((size_t)w);
    int hinfo[2] = {0, 0};
    RC_CUDA(cudaMemcpyAsync(s.data(), ds.p, sizeof(double) * w, cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaMemcpyAsync(hinfo, info.p, sizeof(hinfo), cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    // ?gesdd reports non-convergence as info > 0, which the crate surfaces as LinalgError (src/compute_svd.rs:19-27)
    if (hinfo[0] != 0 && c->defer_depth == 0) RC_THROW(RC_LINALG_ERROR, "svd: the Jacobi iteration did not converge in %d sweeps", hinfo[1]);
    u.reset(um.release());
    wmat.reset(wm.release());
}

template <class T>
void svd_impl(rc_ctx* c, const rc_matrix* arr, bool input_is_conj_transposed, SvdParts& out) {
    const int dtype = arr->dtype;
    const int64_t m = input_is_conj_transposed ? arr->cols : arr->rows;
    const int64_t n = input_is_conj_transposed ? arr->rows : arr->cols;
    RC_REQUIRE(m > 0 && n > 0, "svd: empty matrix");
    const bool sharded = !input_is_conj_transposed && mat_sharded(arr);
    MatPtr uu, ww;
    if (m >= n) {
        // M = U S W^H directly
        MatPtr work(input_is_conj_transposed ? mat_conj_transpose<T>(c, arr) : mat_clone<T>(c, arr));
        svd_tall<T>(c, P<T>(work.get()), work->ld, m, n, dtype, sharded, uu, out.s, ww);
        out.u.reset(uu.release());
        if (sharded) inherit_shard(out.u.get(), arr);
        out.vt.reset(mat_conj_transpose<T>(c, ww.get()));
    } else {
        // M^H (n x m, tall) = Ub S Wb^H  =>  M = Wb S Ub^H : u = Wb, vt = Ub^H
        MatPtr work(input_is_conj_transposed ? mat_clone<T>(c, arr) : mat_conj_transpose<T>(c, arr));
        svd_tall<T>(c, P<T>(work.get()), work->ld, n, m, dtype, false, uu, out.s, ww);
        out.u.reset(ww.release());
        out.vt.reset(mat_conj_transpose<T>(c, uu.get()));
    }
}

// ============================================================================ algorithms
namespace {

template <class T>
rc_matrix* gaussian_new(rc_ctx* c, int dtype, int64_t rows, int64_t cols, uint64_t seed, uint32_t stream, int64_t row_offset) {
    MatPtr o(mat_new(c, dtype, rows, cols));
    k_gaussian<T>(c, P<T>(o.get()), rows, cols, o->ld, seed, stream, row_offset);
    return o.release();
}

// MatMat::matmat (src/types.rs:58-71)
template <class T>
rc_matrix* matmat_impl(rc_ctx* c, const rc_matrix* a, const rc_matrix* x) {
    RC_REQUIRE(a->cols == x->rows, "matmat: operator has %lld columns, X has %lld rows", (long long)a->cols, (long long)x->rows);
    if (a->op_matmat) {       // matrix-free operator: the caller's device callback (MatMat of src/types.rs:58-71)
        MatPtr y(mat_new(c, a->dtype, a->rows, x->cols));
        int rc = a->op_matmat(a->op_user, P<T>(x), x->ld, x->cols, y->data, y->ld, (void*)c->stream);
        if (rc != 0) RC_THROW(RC_LINALG_ERROR, "operator matmat callback failed (%d)", rc);
        c->launches++;
        inherit_shard(y.get(), a);
        return y.release();
    }
    MatPtr y(mat_mul<T>(c, RC_OP_N, a, RC_OP_N, x));
    inherit_shard(y.get(), a);
    return y.release();
}
// ConjMatMat::conj_matmat (src/types.rs:88-101, 129-131); partial sums across row shards.
template <class T>
rc_matrix* conj_matmat_impl(rc_ctx* c, const rc_matrix* a, const rc_matrix* x) {
    RC_REQUIRE(a->rows == x->rows, "conj_matmat: operator has %lld rows, X has %lld rows", (long long)a->rows, (long long)x->rows);
    MatPtr z;
    if (a->op_matmat) {       // ConjMatMat of src/types.rs:88-101 through the caller's callback
        RC_REQUIRE(a->op_conj_matmat, "operator has no conj_matmat callback");
        z.reset(mat_new(c, a->dtype, a->cols, x->cols));
        int rc = a->op_conj_matmat(a->op_user, P<T>(x), x->ld, x->cols, z->data, z->ld, (void*)c->stream);
        if (rc != 0) RC_THROW(RC_LINALG_ERROR, "operator conj_matmat callback failed (%d)", rc);
        c->launches++;
    } else {
        z.reset(mat_new(c, a->dtype, a->cols, x->cols));
        // the row padding (ld > cols: e.g. l = 74 in f32 gives ld = 76) travels through the all-reduce below, so it
        // is zeroed instead of reducing 8192 rows one collective at a time
        if (mat_sharded(a) && z->ld != z->cols)
            RC_CUDA(cudaMemsetAsync(z->data, 0, (size_t)z->rows * z->ld * rc_dtype_size(z->dtype), c->stream));
        gemm<T>(c, RC_OP_H, RC_OP_N, a->cols, x->cols, a->rows, P<T>(a), a->ld, P<T>(x), x->ld, P<T>(z.get()), z->ld,
                rc_one<T>(), rc_zero<T>());
    }
    if (mat_sharded(a)) {
        if (a->op_matmat && z->ld != z->cols) {
            // a callback wrote only the payload: pack, reduce once, unpack
            DevBuf<T> dense(c, (size_t)z->rows * z->cols);
            k_copy<T>(c, dense.p, z->cols, P<T>(z.get()), z->ld, z->rows, z->cols);
            comm_allreduce_sum(c, dense.p, (size_t)z->rows * z->cols, z->dtype);
            k_copy<T>(c, P<T>(z.get()), z->ld, dense.p, z->cols, z->rows, z->cols);
        } else {
            comm_allreduce_sum(c, z->data, (size_t)z->rows * z->ld, z->dtype);      // ONE all-reduce of the n x l partials
        }
    }
    return z.release();
}

// MaxColNorm::max_col_norm (src/random_sampling.rs:175-199)
template <class T>
double max_col_norm_impl(rc_ctx* c, const rc_matrix* m) {
    if (m->cols == 0) return 0.0;
    DevBuf<double> n2(c, (size_t)m->cols);
    k_col_norms2<T>(c, P<T>(m), m->ld, m->rows, m->cols, n2.p);
    if (mat_sharded(m)) comm_allreduce_sum(c, n2.p, (size_t)m->cols, RC_F64);
    std::vector<double> h((size_t)m->cols);
    RC_CUDA(cudaMemcpyAsync(h.data(), n2.p, sizeof(double) * m->cols, cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    double best = 0.0;
    for (double v : h) best = std::max(best, v);
    return std::sqrt(best);
}

// SampleRange::sample_range_by_rank (src/random_sampling.rs:103-118)
template <class T>
rc_matrix* sample_by_rank_impl(rc_ctx* c, const rc_matrix* a, int64_t k, int64_t p, const rc_matrix* omega, uint64_t seed) {
    const int64_t n = a->cols, l = k + p;
    RC_REQUIRE(k > 0 && p >= 0, "sample_range_by_rank: need k > 0, p >= 0");
    MatPtr gen;
    if (!omega) { gen.reset(gaussian_new<T>(c, a->dtype, n, l, seed, 0, 0)); omega = gen.get(); }
    RC_REQUIRE(omega->rows == n && omega->cols == l, "omega must be %lld x %lld", (long long)n, (long long)l);
    rc_trace(c, nullptr);
    MatPtr y(matmat_impl<T>(c, a, omega));                       // :111-112
    rc_trace(c, "by_rank: Y = A Omega");
    QrParts qr;
    qr.want_ind = false;
    int64_t kk = std::min<int64_t>(mat_sharded(a) ? a->global_rows : a->rows, l);
    pivoted_qr_impl<T>(c, y.get(), false, std::min(k, kk), true, qr);   // :114-115 compress(RANK(k))
    return qr.q.release();
}

// SampleRangePowerIteration::sample_range_power_iteration (src/random_sampling.rs:131-160).
// Reference semantics (quirk Q1): `op_omega` inside the loop is a fresh binding, so every trip
// restarts from Y0 = A Omega and the result equals the it_count = 1 result; every trip is
// executed as the reference executes it.  Option "true_power_iteration" advances Y instead.
template <class T>
rc_matrix* sample_power_once(rc_ctx* c, const rc_matrix* a, int64_t k, int64_t p, int64_t it_count,
                             const rc_matrix* omega) {
    const int64_t l = k + p;
    MatPtr y0(matmat_impl<T>(c, a, omega));                      // :140-141
    MatPtr res;                                                  // :142 (res = y0 unless a trip replaces it)
    if (c->true_power_iteration || it_count < 2 || !c->overlap) {
        MatPtr cur;                                              // only for true_power_iteration
        for (int64_t index = 0; index < it_count; ++index) {
            const rc_matrix* start = (c->true_power_iteration && cur.get()) ? cur.get() : y0.get();
            QrParts q1; q1.want_ind = false;
            pivoted_qr_impl<T>(c, start, false, -1, false, q1);      // :145-146 (all columns of Q)
            if (index == 0 && c->defer_depth > 0 && !check_deferred_now(c)) {   // (see the two-stream branch below)
                RedoScope redo(c, a->dtype);
                pivoted_qr_impl<T>(c, start, false, -1, false, q1);
            }
            MatPtr z(conj_matmat_impl<T>(c, a, q1.q.get()));          // :148
            QrParts q2; q2.want_ind = false;
            pivoted_qr_impl<T>(c, z.get(), false, -1, true, q2);     // :148-149
            MatPtr ynew(matmat_impl<T>(c, a, q2.q.get()));            // :150
            if (c->true_power_iteration) { cur.reset(ynew.release()); if (index == it_count - 1) res.reset(cur.release()); }
            else if (index == it_count - 1) res.reset(ynew.release());   // :151-153
        }
    } else {
        // Reference semantics (quirk Q1): every trip starts from the same Y0 and none reads another's result, so the
        // trips are independent.  They are executed -- all of them, as the reference executes them -- on two auxiliary
        // streams, issued stage by stage in lockstep (the order of the collectives is then the same on every rank of a
        // sharded run): while one trip sits in the latency-bound one-CTA kernels of its pivoted QR (Cholesky, pivoting,
        // forming Q), the other trip's kernels fill the rest of the machine.
        struct Trip { QrParts q1, q2; MatPtr z, ynew; };
        std::vector<Trip> trips((size_t)it_count);
        cudaEvent_t ev_y0 = aux_event(c, 0);
        RC_CUDA(cudaEventRecord(ev_y0, c->stream));
        for (int stage = 0; stage < 4; ++stage) {
            for (int64_t t = 0; t < it_count; ++t) {
                Trip& tr = trips[(size_t)t];
                StreamScope on_aux(c, (int)(t & 1));
                if (stage == 0 && t < 2) RC_CUDA(cudaStreamWaitEvent(c->stream, ev_y0, 0));
                switch (stage) {
                    case 0: tr.q1.want_ind = false; pivoted_qr_impl<T>(c, y0.get(), false, -1, false, tr.q1); break;   // :145-146
                    case 1: tr.z.reset(conj_matmat_impl<T>(c, a, tr.q1.q.get())); break;                              // :148
                    case 2: tr.q2.want_ind = false; pivoted_qr_impl<T>(c, tr.z.get(), false, -1, true, tr.q2); break; // :148-149
                    default: tr.ynew.reset(matmat_impl<T>(c, a, tr.q2.q.get())); break;                               // :150
                }
            }
            if (stage == 0 && c->defer_depth > 0 && !check_deferred_now(c)) {
                // Y0 = A Omega is the one sketch of the pipeline whose columns are not graded: when the operator's
                // spectrum falls by more than ~7 decades over the sketch width its Gram matrix is numerically singular
                // and the plain Cholesky-QR2 is rejected.  The later sketches (Z = A^H q, Y = A w) inherit the column
                // order of a pivoted QR, are graded like the spectrum and pass (cholqr2_acceptable, rule (b)).  So the
                // first factorisation is checked here -- one host round trip, ~30 us, at a point where nothing else can
                // run anyway -- and redone at once on the shifted route, instead of finding out at the end of the
                // pipeline and running all of it again (measured on a twelve-decade sketch at configs[1] size: 37 ms
                // with the Householder re-run of the whole pipeline, 24 ms on Householder TSQR throughout).
                RedoScope redo(c, a->dtype);
                for (int64_t t = 0; t < it_count; ++t) {
                    StreamScope on_aux(c, (int)(t & 1));
                    pivoted_qr_impl<T>(c, y0.get(), false, -1, false, trips[(size_t)t].q1);
                }
            }
        }
        for (int s = 0; s < 2; ++s) {                            // join: everything below (and every free) follows both
            cudaEvent_t ev = aux_event(c, 1 + s);
            RC_CUDA(cudaEventRecord(ev, c->aux_stream[s]));
            RC_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
        }
        res.reset(trips[(size_t)it_count - 1].ynew.release());   // :151-153 (the last trip's product)
    }
    const rc_matrix* fin = res.get() ? res.get() : y0.get();
    QrParts qr; qr.want_ind = false;
    int64_t kk = std::min<int64_t>(mat_sharded(a) ? a->global_rows : a->rows, l);
    pivoted_qr_impl<T>(c, fin, false, std::min(k, kk), true, qr);   // :156-159
    return qr.q.release();
}

template <class T>
rc_matrix* sample_power_impl(rc_ctx* c, const rc_matrix* a, int64_t k, int64_t p, int64_t it_count,
                             const rc_matrix* omega, uint64_t seed) {
    const int64_t n = a->cols, l = k + p;
    RC_REQUIRE(k > 0 && p >= 0 && it_count >= 0, "sample_range_power_iteration: bad arguments");
    MatPtr gen;
    if (!omega) { gen.reset(gaussian_new<T>(c, a->dtype, n, l, seed, 0, 0)); omega = gen.get(); }
    RC_REQUIRE(omega->rows == n && omega->cols == l, "omega must be %lld x %lld", (long long)n, (long long)l);
    if (!c->speculate || c->qr_mode == 1 || c->defer_depth > 0) return sample_power_once<T>(c, a, k, p, it_count, omega);
    // Speculative pass: no host synchronisation inside the pipeline (the Cholesky-QR2 panels are assumed acceptable,
    // pivot vectors are not downloaded); one synchronisation at the end checks every panel's status word.  A rejected
    // panel re-runs the sampler, first (double precision) with every tall QR on the shifted Cholesky-QR3 -- still
    // speculative, still GEMM-shaped -- and then on the unconditionally stable Householder TSQR.
    auto speculative_pass = [&](MatPtr& q) -> bool {
        bool ok = false;
        try {
            DeferScope defer(c);
            q.reset(sample_power_once<T>(c, a, k, p, it_count, omega));
            ok = true;
        } catch (...) {
            drop_deferred(c);
            throw;
        }
        if (ok && finish_deferred(c)) return true;
        q.reset(nullptr);
        return false;
    };
    MatPtr q;
    if (speculative_pass(q)) return q.release();
    const bool single_prec = (a->dtype == RC_F32 || a->dtype == RC_C32);
    if (c->shifted_cholqr && !single_prec) {
        c->force_shifted = true;
        bool ok = false;
        try { ok = speculative_pass(q); } catch (...) { c->force_shifted = false; throw; }
        c->force_shifted = false;
        if (ok) return q.release();
    }
    c->force_householder = true;
    try {
        q.reset(sample_power_once<T>(c, a, k, p, it_count, omega));
    } catch (...) { c->force_householder = false; throw; }
    c->force_householder = false;
    return q.release();
}

// AdaptiveSampling::sample_range_adaptive (src/random_sampling.rs:223-274).  Q and B live in
// pre-allocated device buffers with a column / row cursor instead of the crate's re-concatenation.
template <class T>
rc_matrix* sample_adaptive_impl(rc_ctx* c, const rc_matrix* a, double rel_tol_in, int64_t s, const rc_matrix* omega_blocks,
                                uint64_t seed, int64_t max_rank, std::vector<uint64_t>& hist_rank, std::vector<double>& hist_res) {
    using R = RealOf<T>;
    const int64_t m = a->rows, n = a->cols;
    const int64_t m_glob = mat_sharded(a) ? a->global_rows : m;
    RC_REQUIRE(s > 0, "sample_range_adaptive: sample_size must be positive");
    if (max_rank <= 0) max_rank = std::min(m_glob, n);
    const int64_t cap = (max_rank + s - 1) / s * s + s;
    const R tol_factor = (R)(10.0 * std::sqrt(2.0 / 3.14159265358979323846));   // :229-232
    const R rel_tol = (R)rel_tol_in;
    uint32_t draw = 0;
    auto next_omega = [&](MatPtr& holder) -> const rc_matrix* {
        if (omega_blocks) {
            RC_REQUIRE((int64_t)(draw + 1) * s <= omega_blocks->cols, "omega_blocks exhausted after %u draws", draw);
            holder.reset(mat_slice<T>(c, omega_blocks, 0, n, (int64_t)draw * s, (int64_t)(draw + 1) * s));
        } else {
            holder.reset(gaussian_new<T>(c, a->dtype, n, s, seed, draw, 0));
        }
        ++draw;
        return holder.get();
    };
    if (omega_blocks) RC_REQUIRE(omega_blocks->rows == n, "omega_blocks must have %lld rows", (long long)n);
    MatPtr om;
    const rc_matrix* omega = next_omega(om);                                     // :238
    MatPtr y(matmat_impl<T>(c, a, omega));                                       // :239
    const R operator_norm = (R)max_col_norm_impl<T>(c, y.get()) * tol_factor;    // :241
    R max_norm = operator_norm;
    // Q (m x cur) and B (cur x n) grow geometrically (the rank is not known in advance and the crate's
    // only bound is min(m, n): pre-allocating that would be two A-sized buffers)
    int64_t cur = std::min<int64_t>(cap, std::max<int64_t>(8 * s, 512));
    MatPtr qbuf(mat_new(c, a->dtype, m, cur)), bbuf(mat_new(c, a->dtype, cur, n));
    int64_t r = 0;
    while (max_norm / operator_norm >= rel_tol) {                                // :248
        if (r + s > max_rank) RC_THROW(RC_COMPRESSION_ERROR, "adaptive sampler exceeded max_rank %lld", (long long)max_rank);
        if (r + s > cur) {
            const int64_t grown = std::min<int64_t>(cap, std::max<int64_t>(2 * cur, r + s));
            MatPtr q2(mat_new(c, a->dtype, m, grown)), b2(mat_new(c, a->dtype, grown, n));
            k_copy<T>(c, P<T>(q2.get()), q2->ld, P<T>(qbuf.get()), qbuf->ld, m, r);
            k_copy<T>(c, P<T>(b2.get()), b2->ld, P<T>(bbuf.get()), bbuf->ld, r, n);
            qbuf.reset(q2.release()); bbuf.reset(b2.release());
            cur = grown;
        }
        rc_trace(c, nullptr);
        // The next sketch Y' = A Omega' (:265-266) depends on nothing this trip computes: it is drawn and launched now, on
        // an auxiliary stream with all but `side_sms` SMs, while the projection and the pivoted QR of the current sketch
        // -- a chain of small, latency-bound kernels that could not use the machine anyway -- run on the SMs left over.
        // Same draws in the same order, same products; only the order of issue changes (the reference computes A Omega'
        // at the end of the trip whatever the trip finds).
        // (single precision only: the tcgen05 kernels are one CTA per SM, so SMs left out of their grid are EMPTY and can
        // host the one-CTA kernels of the chain -- chol_inv_kernel 1024 threads x 64 registers, pivqr_fused_kernel
        // 512 x 90, neither fits beside anything.  The DMMA products of f64 / c64 run two CTAs per SM, a smaller grid is
        // placed breadth-first and leaves no SM empty, the chain waits for the product after all: measured neutral,
        // 15.0 / 15.1 ms at f64 16384^2, profiles/r2_side_sms_ab.txt.)
        const bool side = c->overlap && c->side_sms > 0 && !c->trace && !a->op_matmat && c->sm_count >= 8 * c->side_sms &&
                          (a->dtype == RC_F32 || a->dtype == RC_C32);
        // How many SMs the chain gets.  The tcgen05 kernels (f32 / c32) hand 128-row work items to one persistent CTA per
        // SM in rounds, and a CTA is rate-limited by its own ingest, so the product takes
        // ceil(items / CTAs) x the time of one item: at config 3, 256 items take two rounds on 148 SMs and two rounds on
        // 128 -- 20 SMs cost nothing (measured: 15.3 ms without the overlap, 14.5 / 13.8 ms with 8 / 16 SMs, 15.2 ms
        // with 24 or 32, which push the product into a third round).  So the share is raised from `side_sms` up to three
        // times that as long as the number of rounds stays what it is on the whole device.  The DMMA kernels (f64 / c64)
        // draw 64-row tiles from a counter and slow down in proportion: they keep the minimum.
        int side_n = c->side_sms;
        if (side) {
            const int64_t cols = (a->dtype == RC_C32 ? 2 : 1) * s;
            const int64_t items = ((m + 127) / 128) * ((cols + 95) / 96);
            const int64_t rounds_full = (items + c->sm_count - 1) / c->sm_count;
            for (int cand = 3 * c->side_sms; cand > c->side_sms; cand -= 2)
                if ((items + (c->sm_count - cand) - 1) / (c->sm_count - cand) == rounds_full) { side_n = cand; break; }
        }
        MatPtr om_next, y_next;
        if (side) {
            next_omega(om_next);                                                 // :265
            cudaEvent_t ev_in = aux_event(c, 0), ev_out = aux_event(c, 1);
            RC_CUDA(cudaEventRecord(ev_in, c->stream));
            try {
                StreamScope on_aux(c, 0);
                SmBudget budget(c, c->sm_count - side_n);
                RC_CUDA(cudaStreamWaitEvent(c->stream, ev_in, 0));
                y_next.reset(matmat_impl<T>(c, a, om_next.get()));               // :266  A Omega'
                RC_CUDA(cudaEventRecord(ev_out, c->stream));
            } catch (...) {
                if (c->aux_stream[0]) cudaStreamSynchronize(c->aux_stream[0]);
                throw;
            }
        }
        QrParts qq;
        qq.want_ind = false;                 // (the pivots of the sketch are not used: no download, no host synchronisation)
        try {
            SmBudget budget(c, side ? side_n : c->sm_count);
            if (r > 0) {                                                         // :250-252
                DevBuf<T> t(c, (size_t)r * s);
                gemm<T>(c, RC_OP_H, RC_OP_N, r, s, m, P<T>(qbuf.get()), qbuf->ld, P<T>(y.get()), y->ld, t.p, s, rc_one<T>(), rc_zero<T>());
                if (mat_sharded(a)) comm_allreduce_sum(c, t.p, (size_t)r * s, a->dtype);
                gemm<T>(c, RC_OP_N, RC_OP_N, m, s, r, P<T>(qbuf.get()), qbuf->ld, t.p, s, P<T>(y.get()), y->ld, rc_make<T>(-1.0, 0.0), rc_one<T>());
            }
            rc_trace(c, "adaptive: Y -= Q (Q^H Y)");
            pivoted_qr_impl<T>(c, y.get(), false, -1, true, qq);                 // :254
            rc_trace(c, "adaptive: pivoted QR of Y");
        } catch (...) {
            if (side && c->aux_stream[0]) cudaStreamSynchronize(c->aux_stream[0]);   // Y' is still being written
            throw;
        }
        if (side) RC_CUDA(cudaStreamWaitEvent(c->stream, aux_event(c, 1), 0));
        const int64_t sq = qq.q->cols;
        MatPtr z(conj_matmat_impl<T>(c, a, qq.q.get()));                          // :256-260  (A^H q)
        rc_trace(c, "adaptive: A^H q");
        k_transpose<T>(c, P<T>(bbuf.get()) + r * bbuf->ld, bbuf->ld, P<T>(z.get()), z->ld, n, sq, true);
        k_copy<T>(c, P<T>(qbuf.get()) + r, qbuf->ld, P<T>(qq.q.get()), qq.q->ld, m, sq);   // :262
        r += sq;
        if (side) {
            om.reset(om_next.release());
            omega = om.get();
            y.reset(y_next.release());
        } else {
            omega = next_omega(om);                                              // :265
            rc_trace(c, "adaptive: append B, Q, draw Omega");
            y.reset(matmat_impl<T>(c, a, omega));                                // :266  A Omega
            rc_trace(c, "adaptive: A Omega");
        }
        {
            DevBuf<T> t(c, (size_t)r * s);
            gemm<T>(c, RC_OP_N, RC_OP_N, r, s, n, P<T>(bbuf.get()), bbuf->ld, P<T>(omega), omega->ld, t.p, s, rc_one<T>(), rc_zero<T>());
            gemm<T>(c, RC_OP_N, RC_OP_N, m, s, r, P<T>(qbuf.get()), qbuf->ld, t.p, s, P<T>(y.get()), y->ld, rc_make<T>(-1.0, 0.0), rc_one<T>());
        }
        rc_trace(c, "adaptive: Y -= Q (B Omega)");
        max_norm = (R)max_col_norm_impl<T>(c, y.get()) * tol_factor;             // :269
        rc_trace(c, "adaptive: max_col_norm");
        hist_rank.push_back((uint64_t)r);
        hist_res.push_back((double)(max_norm / operator_norm));                  // :270
    }
    MatPtr q(mat_slice<T>(c, qbuf.get(), 0, m, 0, r));
    inherit_shard(q.get(), a);
    if (a->owns) {      // B = Q^H A is a by-product: keep it with Q for compute_from_range_estimate
        q->companion = mat_slice<T>(c, bbuf.get(), 0, r, 0, n);
        q->companion_op_id = a->id;
    }
    return q.release();
}

// A^H Q (n x k) for a range estimate Q: the conj_matmat of src/qr.rs:315 / src/svd.rs:175, or -- when Q
// came out of the adaptive sampler on this very operator -- the transpose of the B = Q^H A it built.
template <class T>
rc_matrix* ah_range(rc_ctx* c, const rc_matrix* op, const rc_matrix* range) {
    const rc_matrix* b = range->companion;
    if (c->reuse_range_b && b && op->owns && range->companion_op_id == op->id && b->rows == range->cols &&
        b->cols == op->cols && b->dtype == op->dtype) {
        c->range_b_reused++;
        return mat_conj_transpose<T>(c, b);
    }
    return conj_matmat_impl<T>(c, op, range);
}

// QRTraits::compute_from_range_estimate (src/qr.rs:311-323)
template <class T>
void qr_from_range_impl(rc_ctx* c, const rc_matrix* range, const rc_matrix* op, QrParts& out) {
    RC_REQUIRE(range->rows == op->rows, "range estimate and operator row counts differ");
    rc_trace(c, nullptr);
    MatPtr z(ah_range<T>(c, op, range));                         // A^H Q  (n x k)
    rc_trace(c, "qr_from_range: A^H Q");
    QrParts qb;
    pivoted_qr_impl<T>(c, z.get(), true, -1, false, qb);         // pivoted QR of b = (A^H Q)^H, :315-316
    rc_trace(c, "qr_from_range: pivoted QR of b");
    out.q.reset(mat_mul<T>(c, RC_OP_N, range, RC_OP_N, qb.q.get()));   // :319
    rc_trace(c, "qr_from_range: Q q_b");
    inherit_shard(out.q.get(), range);
    out.r.reset(qb.r.release());
    out.ind = qb.ind;
}

// SVDTraits::compute_from_range_estimate (src/svd.rs:171-183)
template <class T>
void svd_from_range_impl(rc_ctx* c, const rc_matrix* range, const rc_matrix* op, SvdParts& out) {
    RC_REQUIRE(range->rows == op->rows, "range estimate and operator row counts differ");
    MatPtr z(ah_range<T>(c, op, range));                         // A^H Q = b^H  (n x k)
    SvdParts sb;
    svd_impl<T>(c, z.get(), true, sb);                           // SVD of b, :175-176
    out.u.reset(mat_mul<T>(c, RC_OP_N, range, RC_OP_N, sb.u.get()));   // :179
    inherit_shard(out.u.get(), range);
    out.vt.reset(sb.vt.release());
    out.s = sb.s;
}

template <class T>
rc_matrix* permute_impl(rc_ctx* c, const rc_matrix* m, const std::vector<uint64_t>& idx, int mode) {
    // src/permutation.rs:84-144
    const bool cols = (mode == RC_PERM_COL || mode == RC_PERM_COLINV);
    RC_REQUIRE((int64_t)idx.size() == (cols ? m->cols : m->rows),
               cols ? "Length of index array and number of columns differ." : "Length of index array and number of rows differ.");
    std::vector<uint64_t> use = (mode == RC_PERM_COLINV || mode == RC_PERM_ROWINV) ? invert_perm(idx) : idx;
    for (uint64_t v : use) RC_REQUIRE(v < use.size(), "index out of range in permutation");
    DevIndex di(c, use);
    MatPtr o(mat_new(c, m->dtype, m->rows, m->cols));
    if (cols) k_gather_cols<T>(c, P<T>(o.get()), o->ld, P<T>(m), m->ld, m->rows, m->cols, di.d.p);
    else k_gather_rows<T>(c, P<T>(o.get()), o->ld, P<T>(m), m->ld, m->rows, m->cols, di.d.p);
    return o.release();
}

// QRTraits::column_id (src/qr.rs:270-309)
template <class T>
void column_id_impl(rc_ctx* c, const rc_matrix* q, const rc_matrix* r, const std::vector<uint64_t>& ind, rc_column_id* out) {
    const int64_t rank = q->cols, ncols = r->cols;
    MatPtr cm, z(mat_new(c, q->dtype, rank, ncols));
    if (rank == ncols) {
        cm.reset(mat_mul<T>(c, RC_OP_N, q, RC_OP_N, r));                         // :277
        k_eye<T>(c, P<T>(z.get()), rank, ncols, z->ld);
    } else {
        MatPtr r11(mat_slice<T>(c, r, 0, rank, 0, rank));                        // :286
        cm.reset(mat_mul<T>(c, RC_OP_N, q, RC_OP_N, r11.get()));                 // :287
        k_eye<T>(c, P<T>(z.get()), rank, ncols, z->ld);                          // [I | .]
        // all n - k triangular solves at once (:290-301)
        trsm_upper<T>(c, P<T>(r11.get()), r11->ld, false, rank, P<T>(r) + rank, r->ld, ncols - rank,
                      P<T>(z.get()) + rank, z->ld);
    }
    MatPtr zp(permute_impl<T>(c, z.get(), ind, RC_PERM_COLINV));                 // :278, :305
    inherit_shard(cm.get(), q);
    out->c = cm.release();
    out->z = zp.release();
    out->col_ind = ind;
}

// LQTraits::row_id (src/qr.rs:363-403)
template <class T>
void row_id_impl(rc_ctx* c, const rc_matrix* l, const rc_matrix* q, const std::vector<uint64_t>& ind, rc_row_id* out) {
    const int64_t rank = q->rows, nrows = l->rows;
    MatPtr rm, x(mat_new(c, l->dtype, nrows, rank));
    if (rank == nrows) {
        rm.reset(mat_mul<T>(c, RC_OP_N, l, RC_OP_N, q));                         // :371
        k_eye<T>(c, P<T>(x.get()), nrows, rank, x->ld);
    } else {
        MatPtr l11(mat_slice<T>(c, l, 0, rank, 0, rank));                        // :379
        rm.reset(mat_mul<T>(c, RC_OP_N, l11.get(), RC_OP_N, q));                 // :380
        k_eye<T>(c, P<T>(x.get()), nrows, rank, x->ld);
        // X2 L11 = L21  <=>  L11^T X2^T = L21^T : upper solve with the plain transpose (:383-395)
        const int64_t nr = nrows - rank;
        DevBuf<T> bt(c, (size_t)rank * nr), xt(c, (size_t)rank * nr);
        k_transpose<T>(c, bt.p, nr, P<T>(l) + rank * l->ld, l->ld, nr, rank, false);
        trsm_upper<T>(c, P<T>(l11.get()), l11->ld, true, rank, bt.p, nr, nr, xt.p, nr);
        k_transpose<T>(c, P<T>(x.get()) + rank * x->ld, x->ld, xt.p, nr, rank, nr, false);
    }
    MatPtr xp(permute_impl<T>(c, x.get(), ind, RC_PERM_ROWINV));                 // :369, :399
    out->x = xp.release();
    out->r = rm.release();
    out->row_ind = ind;
}

template <class T>
double rel_diff_impl(rc_ctx* c, const rc_matrix* first, const rc_matrix* second) {
    // src/types.rs:182-188, 190-196
    RC_REQUIRE(first->rows == second->rows && first->cols == second->cols, "rel_diff: shapes differ");
    DevBuf<double> d(c, 2);
    k_diff_fro2<T>(c, P<T>(first), first->ld, P<T>(second), second->ld, first->rows, first->cols, d.p);
    k_fro2<T>(c, P<T>(second), second->ld, second->rows, second->cols, d.p + 1);
    if (mat_sharded(second)) comm_allreduce_sum(c, d.p, 2, RC_F64);
    double h[2];
    RC_CUDA(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    return std::sqrt(h[0]) / std::sqrt(h[1]);
}

// RandomMatrix::random_orthogonal_matrix (src/random_matrix.rs:35-56): left singular vectors of a Gaussian.
template <class T>
rc_matrix* random_orthogonal_impl(rc_ctx* c, int dtype, int64_t rows, int64_t cols, uint64_t seed, uint32_t stream) {
    int64_t m = rows, n = cols;
    const bool swap = cols > rows;
    if (swap) std::swap(m, n);
    MatPtr g(gaussian_new<T>(c, dtype, m, n, seed, stream, 0));
    SvdParts sv;
    svd_impl<T>(c, g.get(), false, sv);
    if (swap) return mat_conj_transpose<T>(c, sv.u.get());
    return sv.u.release();
}

template <class T>
rc_matrix* low_rank_from_factors(rc_ctx* c, const rc_matrix* u, const std::vector<double>& sig, const rc_matrix* vt) {
    using R = RealOf<T>;
    std::vector<R> hs(sig.size());
    for (size_t i = 0; i < sig.size(); ++i) hs[i] = (R)sig[i];
    DevBuf<R> ds(c, hs.size());
    RC_CUDA(cudaMemcpyAsync(ds.p, hs.data(), sizeof(R) * hs.size(), cudaMemcpyHostToDevice, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    MatPtr sv(mat_clone<T>(c, vt));
    k_scale_rows<T>(c, P<T>(sv.get()), sv->rows, sv->cols, sv->ld, ds.p);
    return mat_mul<T>(c, RC_OP_N, u, RC_OP_N, sv.get());
}

std::vector<uint64_t> vec_from(const uint64_t* p, size_t n) {
    RC_REQUIRE(p || n == 0, "null index array");
    return std::vector<uint64_t>(p, p + n);
}
void copy_ind(const std::vector<uint64_t>& v, uint64_t* out, size_t n) {
    RC_REQUIRE(out && n >= v.size(), "index output buffer too small (%zu < %zu)", n, v.size());
    memcpy(out, v.data(), v.size() * sizeof(uint64_t));
}

}  // namespace

template <class T>
static rc_matrix* scaled_vt(rc_ctx* c, const rc_svd* svd) {
    using R = RealOf<T>;
    std::vector<R> hs(svd->s.size());
    for (size_t i = 0; i < hs.size(); ++i) hs[i] = (R)svd->s[i];
    DevBuf<R> ds(c, std::max<size_t>(hs.size(), 1));
    if (!hs.empty()) {
        RC_CUDA(cudaMemcpyAsync(ds.p, hs.data(), sizeof(R) * hs.size(), cudaMemcpyHostToDevice, c->stream));
        RC_CUDA(cudaStreamSynchronize(c->stream));
    }
    MatPtr sv(mat_clone<T>(c, svd->vt));
    k_scale_rows<T>(c, P<T>(sv.get()), sv->rows, sv->cols, sv->ld, ds.p);
    return sv.release();
}
template <class T>
static rc_matrix* chain_apply(rc_ctx* c, std::initializer_list<const rc_matrix*> factors, const rc_matrix* rhs) {
    // right-to-left: f0 (f1 (... rhs))   (src/col_interp_decomp.rs:141, src/two_sided_interp_decomp.rs:160)
    MatPtr cur;
    const rc_matrix* x = rhs;
    std::vector<const rc_matrix*> fs(factors);
    for (size_t i = fs.size(); i-- > 0;) {
        MatPtr nxt(mat_mul<T>(c, RC_OP_N, fs[i], RC_OP_N, x));
        cur.reset(nxt.release());
        x = cur.get();
    }
    return cur.release();
}

// ============================================================================ extern "C"
extern "C" {

int rc_version(void) { return 100; }

rc_status rc_ctx_create(int device, rc_ctx** out) {
    if (!out) return RC_INVALID_ARGUMENT;
    *out = nullptr;
    rc_ctx* c = new rc_ctx();
    int caller_device = -1;
    cudaGetDevice(&caller_device);
    rc_status st = guard(c, [&] {
        int ndev = 0;
        RC_CUDA(cudaGetDeviceCount(&ndev));
        RC_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d visible)", device, ndev);
        RC_CUDA(cudaSetDevice(device));
        c->device = device;
        cudaDeviceProp prop;
        RC_CUDA(cudaGetDeviceProperties(&prop, device));
        RC_REQUIRE(prop.major == 10, "this library is built for sm_100a only (device is sm_%d%d)", prop.major, prop.minor);
        c->sm_count = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        RC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
        // keep freed blocks in the stream-ordered pool: workspaces are re-used every call
        cudaMemPool_t pool;
        RC_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = ~0ull;
        RC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        // memory freed on one stream is not handed to another stream by making it WAIT for the free (the pool's default):
        // that orders the small-kernel chain on one stream behind the big product on the other (measured: the overlap of
        // the adaptive sampler did not happen at all); reuse is opportunistic -- only once the free has completed
        int off = 0;
        RC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
    });
    if (caller_device >= 0 && caller_device != device) cudaSetDevice(caller_device);   // leave the caller's device as it was
    if (st != RC_OK) { fprintf(stderr, "rc_ctx_create: %s\n", c->err.c_str()); delete c; return st; }
    *out = c;
    return RC_OK;
}

rc_status rc_ctx_destroy(rc_ctx* c) {
    if (!c) return RC_OK;
    {
        DeviceGuard dg(c->device);
        cudaStreamSynchronize(c->stream);
        rc_cache_release(c);
        cudaStreamSynchronize(c->stream);
        for (cudaEvent_t ev : c->event_pool) cudaEventDestroy(ev);
        c->event_pool.clear();
        try { comm_destroy(c); } catch (...) {}
        if (c->tile_counter) cudaFree(c->tile_counter);
        for (int i = 0; i < 2; ++i) {
            if (c->aux_stream[i]) { cudaStreamSynchronize(c->aux_stream[i]); cudaStreamDestroy(c->aux_stream[i]); }
            if (c->aux_tile_counter[i]) cudaFree(c->aux_tile_counter[i]);
        }
        for (int i = 0; i < 4; ++i) if (c->aux_event[i]) cudaEventDestroy(c->aux_event[i]);
        if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
        if (c->own_stream) cudaStreamDestroy(c->stream);
    }
    delete c;
    return RC_OK;
}

rc_status rc_ctx_set_stream(rc_ctx* c, void* s) {
    if (!c) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        rc_cache_release(c);                         // cached workspaces are tied to the stream they were freed on
        RC_CUDA(cudaStreamSynchronize(c->stream));
        if (c->own_stream) RC_CUDA(cudaStreamDestroy(c->stream));
        c->stream = (cudaStream_t)s;
        c->own_stream = false;
    });
}
rc_status rc_ctx_synchronize(rc_ctx* c) {
    if (!c) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_CUDA(cudaStreamSynchronize(c->stream)); });
}
const char* rc_last_error_string(rc_ctx* c) { return c ? c->err.c_str() : "null context"; }

rc_status rc_ctx_set_option(rc_ctx* c, const char* key, int64_t v) {
    if (!c || !key) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (!strcmp(key, "gemm_impl")) c->gemm_impl = (int)v;
        else if (!strcmp(key, "dmma_tail")) c->dmma_tail = (int)v;
        else if (!strcmp(key, "tf32_ring")) c->tf32_ring = (int)v;
        else if (!strcmp(key, "f32_precision")) { RC_REQUIRE(v == 0 || v == 1, "f32_precision: 0 (3xTF32) or 1 (bf16)"); c->f32_precision = (int)v; }
        else if (!strcmp(key, "true_power_iteration")) c->true_power_iteration = (int)v;
        else if (!strcmp(key, "qr_mode")) c->qr_mode = (int)v;
        else if (!strcmp(key, "pivot_precision")) c->pivot_f64 = (v != 0);
        else if (!strcmp(key, "speculate")) c->speculate = (int)v;
        else if (!strcmp(key, "shifted_cholqr")) c->shifted_cholqr = (int)v;
        else if (!strcmp(key, "fused_small_qr")) c->fused_small_qr = (int)v;
        else if (!strcmp(key, "cluster_qr")) c->cluster_qr = (int)v;
        else if (!strcmp(key, "svd_precondition")) c->svd_precondition = (int)v;
        else if (!strcmp(key, "workspace_cache")) { c->block_cache_on = (int)v; if (!v) rc_cache_release(c); }
        else if (!strcmp(key, "release_workspaces")) { trim_pool(c); }
        else if (!strcmp(key, "overlap")) c->overlap = (int)v;
        else if (!strcmp(key, "side_sms")) { RC_REQUIRE(v >= 0 && v <= 16, "side_sms: 0 (off) .. 16"); c->side_sms = (int)v; }
        else if (!strcmp(key, "reuse_range_b")) c->reuse_range_b = (int)v;
        else if (!strcmp(key, "trace")) c->trace = (int)v;
        else RC_THROW(RC_INVALID_ARGUMENT, "unknown option '%s'", key);
    });
}
rc_status rc_ctx_get_counter(rc_ctx* c, const char* key, int64_t* out) {
    if (!c || !key || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (!strcmp(key, "kernel_launches")) *out = c->launches;
        else if (!strcmp(key, "gemm_flops")) *out = c->gemm_flops;
        else if (!strcmp(key, "h2d_bytes")) *out = c->h2d_bytes;
        else if (!strcmp(key, "d2h_bytes")) *out = c->d2h_bytes;
        else if (!strcmp(key, "cholqr_used")) *out = c->cholqr_used;
        else if (!strcmp(key, "cholqr_fallbacks")) *out = c->cholqr_fallbacks;
        else if (!strcmp(key, "cholqr_shifted")) *out = c->cholqr_shifted;
        else if (!strcmp(key, "range_b_reused")) *out = c->range_b_reused;
        else if (!strcmp(key, "workspace_cache_hits")) *out = c->cache_hits;
        else if (!strcmp(key, "workspace_cache_misses")) *out = c->cache_misses;
        else if (!strcmp(key, "workspace_cached_bytes")) *out = (int64_t)c->cached_bytes;
        else RC_THROW(RC_INVALID_ARGUMENT, "unknown counter '%s'", key);
    });
}
rc_status rc_ctx_reset_counters(rc_ctx* c) {
    if (!c) return RC_INVALID_ARGUMENT;
    c->launches = c->gemm_flops = c->h2d_bytes = c->d2h_bytes = 0;
    c->cholqr_used = c->cholqr_fallbacks = c->cholqr_shifted = c->range_b_reused = 0;
    return RC_OK;
}

rc_status rc_comm_get_unique_id(void* out) {
    if (!out) return RC_INVALID_ARGUMENT;
    return guard(nullptr, [&] { comm_get_unique_id(out); });
}
rc_status rc_ctx_comm_init(rc_ctx* c, const void* id, int rank, int nranks) {
    if (!c || !id) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { comm_init(c, id, rank, nranks); });
}
rc_status rc_ctx_comm_info(rc_ctx* c, int* rank, int* nranks) {
    if (!c) return RC_INVALID_ARGUMENT;
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return RC_OK;
}

// ---------------------------------------------------------------- matrices
rc_status rc_matrix_create(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(dt >= 0 && dt <= 3, "bad dtype");
        *out = mat_new(c, dt, rows, cols);
    });
}
rc_status rc_matrix_from_host(rc_ctx* c, rc_dtype dt, const void* host, int64_t rows, int64_t cols,
                              int64_t rs, int64_t cs, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(dt >= 0 && dt <= 3, "bad dtype");
        RC_REQUIRE(host || rows * cols == 0, "null host pointer");
        RC_REQUIRE(rows >= 0 && cols >= 0, "negative dimension");
        MatPtr m(mat_new(c, dt, rows, cols));
        const size_t es = rc_dtype_size(dt);
        if (rows * cols > 0) {
            if (cs == 1 && rs >= cols) {
                RC_CUDA(cudaMemcpy2DAsync(m->data, m->ld * es, host, rs * es, cols * es, rows, cudaMemcpyHostToDevice, c->stream));
            } else {
                // general strided view (e.g. a transposed ndarray view): stage the spanned host
                // range, then gather on device
                RC_REQUIRE(rs >= 0 && cs >= 0, "negative strides are not supported");   // LayoutError analogue
                size_t span = (size_t)((rows - 1) * rs + (cols - 1) * cs + 1);
                DevBuf<char> stage(c, span * es);
                RC_CUDA(cudaMemcpyAsync(stage.p, host, span * es, cudaMemcpyHostToDevice, c->stream));
                RC_DISPATCH(dt, k_strided_to_dense<T>(c, P<T>(m.get()), m->ld, reinterpret_cast<const T*>(stage.p), rs, cs, rows, cols));
            }
            RC_CUDA(cudaStreamSynchronize(c->stream));   // host buffer is not retained
            c->h2d_bytes += rows * cols * (int64_t)es;
        }
        *out = m.release();
    });
}
// Pipelined upload: the copy is queued on the context's copy stream and the call returns at once, so the transfer
// of the NEXT operator runs under the kernels working on the current one (the PCIe link and the SMs are independent
// resources; for configs[1] the 4 GiB upload is 78 ms against 16 ms of compute).  rc_matrix_await orders the context
// stream behind the copy.
rc_status rc_matrix_from_host_async(rc_ctx* c, rc_dtype dt, const void* host, int64_t rows, int64_t cols,
                                    int64_t rs, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(dt >= 0 && dt <= 3, "bad dtype");
        RC_REQUIRE(host || rows * cols == 0, "null host pointer");
        RC_REQUIRE(rows >= 0 && cols >= 0 && rs >= cols, "from_host_async takes a row-major view with row stride >= cols");
        MatPtr m(mat_new(c, dt, rows, cols));
        const size_t es = rc_dtype_size(dt);
        if (rows * cols > 0) {
            if (!c->copy_stream) RC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            // the buffer was allocated (or taken from the workspace cache) in the order of the context stream
            cudaEvent_t ev;
            RC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            m->upload_done = ev;
            RC_CUDA(cudaEventRecord(ev, c->stream));
            RC_CUDA(cudaStreamWaitEvent(c->copy_stream, ev, 0));
            RC_CUDA(cudaMemcpy2DAsync(m->data, m->ld * es, host, rs * es, cols * es, rows, cudaMemcpyHostToDevice, c->copy_stream));
            RC_CUDA(cudaEventRecord(ev, c->copy_stream));
            c->h2d_bytes += rows * cols * (int64_t)es;
        }
        *out = m.release();
    });
}
// Page-lock (and unlock) host memory the caller already owns -- an ndarray allocation, say -- so that uploads from it
// run at the full rate of the link and rc_matrix_from_host_async really is asynchronous (a copy from pageable memory
// is staged through a driver bounce buffer and blocks the host).
rc_status rc_host_register(rc_ctx* c, void* host, size_t bytes) {
    if (!c || !host || bytes == 0) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_CUDA(cudaHostRegister(host, bytes, cudaHostRegisterPortable)); });
}
rc_status rc_host_unregister(rc_ctx* c, void* host) {
    if (!c || !host) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_CUDA(cudaHostUnregister(host)); });
}
rc_status rc_matrix_await(rc_ctx* c, rc_matrix* m, int block_host) {
    if (!c || !m) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (!m->upload_done) return;
        if (block_host) RC_CUDA(cudaEventSynchronize(m->upload_done));
        RC_CUDA(cudaStreamWaitEvent(c->stream, m->upload_done, 0));
        RC_CUDA(cudaEventDestroy(m->upload_done));
        m->upload_done = nullptr;
    });
}
rc_status rc_matrix_wrap_device(rc_ctx* c, rc_dtype dt, void* dptr, int64_t rows, int64_t cols, int64_t ld, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(dt >= 0 && dt <= 3 && dptr && rows >= 0 && cols >= 0 && ld >= cols, "bad arguments");
        rc_matrix* m = new rc_matrix();
        m->ctx = c; m->dtype = dt; m->rows = rows; m->cols = cols; m->ld = ld; m->data = dptr; m->owns = false;
        m->id = rc_next_matrix_id();
        *out = m;
    });
}
/* Matrix-free operator handle: the crate's plugin API (MatVec / ConjMatVec / MatMat / ConjMatMat implemented by
 * the caller, src/types.rs:40-101) for operators that are never materialised. */
rc_status rc_operator_create(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, rc_matmat_fn matmat, rc_matmat_fn conj_matmat,
                             void* user, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(dt >= 0 && dt <= 3 && rows > 0 && cols > 0 && matmat, "bad arguments");
        rc_matrix* m = new rc_matrix();
        m->ctx = c; m->dtype = dt; m->rows = rows; m->cols = cols; m->ld = cols; m->data = nullptr; m->owns = false;
        m->id = rc_next_matrix_id();
        m->op_matmat = matmat; m->op_conj_matmat = conj_matmat; m->op_user = user;
        *out = m;
    });
}
rc_status rc_matrix_to_host(rc_ctx* c, const rc_matrix* m, void* host) {
    if (!c || !m) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (m->rows * m->cols == 0) return;
        RC_REQUIRE(!m->op_matmat, "a matrix-free operator has no entries to copy");
        RC_REQUIRE(host, "null host pointer");
        const size_t es = rc_dtype_size(m->dtype);
        RC_CUDA(cudaMemcpy2DAsync(host, m->cols * es, m->data, m->ld * es, m->cols * es, m->rows, cudaMemcpyDeviceToHost, c->stream));
        RC_CUDA(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += m->rows * m->cols * (int64_t)es;
    });
}
rc_status rc_matrix_to_device(rc_ctx* c, const rc_matrix* m, void* dptr) {
    if (!c || !m) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (m->rows * m->cols == 0) return;
        RC_REQUIRE(!m->op_matmat, "a matrix-free operator has no entries to copy");
        RC_REQUIRE(dptr, "null device pointer");
        const size_t es = rc_dtype_size(m->dtype);
        RC_CUDA(cudaMemcpy2DAsync(dptr, m->cols * es, m->data, m->ld * es, m->cols * es, m->rows, cudaMemcpyDeviceToDevice, c->stream));
    });
}
/* dst <- src (device to device, same shape and scalar type; either side may be a wrapped buffer with its own
 * leading dimension) -- what an operator callback needs to hand a product computed elsewhere back to the library. */
rc_status rc_matrix_copy(rc_ctx* c, const rc_matrix* src, rc_matrix* dst) {
    if (!c || !src || !dst) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(src, dst);
        RC_REQUIRE(src->rows == dst->rows && src->cols == dst->cols, "rc_matrix_copy: shapes differ");
        RC_DISPATCH(src->dtype, k_copy<T>(c, P<T>(dst), dst->ld, P<T>(src), src->ld, src->rows, src->cols));
        // dst's contents changed: anything cached against its identity (the adaptive sampler's B = Q^H A kept for
        // compute_from_range_estimate, see ah_range) must not be matched any more
        dst->id = rc_next_matrix_id();
        if (dst->companion) { mat_free(dst->companion); dst->companion = nullptr; dst->companion_op_id = 0; }
    });
}
rc_status rc_matrix_free(rc_matrix* m) { mat_free(m); return RC_OK; }
int64_t rc_matrix_rows(const rc_matrix* m) { return m ? m->rows : -1; }
int64_t rc_matrix_cols(const rc_matrix* m) { return m ? m->cols : -1; }
int64_t rc_matrix_ld(const rc_matrix* m) { return m ? m->ld : -1; }
int rc_matrix_dtype(const rc_matrix* m) { return m ? m->dtype : -1; }
void* rc_matrix_device_ptr(const rc_matrix* m) { return m ? m->data : nullptr; }
rc_status rc_matrix_set_shard(rc_matrix* m, int64_t global_rows, int64_t row_offset) {
    if (!m) return RC_INVALID_ARGUMENT;
    return guard(m->ctx, [&] {
        RC_REQUIRE(global_rows >= m->rows && row_offset >= 0 && row_offset + m->rows <= global_rows, "bad shard description");
        m->global_rows = global_rows; m->row_offset = row_offset;
    });
}

rc_status rc_matmat(rc_ctx* c, const rc_matrix* a, const rc_matrix* x, rc_matrix** y) {
    if (!c || !a || !x || !y) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(a, x); RC_DISPATCH(a->dtype, *y = matmat_impl<T>(c, a, x)); });
}
rc_status rc_conj_matmat(rc_ctx* c, const rc_matrix* a, const rc_matrix* x, rc_matrix** z) {
    if (!c || !a || !x || !z) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(a, x); RC_DISPATCH(a->dtype, *z = conj_matmat_impl<T>(c, a, x)); });
}

rc_status rc_random_gaussian(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, uint64_t seed, uint32_t stream,
                             int64_t row_offset, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_DISPATCH(dt, *out = gaussian_new<T>(c, dt, rows, cols, seed, stream, row_offset)); });
}
rc_status rc_random_orthogonal_matrix(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, uint64_t seed, uint32_t stream, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(rows > 0 && cols > 0, "empty matrix");
        RC_DISPATCH(dt, *out = random_orthogonal_impl<T>(c, dt, rows, cols, seed, stream));
    });
}
rc_status rc_random_approximate_low_rank_matrix(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, double smax, double smin,
                                                uint64_t seed, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(smin < smax, "`sigma_min` must be smaller than `sigma_max`");     // src/random_matrix.rs:78-82
        RC_REQUIRE(smin > 0.0, "`sigma_min` must be positive.");
        RC_REQUIRE(rows > 0 && cols > 0, "empty matrix");
        const int64_t md = std::min(rows, cols);
        std::vector<double> sig((size_t)md);
        for (int64_t i = 0; i < md; ++i)      // geomspace(sigma_min, sigma_max): ascending (quirk Q4)
            sig[i] = (md == 1) ? smin : smin * std::pow(smax / smin, (double)i / (double)(md - 1));
        RC_DISPATCH(dt, {
            MatPtr u(random_orthogonal_impl<T>(c, dt, rows, md, seed, 101));
            MatPtr vt(random_orthogonal_impl<T>(c, dt, md, cols, seed, 102));
            *out = low_rank_from_factors<T>(c, u.get(), sig, vt.get());
        });
    });
}
rc_status rc_decaying_spectrum_matrix(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, int64_t r0, double decade_every,
                                      uint64_t seed, int64_t row_offset, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(rows > 0 && cols > 0 && r0 > 0 && decade_every > 0, "bad arguments");
        r0 = std::min(r0, std::min(rows, cols));
        std::vector<double> sig((size_t)r0);
        for (int64_t j = 0; j < r0; ++j) sig[j] = std::pow(10.0, -(double)j / decade_every);
        RC_DISPATCH(dt, {
            // orthonormal factors: Q of the TSQR of seeded Gaussians
            auto ortho = [&](int64_t m, uint32_t stream, int64_t off) -> rc_matrix* {
                MatPtr g(gaussian_new<T>(c, dt, m, r0, seed, stream, off));
                QrParts qr;
                pivoted_qr_impl<T>(c, g.get(), false, -1, true, qr);
                return qr.q.release();
            };
            MatPtr u(ortho(rows, 201, row_offset)), v(ortho(cols, 202, 0));
            MatPtr vt(mat_conj_transpose<T>(c, v.get()));
            *out = low_rank_from_factors<T>(c, u.get(), sig, vt.get());
        });
        trim_pool(c);
    });
}

// BASELINE config 4 input: rows [row_offset, row_offset + rows) of
// A = m_total^(-1/2) G diag(10^(-j/decade_every)) V^H, G_ij ~ N(0,1) from Philox keyed by (global row, j)
// (stream 7, like oracle/inputs.py::tall_shard_matrix), V (cols x r0) a shared orthonormal factor.  The left
// factor is NOT orthonormalised, so any rank of any world size regenerates exactly its own rows.
rc_status rc_tall_shard_matrix(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, int64_t r0, double decade_every,
                               uint64_t seed, int64_t row_offset, int64_t m_total, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(rows > 0 && cols > 0 && r0 > 0 && decade_every > 0 && m_total >= row_offset + rows, "bad arguments");
        r0 = std::min(r0, cols);
        std::vector<double> sig((size_t)r0);
        for (int64_t j = 0; j < r0; ++j) sig[j] = std::pow(10.0, -(double)j / decade_every) / std::sqrt((double)m_total);
        RC_DISPATCH(dt, {
            MatPtr g(gaussian_new<T>(c, dt, rows, r0, seed, 7, row_offset));
            MatPtr gv(gaussian_new<T>(c, dt, cols, r0, seed, 202, 0));
            QrParts qr;
            pivoted_qr_impl<T>(c, gv.get(), false, -1, true, qr);
            MatPtr vt(mat_conj_transpose<T>(c, qr.q.get()));
            *out = low_rank_from_factors<T>(c, g.get(), sig, vt.get());
        });
        trim_pool(c);
    });
}

// BASELINE config 5 input: the "low-rank kernel matrix" exp(i kappa |x_i - y_j|) / |x_i - y_j| of
// two well-separated unit boxes, generated on device (real dtypes take the real part).
rc_status rc_helmholtz_kernel_matrix(rc_ctx* c, rc_dtype dt, int64_t rows, int64_t cols, uint64_t seed, double kappa,
                                     double shift, int64_t row_offset, rc_matrix** out) {
    if (!c || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(rows > 0 && cols > 0 && shift > 1.0, "bad arguments (the boxes must not touch: shift > 1)");
        RC_DISPATCH(dt, {
            MatPtr a(mat_new(c, dt, rows, cols));
            k_helmholtz<T>(c, P<T>(a.get()), rows, cols, a->ld, seed, kappa, shift, row_offset);
            *out = a.release();
        });
    });
}

rc_status rc_rel_diff_fro(rc_ctx* c, const rc_matrix* a, const rc_matrix* b, double* out) {
    if (!c || !a || !b || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(a, b); RC_DISPATCH(a->dtype, *out = rel_diff_impl<T>(c, a, b)); });
}
rc_status rc_rel_diff_l2(rc_ctx* c, const rc_matrix* a, const rc_matrix* b, double* out) {
    return rc_rel_diff_fro(c, a, b, out);   // same arithmetic on a 1 x n / n x 1 matrix
}
rc_status rc_max_col_norm(rc_ctx* c, const rc_matrix* m, double* out) {
    if (!c || !m || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_DISPATCH(m->dtype, *out = max_col_norm_impl<T>(c, m)); });
}

// ---------------------------------------------------------------- permutations
rc_status rc_invert_permutation_vector(const uint64_t* perm, size_t n, uint64_t* inverse) {
    if ((!perm || !inverse) && n) return RC_INVALID_ARGUMENT;
    for (size_t i = 0; i < n; ++i) if (perm[i] >= n) return RC_INVALID_ARGUMENT;
    for (size_t i = 0; i < n; ++i) inverse[perm[i]] = i;     // src/permutation.rs:33-35
    return RC_OK;
}
rc_status rc_apply_permutation_matrix(rc_ctx* c, const rc_matrix* m, const uint64_t* idx, size_t n, rc_perm_mode mode, rc_matrix** out) {
    if (!c || !m || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(mode >= 0 && mode <= 3, "bad permutation mode");
        std::vector<uint64_t> v = vec_from(idx, n);
        RC_DISPATCH(m->dtype, *out = permute_impl<T>(c, m, v, mode));
    });
}
rc_status rc_apply_permutation_vector(rc_ctx* c, const rc_matrix* v, const uint64_t* idx, size_t n, rc_vperm_mode mode, rc_matrix** out) {
    if (!c || !v || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(v->rows == 1 || v->cols == 1, "vector permutation expects a 1 x n or n x 1 matrix");
        RC_REQUIRE((int64_t)n == v->rows * v->cols, "The input vector and the index array must have the same length");   // :161-164
        std::vector<uint64_t> iv = vec_from(idx, n);
        int mm = (v->rows == 1) ? (mode == RC_VPERM_INV ? RC_PERM_COLINV : RC_PERM_COL)
                                : (mode == RC_VPERM_INV ? RC_PERM_ROWINV : RC_PERM_ROW);
        RC_DISPATCH(v->dtype, *out = permute_impl<T>(c, v, iv, mm));
    });
}

// ---------------------------------------------------------------- samplers
rc_status rc_sample_range_by_rank(rc_ctx* c, const rc_matrix* a, int64_t k, int64_t p, const rc_matrix* omega, uint64_t seed, rc_matrix** q) {
    if (!c || !a || !q) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (omega) check_same(a, omega);
        RC_DISPATCH(a->dtype, *q = sample_by_rank_impl<T>(c, a, k, p, omega, seed));
    });
}
rc_status rc_sample_range_power_iteration(rc_ctx* c, const rc_matrix* a, int64_t k, int64_t p, int64_t it, const rc_matrix* omega,
                                          uint64_t seed, rc_matrix** q) {
    if (!c || !a || !q) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (omega) check_same(a, omega);
        RC_DISPATCH(a->dtype, *q = sample_power_impl<T>(c, a, k, p, it, omega, seed));
    });
}
rc_status rc_sample_range_adaptive(rc_ctx* c, const rc_matrix* a, double rel_tol, int64_t s, const rc_matrix* omega_blocks, uint64_t seed,
                                   int64_t max_rank, rc_matrix** q, uint64_t* hist_rank, double* hist_res, size_t hist_cap, size_t* hist_len) {
    if (!c || !a || !q) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        if (omega_blocks) check_same(a, omega_blocks);
        std::vector<uint64_t> hr; std::vector<double> hv;
        MatPtr res;
        RC_DISPATCH(a->dtype, res.reset(sample_adaptive_impl<T>(c, a, rel_tol, s, omega_blocks, seed, max_rank, hr, hv)));
        if (hist_len) *hist_len = hr.size();
        size_t ncopy = std::min(hist_cap, hr.size());
        if (hist_rank) for (size_t i = 0; i < ncopy; ++i) hist_rank[i] = hr[i];
        if (hist_res) for (size_t i = 0; i < ncopy; ++i) hist_res[i] = hv[i];
        *q = res.release();
    });
}

// ---------------------------------------------------------------- QR / LQ
static rc_qr* qr_from_parts(QrParts& p) {
    rc_qr* h = new rc_qr();
    h->q = p.q.release(); h->r = p.r.release(); h->ind = std::move(p.ind);
    return h;
}
rc_status rc_qr_compute_from(rc_ctx* c, const rc_matrix* arr, rc_qr** out) {
    if (!c || !arr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        QrParts p;
        RC_DISPATCH(arr->dtype, pivoted_qr_impl<T>(c, arr, false, -1, false, p));
        *out = qr_from_parts(p);
    });
}
// QR { q, r, ind } assembled by the caller (pub fields, src/qr.rs:31-40)
rc_status rc_qr_new(rc_ctx* c, const rc_matrix* q, const rc_matrix* r, const uint64_t* ind, size_t n, rc_qr** out) {
    if (!c || !q || !r || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(q, r);
        RC_REQUIRE(q->cols == r->rows, "QR: q has %lld columns, r has %lld rows", (long long)q->cols, (long long)r->rows);
        RC_REQUIRE((int64_t)n == r->cols, "QR: ind has length %zu, r has %lld columns", n, (long long)r->cols);
        std::unique_ptr<rc_qr> h(new rc_qr());
        h->ind = vec_from(ind, n);
        for (uint64_t v : h->ind) RC_REQUIRE(v < n, "QR: ind entry %llu out of range", (unsigned long long)v);
        RC_DISPATCH(q->dtype, { MatPtr a(mat_clone<T>(c, q)); MatPtr b(mat_clone<T>(c, r)); h->q = a.release(); h->r = b.release(); });
        *out = h.release();
    });
}
rc_status rc_qr_compute_from_range_estimate(rc_ctx* c, const rc_matrix* range, const rc_matrix* op, rc_qr** out) {
    if (!c || !range || !op || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(range, op);
        QrParts p;
        RC_DISPATCH(op->dtype, qr_from_range_impl<T>(c, range, op, p));
        *out = qr_from_parts(p);
    });
}
rc_status rc_qr_compress_rank(rc_ctx* c, const rc_qr* qr, int64_t max_rank, rc_qr** out) {
    if (!c || !qr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(max_rank >= 0, "negative rank");
        max_rank = std::min(max_rank, qr->q->cols);                              // src/qr.rs:172-174
        std::unique_ptr<rc_qr> h(new rc_qr());
        RC_DISPATCH(qr->q->dtype, {
            MatPtr q(mat_slice<T>(c, qr->q, 0, qr->q->rows, 0, max_rank));
            MatPtr r(mat_slice<T>(c, qr->r, 0, max_rank, 0, qr->r->cols));
            h->q = q.release(); h->r = r.release();
        });
        h->ind = qr->ind;                                                        // full length (quirk Q8)
        *out = h.release();
    });
}
static int64_t first_below(rc_ctx* c, const rc_matrix* tri, double tol) {
    // first i with |d_ii / d_00| < tol, ratio formed in the scalar type, compared in f64 (src/qr.rs:190-194)
    int64_t nd = std::min(tri->rows, tri->cols);
    const size_t es = rc_dtype_size(tri->dtype);
    std::vector<char> diag((size_t)nd * es);
    RC_CUDA(cudaMemcpy2DAsync(diag.data(), es, tri->data, (tri->ld + 1) * es, es, nd, cudaMemcpyDeviceToHost, c->stream));
    RC_CUDA(cudaStreamSynchronize(c->stream));
    auto ratio = [&](int64_t i) -> double {
        switch (tri->dtype) {
            case RC_F32: { const float* d = (const float*)diag.data(); return (double)std::fabs(d[i] / d[0]); }
            case RC_F64: { const double* d = (const double*)diag.data(); return std::fabs(d[i] / d[0]); }
            case RC_C32: { const c32* d = (const c32*)diag.data(); c32 q = d[i] / d[0]; return (double)std::hypot(q.re, q.im); }
            default: { const c64* d = (const c64*)diag.data(); c64 q = d[i] / d[0]; return std::hypot(q.re, q.im); }
        }
    };
    for (int64_t i = 0; i < nd; ++i) if (ratio(i) < tol) return i;
    return -1;
}
rc_status rc_qr_compress_tolerance(rc_ctx* c, const rc_qr* qr, double tol, rc_qr** out) {
    if (!c || !qr || !out) return RC_INVALID_ARGUMENT;
    int64_t pos = -1;
    rc_status st = guard(c, [&] {
        RC_REQUIRE((tol < 1.0) && (0.0 <= tol), "Require 0 <= tol < 1.0");      // src/qr.rs:188
        pos = first_below(c, qr->r, tol);
        if (pos < 0) RC_THROW(RC_COMPRESSION_ERROR, "Could not compress to desired tolerance");   // quirk Q3
    });
    if (st != RC_OK) return st;
    return rc_qr_compress_rank(c, qr, pos, out);
}
rc_status rc_qr_to_mat(rc_ctx* c, const rc_qr* qr, rc_matrix** out) {
    if (!c || !qr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_DISPATCH(qr->q->dtype, {
            MatPtr rp(permute_impl<T>(c, qr->r, qr->ind, RC_PERM_COLINV));       // src/qr.rs:160-166
            *out = mat_mul<T>(c, RC_OP_N, qr->q, RC_OP_N, rp.get());
            inherit_shard(*out, qr->q);
        });
    });
}
rc_status rc_qr_column_id(rc_ctx* c, const rc_qr* qr, rc_column_id** out) {
    if (!c || !qr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        std::unique_ptr<rc_column_id> h(new rc_column_id());
        RC_DISPATCH(qr->q->dtype, column_id_impl<T>(c, qr->q, qr->r, qr->ind, h.get()));
        *out = h.release();
    });
}
const rc_matrix* rc_qr_get_q(const rc_qr* qr) { return qr ? qr->q : nullptr; }
const rc_matrix* rc_qr_get_r(const rc_qr* qr) { return qr ? qr->r : nullptr; }
int64_t rc_qr_rank(const rc_qr* qr) { return qr ? qr->q->cols : -1; }
int64_t rc_qr_nrows(const rc_qr* qr) { return qr ? qr->q->rows : -1; }
int64_t rc_qr_ncols(const rc_qr* qr) { return qr ? qr->r->cols : -1; }
rc_status rc_qr_get_ind(const rc_qr* qr, uint64_t* out, size_t n) {
    if (!qr) return RC_INVALID_ARGUMENT;
    return guard(qr->q->ctx, [&] { copy_ind(qr->ind, out, n); });
}
rc_status rc_qr_free(rc_qr* qr) {
    if (qr) { mat_free(qr->q); mat_free(qr->r); delete qr; }
    return RC_OK;
}

rc_status rc_lq_compute_from(rc_ctx* c, const rc_matrix* arr, rc_lq** out) {
    if (!c || !arr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        // pivoted QR of arr^H, transposed back (src/qr.rs:354-362)
        QrParts p;
        std::unique_ptr<rc_lq> h(new rc_lq());
        RC_DISPATCH(arr->dtype, {
            pivoted_qr_impl<T>(c, arr, true, -1, false, p);
            MatPtr l(mat_conj_transpose<T>(c, p.r.get()));
            MatPtr q(mat_conj_transpose<T>(c, p.q.get()));
            h->l = l.release(); h->q = q.release();
        });
        h->ind = std::move(p.ind);
        *out = h.release();
    });
}
// LQ { l, q, ind } assembled by the caller (pub fields, src/qr.rs:42-51)
rc_status rc_lq_new(rc_ctx* c, const rc_matrix* l, const rc_matrix* q, const uint64_t* ind, size_t n, rc_lq** out) {
    if (!c || !l || !q || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(l, q);
        RC_REQUIRE(l->cols == q->rows, "LQ: l has %lld columns, q has %lld rows", (long long)l->cols, (long long)q->rows);
        RC_REQUIRE((int64_t)n == l->rows, "LQ: ind has length %zu, l has %lld rows", n, (long long)l->rows);
        std::unique_ptr<rc_lq> h(new rc_lq());
        h->ind = vec_from(ind, n);
        for (uint64_t v : h->ind) RC_REQUIRE(v < n, "LQ: ind entry %llu out of range", (unsigned long long)v);
        RC_DISPATCH(l->dtype, { MatPtr a(mat_clone<T>(c, l)); MatPtr b(mat_clone<T>(c, q)); h->l = a.release(); h->q = b.release(); });
        *out = h.release();
    });
}
rc_status rc_lq_compress_rank(rc_ctx* c, const rc_lq* lq, int64_t max_rank, rc_lq** out) {
    if (!c || !lq || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(max_rank >= 0, "negative rank");
        max_rank = std::min(max_rank, lq->q->rows);                              // src/qr.rs:84-86
        std::unique_ptr<rc_lq> h(new rc_lq());
        RC_DISPATCH(lq->q->dtype, {
            MatPtr q(mat_slice<T>(c, lq->q, 0, max_rank, 0, lq->q->cols));
            MatPtr l(mat_slice<T>(c, lq->l, 0, lq->l->rows, 0, max_rank));
            h->q = q.release(); h->l = l.release();
        });
        h->ind = lq->ind;
        *out = h.release();
    });
}
rc_status rc_lq_compress_tolerance(rc_ctx* c, const rc_lq* lq, double tol, rc_lq** out) {
    if (!c || !lq || !out) return RC_INVALID_ARGUMENT;
    int64_t pos = -1;
    rc_status st = guard(c, [&] {
        RC_REQUIRE((tol < 1.0) && (0.0 <= tol), "Require 0 <= tol < 1.0");      // src/qr.rs:99
        pos = first_below(c, lq->l, tol);
        if (pos < 0) RC_THROW(RC_COMPRESSION_ERROR, "Could not compress to desired tolerance");
    });
    if (st != RC_OK) return st;
    return rc_lq_compress_rank(c, lq, pos, out);
}
rc_status rc_lq_to_mat(rc_ctx* c, const rc_lq* lq, rc_matrix** out) {
    if (!c || !lq || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_DISPATCH(lq->q->dtype, {
            MatPtr lp(permute_impl<T>(c, lq->l, lq->ind, RC_PERM_ROWINV));       // src/qr.rs:73-78
            *out = mat_mul<T>(c, RC_OP_N, lp.get(), RC_OP_N, lq->q);
        });
    });
}
rc_status rc_lq_row_id(rc_ctx* c, const rc_lq* lq, rc_row_id** out) {
    if (!c || !lq || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        std::unique_ptr<rc_row_id> h(new rc_row_id());
        RC_DISPATCH(lq->q->dtype, row_id_impl<T>(c, lq->l, lq->q, lq->ind, h.get()));
        *out = h.release();
    });
}
const rc_matrix* rc_lq_get_l(const rc_lq* lq) { return lq ? lq->l : nullptr; }
const rc_matrix* rc_lq_get_q(const rc_lq* lq) { return lq ? lq->q : nullptr; }
int64_t rc_lq_rank(const rc_lq* lq) { return lq ? lq->q->rows : -1; }
int64_t rc_lq_nrows(const rc_lq* lq) { return lq ? lq->l->rows : -1; }
int64_t rc_lq_ncols(const rc_lq* lq) { return lq ? lq->q->cols : -1; }
rc_status rc_lq_get_ind(const rc_lq* lq, uint64_t* out, size_t n) {
    if (!lq) return RC_INVALID_ARGUMENT;
    return guard(lq->q->ctx, [&] { copy_ind(lq->ind, out, n); });
}
rc_status rc_lq_free(rc_lq* lq) {
    if (lq) { mat_free(lq->l); mat_free(lq->q); delete lq; }
    return RC_OK;
}

// ---------------------------------------------------------------- SVD
static rc_svd* svd_from_parts(SvdParts& p) {
    rc_svd* h = new rc_svd();
    h->u = p.u.release(); h->vt = p.vt.release(); h->s = std::move(p.s);
    return h;
}
rc_status rc_svd_compute_from(rc_ctx* c, const rc_matrix* arr, rc_svd** out) {
    if (!c || !arr || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        SvdParts p;
        RC_DISPATCH(arr->dtype, svd_impl<T>(c, arr, false, p));
        *out = svd_from_parts(p);
    });
}
// SVD { u, s, vt } assembled by the caller (pub fields, src/svd.rs:13-20)
rc_status rc_svd_new(rc_ctx* c, const rc_matrix* u, const double* s, size_t ns, const rc_matrix* vt, rc_svd** out) {
    if (!c || !u || !vt || !out || (!s && ns)) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(u, vt);
        RC_REQUIRE(u->cols == (int64_t)ns && vt->rows == (int64_t)ns, "SVD: u is %lld x %lld, s has %zu entries, vt is %lld x %lld",
                   (long long)u->rows, (long long)u->cols, ns, (long long)vt->rows, (long long)vt->cols);
        std::unique_ptr<rc_svd> h(new rc_svd());
        h->s.assign(s, s + ns);
        RC_DISPATCH(u->dtype, { MatPtr a(mat_clone<T>(c, u)); MatPtr b(mat_clone<T>(c, vt)); h->u = a.release(); h->vt = b.release(); });
        *out = h.release();
    });
}
rc_status rc_svd_compute_from_range_estimate(rc_ctx* c, const rc_matrix* range, const rc_matrix* op, rc_svd** out) {
    if (!c || !range || !op || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(range, op);
        SvdParts p;
        RC_DISPATCH(op->dtype, svd_from_range_impl<T>(c, range, op, p));
        *out = svd_from_parts(p);
    });
}
rc_status rc_svd_compress_rank(rc_ctx* c, const rc_svd* svd, int64_t max_rank, rc_svd** out) {
    if (!c || !svd || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_REQUIRE(max_rank >= 0, "negative rank");
        max_rank = std::min<int64_t>(max_rank, (int64_t)svd->s.size());          // src/svd.rs:71-73
        std::unique_ptr<rc_svd> h(new rc_svd());
        RC_DISPATCH(svd->u->dtype, {
            MatPtr u(mat_slice<T>(c, svd->u, 0, svd->u->rows, 0, max_rank));
            MatPtr vt(mat_slice<T>(c, svd->vt, 0, max_rank, 0, svd->vt->cols));
            h->u = u.release(); h->vt = vt.release();
        });
        h->s.assign(svd->s.begin(), svd->s.begin() + max_rank);
        *out = h.release();
    });
}
rc_status rc_svd_compress_tolerance(rc_ctx* c, const rc_svd* svd, double tol, rc_svd** out) {
    if (!c || !svd || !out) return RC_INVALID_ARGUMENT;
    int64_t pos = -1;
    rc_status st = guard(c, [&] {
        RC_REQUIRE((tol < 1.0) && (0.0 <= tol), "Require 0 <= tol < 1.0");      // src/svd.rs:88
        RC_REQUIRE(!svd->s.empty(), "empty SVD");
        const bool single = (svd->u->dtype == RC_F32 || svd->u->dtype == RC_C32);
        for (size_t i = 0; i < svd->s.size(); ++i) {                             // src/svd.rs:92-95
            double ratio = single ? (double)((float)svd->s[i] / (float)svd->s[0]) : svd->s[i] / svd->s[0];
            if (ratio < tol) { pos = (int64_t)i; break; }
        }
        if (pos < 0) RC_THROW(RC_COMPRESSION_ERROR, "Could not compress to desired tolerance");
    });
    if (st != RC_OK) return st;
    return rc_svd_compress_rank(c, svd, pos, out);
}
rc_status rc_svd_to_mat(rc_ctx* c, const rc_svd* svd, rc_matrix** out) {
    if (!c || !svd || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_DISPATCH(svd->u->dtype, {
            MatPtr sv(scaled_vt<T>(c, svd));                                     // src/svd.rs:42-54
            *out = mat_mul<T>(c, RC_OP_N, svd->u, RC_OP_N, sv.get());
            inherit_shard(*out, svd->u);
        });
    });
}
rc_status rc_svd_to_qr(rc_ctx* c, const rc_svd* svd, rc_qr** out) {
    if (!c || !svd || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        QrParts p;
        RC_DISPATCH(svd->u->dtype, {
            MatPtr sv(scaled_vt<T>(c, svd));                                     // src/svd.rs:150-163
            pivoted_qr_impl<T>(c, sv.get(), false, -1, true, p);
            MatPtr q(mat_mul<T>(c, RC_OP_N, svd->u, RC_OP_N, p.q.get()));
            inherit_shard(q.get(), svd->u);
            p.q.reset(q.release());
        });
        *out = qr_from_parts(p);
    });
}
const rc_matrix* rc_svd_get_u(const rc_svd* svd) { return svd ? svd->u : nullptr; }
const rc_matrix* rc_svd_get_vt(const rc_svd* svd) { return svd ? svd->vt : nullptr; }
int64_t rc_svd_rank(const rc_svd* svd) { return svd ? (int64_t)svd->s.size() : -1; }
rc_status rc_svd_get_s(const rc_svd* svd, double* out, size_t n) {
    if (!svd || (!out && n)) return RC_INVALID_ARGUMENT;
    if (n < svd->s.size()) return RC_INVALID_ARGUMENT;
    memcpy(out, svd->s.data(), svd->s.size() * sizeof(double));
    return RC_OK;
}
rc_status rc_svd_free(rc_svd* svd) {
    if (svd) { mat_free(svd->u); mat_free(svd->vt); delete svd; }
    return RC_OK;
}

// ---------------------------------------------------------------- interpolative decompositions
rc_status rc_column_id_new(rc_ctx* c, const rc_matrix* cm, const rc_matrix* z, const uint64_t* col_ind, size_t n, rc_column_id** out) {
    if (!c || !cm || !z || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(cm, z);
        std::unique_ptr<rc_column_id> h(new rc_column_id());
        h->col_ind = vec_from(col_ind, n);
        RC_DISPATCH(cm->dtype, { MatPtr a(mat_clone<T>(c, cm)); MatPtr b(mat_clone<T>(c, z)); h->c = a.release(); h->z = b.release(); });
        *out = h.release();
    });
}
const rc_matrix* rc_column_id_get_c(const rc_column_id* id) { return id ? id->c : nullptr; }
const rc_matrix* rc_column_id_get_z(const rc_column_id* id) { return id ? id->z : nullptr; }
rc_status rc_column_id_get_col_ind(const rc_column_id* id, uint64_t* out, size_t n) {
    if (!id) return RC_INVALID_ARGUMENT;
    return guard(id->c->ctx, [&] { copy_ind(id->col_ind, out, n); });
}
rc_status rc_column_id_to_mat(rc_ctx* c, const rc_column_id* id, rc_matrix** out) {
    if (!c || !id || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_DISPATCH(id->c->dtype, *out = mat_mul<T>(c, RC_OP_N, id->c, RC_OP_N, id->z)); });   // :64
}
rc_status rc_column_id_apply(rc_ctx* c, const rc_column_id* id, const rc_matrix* rhs, rc_matrix** out) {
    if (!c || !id || !rhs || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(id->c, rhs); RC_DISPATCH(id->c->dtype, *out = chain_apply<T>(c, {id->c, id->z}, rhs)); });
}
rc_status rc_column_id_two_sided_id(rc_ctx* c, const rc_column_id* id, rc_two_sided_id** out) {
    if (!c || !id || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        // LQ::compute_from(C).row_id() (src/col_interp_decomp.rs:116-125; quirk Q10: uncompressed)
        std::unique_ptr<rc_two_sided_id> h(new rc_two_sided_id());
        RC_DISPATCH(id->c->dtype, {
            // Row-sharded C: the m columns of C^H are spread over the ranks.  C is small
            // (m x k), so it is all-gathered and the pivoted LQ / row ID run redundantly on every rank --
            // identical pivots to the single-GPU run by construction (a tournament pivoting would not be);
            // every rank keeps its own rows of the m x k factor X.
            const rc_matrix* cm = id->c;
            MatPtr cfull;
            const bool sharded = mat_sharded(id->c);
            if (sharded) {
                RC_REQUIRE(id->c->rows * c->nranks == id->c->global_rows, "two_sided_id: row shards must have equal sizes");
                cfull.reset(mat_new(c, id->c->dtype, id->c->global_rows, id->c->cols));
                RC_REQUIRE(cfull->ld == id->c->ld, "two_sided_id: unexpected leading dimension");
                comm_allgather(c, id->c->data, cfull->data, (size_t)id->c->rows * id->c->ld * sizeof(T));
                cm = cfull.get();
            }
            QrParts p;
            pivoted_qr_impl<T>(c, cm, true, -1, false, p);
            MatPtr l(mat_conj_transpose<T>(c, p.r.get()));
            MatPtr q(mat_conj_transpose<T>(c, p.q.get()));
            rc_row_id rid;
            row_id_impl<T>(c, l.get(), q.get(), p.ind, &rid);
            if (sharded) {
                MatPtr xl(mat_slice<T>(c, rid.x, id->c->row_offset, id->c->row_offset + id->c->rows, 0, rid.x->cols));
                inherit_shard(xl.get(), id->c);
                mat_free(rid.x);
                rid.x = xl.release();
            }
            h->c = rid.x; h->x = rid.r; h->row_ind = rid.row_ind;
            MatPtr zc(mat_clone<T>(c, id->z));
            h->r = zc.release();
        });
        h->col_ind = id->col_ind;
        *out = h.release();
    });
}
/* lengths of the (always full-length, quirk Q8) index vectors: with row-sharded factors the local row count of
 * c / x is not the length of row_ind */
size_t rc_column_id_col_ind_len(const rc_column_id* id) { return id ? id->col_ind.size() : 0; }
size_t rc_row_id_row_ind_len(const rc_row_id* id) { return id ? id->row_ind.size() : 0; }
size_t rc_two_sided_id_row_ind_len(const rc_two_sided_id* id) { return id ? id->row_ind.size() : 0; }
size_t rc_two_sided_id_col_ind_len(const rc_two_sided_id* id) { return id ? id->col_ind.size() : 0; }
rc_status rc_column_id_free(rc_column_id* id) {
    if (id) { mat_free(id->c); mat_free(id->z); delete id; }
    return RC_OK;
}

rc_status rc_row_id_new(rc_ctx* c, const rc_matrix* x, const rc_matrix* r, const uint64_t* row_ind, size_t n, rc_row_id** out) {
    if (!c || !x || !r || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(x, r);
        std::unique_ptr<rc_row_id> h(new rc_row_id());
        h->row_ind = vec_from(row_ind, n);
        RC_DISPATCH(x->dtype, { MatPtr a(mat_clone<T>(c, x)); MatPtr b(mat_clone<T>(c, r)); h->x = a.release(); h->r = b.release(); });
        *out = h.release();
    });
}
const rc_matrix* rc_row_id_get_x(const rc_row_id* id) { return id ? id->x : nullptr; }
const rc_matrix* rc_row_id_get_r(const rc_row_id* id) { return id ? id->r : nullptr; }
rc_status rc_row_id_get_row_ind(const rc_row_id* id, uint64_t* out, size_t n) {
    if (!id) return RC_INVALID_ARGUMENT;
    return guard(id->x->ctx, [&] { copy_ind(id->row_ind, out, n); });
}
rc_status rc_row_id_to_mat(rc_ctx* c, const rc_row_id* id, rc_matrix** out) {
    if (!c || !id || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { RC_DISPATCH(id->x->dtype, *out = mat_mul<T>(c, RC_OP_N, id->x, RC_OP_N, id->r)); });   // :66
}
rc_status rc_row_id_apply(rc_ctx* c, const rc_row_id* id, const rc_matrix* rhs, rc_matrix** out) {
    if (!c || !id || !rhs || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(id->x, rhs); RC_DISPATCH(id->x->dtype, *out = chain_apply<T>(c, {id->x, id->r}, rhs)); });
}
rc_status rc_row_id_two_sided_id(rc_ctx* c, const rc_row_id* id, rc_two_sided_id** out) {
    if (!c || !id || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        // QR::compute_from(R).column_id() (src/row_interp_decomp.rs:120-130)
        std::unique_ptr<rc_two_sided_id> h(new rc_two_sided_id());
        RC_DISPATCH(id->x->dtype, {
            QrParts p;
            pivoted_qr_impl<T>(c, id->r, false, -1, false, p);
            rc_column_id cid;
            column_id_impl<T>(c, p.q.get(), p.r.get(), p.ind, &cid);
            h->x = cid.c; h->r = cid.z; h->col_ind = cid.col_ind;
            MatPtr xc(mat_clone<T>(c, id->x));
            h->c = xc.release();
        });
        h->row_ind = id->row_ind;
        *out = h.release();
    });
}
rc_status rc_row_id_free(rc_row_id* id) {
    if (id) { mat_free(id->x); mat_free(id->r); delete id; }
    return RC_OK;
}

rc_status rc_two_sided_id_new(rc_ctx* c, const rc_matrix* x, const rc_matrix* r, const rc_matrix* cm, const uint64_t* col_ind, size_t n_col,
                              const uint64_t* row_ind, size_t n_row, rc_two_sided_id** out) {
    if (!c || !x || !r || !cm || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        check_same(x, r); check_same(x, cm);
        std::unique_ptr<rc_two_sided_id> h(new rc_two_sided_id());
        h->col_ind = vec_from(col_ind, n_col);
        h->row_ind = vec_from(row_ind, n_row);
        RC_DISPATCH(x->dtype, {
            MatPtr a(mat_clone<T>(c, cm)); MatPtr b(mat_clone<T>(c, x)); MatPtr d(mat_clone<T>(c, r));
            h->c = a.release(); h->x = b.release(); h->r = d.release();
        });
        *out = h.release();
    });
}
const rc_matrix* rc_two_sided_id_get_c(const rc_two_sided_id* id) { return id ? id->c : nullptr; }
const rc_matrix* rc_two_sided_id_get_x(const rc_two_sided_id* id) { return id ? id->x : nullptr; }
const rc_matrix* rc_two_sided_id_get_r(const rc_two_sided_id* id) { return id ? id->r : nullptr; }
rc_status rc_two_sided_id_get_row_ind(const rc_two_sided_id* id, uint64_t* out, size_t n) {
    if (!id) return RC_INVALID_ARGUMENT;
    return guard(id->c->ctx, [&] { copy_ind(id->row_ind, out, n); });
}
rc_status rc_two_sided_id_get_col_ind(const rc_two_sided_id* id, uint64_t* out, size_t n) {
    if (!id) return RC_INVALID_ARGUMENT;
    return guard(id->c->ctx, [&] { copy_ind(id->col_ind, out, n); });
}
rc_status rc_two_sided_id_to_mat(rc_ctx* c, const rc_two_sided_id* id, rc_matrix** out) {
    if (!c || !id || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] {
        RC_DISPATCH(id->c->dtype, {
            MatPtr xr(mat_mul<T>(c, RC_OP_N, id->x, RC_OP_N, id->r));            // c . (x . r)  (:62-64)
            *out = mat_mul<T>(c, RC_OP_N, id->c, RC_OP_N, xr.get());
        });
    });
}
rc_status rc_two_sided_id_apply(rc_ctx* c, const rc_two_sided_id* id, const rc_matrix* rhs, rc_matrix** out) {
    if (!c || !id || !rhs || !out) return RC_INVALID_ARGUMENT;
    return guard(c, [&] { check_same(id->c, rhs); RC_DISPATCH(id->c->dtype, *out = chain_apply<T>(c, {id->c, id->x, id->r}, rhs)); });
}
rc_status rc_two_sided_id_free(rc_two_sided_id* id) {
    if (id) { mat_free(id->c); mat_free(id->x); mat_free(id->r); delete id; }
    return RC_OK;
}

}  // extern "C"
