// One-sided (Hestenes) Jacobi SVD of a small square matrix: the B200 replacement for LAPACK
// ?gesdd jobz='S' (reference N3: src/compute_svd.rs:19) after the k x n factor has been reduced
// to k x k by a TSQR of its conjugate transpose (host_linalg.cu: svd_impl).
//
// G (n x n, column-major) is rotated from the right until its columns are mutually orthogonal:
// G J = U diag(s), J accumulated in V.  One warp per column pair, round-robin ordering, all pairs
// of a round in parallel; Gram entries accumulated in double.  Singular values come out with
// high relative accuracy, sorted descending like ?gesdd's.
#include <cooperative_groups.h>
#include "rc_internal.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int JT = 512;            // at most 16 warps (128 registers per thread: the register tiles of a pair)

// LPP lanes share a column pair, so a warp rotates 32 / LPP pairs at once.  A round is ISSUE-bound, not latency-bound:
// with one pair per warp every one of the 32 lanes executed the ~150-instruction scalar chain of its pair redundantly
// and the 32 warps of a round cost ~2 400 issue cycles on the four schedulers (measured 3 300 cycles per round at
// n = 64); with 8 lanes per pair the same round is 8 warps of about the same length.
template <class T, bool SMEM, int LPP>
__global__ void __launch_bounds__(JT)
jacobi_kernel(T* __restrict__ Gg, T* __restrict__ Vg, int rows, int n, int ldg, int max_sweeps, double tol,
              T* __restrict__ u, int64_t ldu, double* __restrict__ s_out, T* __restrict__ w, int64_t ldw,
              double* __restrict__ sig_scratch, int* __restrict__ info) {
    __shared__ int s_rot;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, JW = NT >> 5;
    // SMEM: G (rows x n, column stride ldg) and V (n x n) are resident in shared memory for the whole iteration
    T* G = SMEM ? reinterpret_cast<T*>(smem_raw) : Gg;
    T* V = SMEM ? G + (size_t)ldg * n : Vg;
    const int ldgg = SMEM ? ldg : rows;             // the global copy is dense
    if (SMEM) {
        for (int e = tid; e < rows * n; e += NT) { int c = e / rows, r = e - c * rows; G[(size_t)c * ldg + r] = Gg[e]; }
    }
    const int npad = n + (n & 1);
    const int half = npad / 2;
    // V = I
    for (int e = tid; e < n * n; e += NT) {
        int c = e / n, r = e - c * n;
        V[e] = (r == c) ? rc_one<T>() : rc_zero<T>();
    }
    __syncthreads();
    constexpr int PPW = 32 / LPP;                   // pairs per warp
    const int sl = lane % LPP, sg = lane / LPP;
    const double tol2 = tol * tol;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < npad - 1; ++round) {
            for (int k0 = warp * PPW; k0 < half; k0 += JW * PPW) {          // warp-uniform trip count (shuffles below)
                const int k = k0 + sg;
                int a = npad - 1, b = round;
                if (k > 0) {
                    a = round + k; if (a >= npad - 1) a -= npad - 1;
                    b = round - k; if (b < 0) b += npad - 1;
                }
                const int p = min(a, b), q = max(a, b);
                const bool valid = (k < half) && (q < n);                   // q == n: padding column of an odd n
                T* gp = G + (int64_t)(valid ? p : 0) * ldgg;
                T* gq = G + (int64_t)(valid ? q : 0) * ldgg;
                T* vp = V + (int64_t)(valid ? p : 0) * n;
                T* vq = V + (int64_t)(valid ? q : 0) * n;
                double alpha = 0.0, beta = 0.0, gre = 0.0, gim = 0.0;
                // Register tile: each lane's share of the two columns (of G and of V) is loaded ONCE, up front, and
                // stored once at the end.  A load-rotate-store loop over shared memory serialises on its own stores
                // (the next iteration's loads may alias them), which tripled the length of a round.
                constexpr int RPL = ScalarTraits<T>::is_complex ? 4 : 8;
                const bool tiled = (rows <= RPL * LPP) && (n <= RPL * LPP);
                T xg[RPL], yg[RPL], xv[RPL], yv[RPL];
                if (tiled) {
#pragma unroll
                    for (int u = 0; u < RPL; ++u) {
                        const int r = sl + u * LPP;
                        const bool in = valid && r < rows, inv = valid && r < n;
                        xg[u] = in ? gp[r] : rc_zero<T>();
                        yg[u] = in ? gq[r] : rc_zero<T>();
                        xv[u] = inv ? vp[r] : rc_zero<T>();
                        yv[u] = inv ? vq[r] : rc_zero<T>();
                    }
#pragma unroll
                    for (int u = 0; u < RPL; ++u) {
                        alpha += rc_abs2(xg[u]);
                        beta += rc_abs2(yg[u]);
                        const double xr = (double)rc_real(xg[u]), xi = (double)rc_imag(xg[u]);
                        const double yr = (double)rc_real(yg[u]), yi = (double)rc_imag(yg[u]);
                        gre += xr * yr + xi * yi;
                        gim += xr * yi - xi * yr;
                    }
                } else if (valid) {
                    for (int r = sl; r < rows; r += LPP) {
                        T x = gp[r], y = gq[r];
                        alpha += rc_abs2(x);
                        beta += rc_abs2(y);
                        // gamma = conj(x) * y
                        double xr = (double)rc_real(x), xi = (double)rc_imag(x);
                        double yr = (double)rc_real(y), yi = (double)rc_imag(y);
                        gre += xr * yr + xi * yi;
                        gim += xr * yi - xi * yr;
                    }
                }
#pragma unroll
                for (int m = LPP / 2; m > 0; m >>= 1) {
                    alpha += __shfl_xor_sync(0xffffffffu, alpha, m);
                    beta += __shfl_xor_sync(0xffffffffu, beta, m);
                    gre += __shfl_xor_sync(0xffffffffu, gre, m);
                    if (ScalarTraits<T>::is_complex) gim += __shfl_xor_sync(0xffffffffu, gim, m);
                }
                // The scalar chain at the head of every rotation is kept to reciprocal square roots and one
                // reciprocal: no hypot, no divisions, no square root of alpha * beta (the threshold is compared in
                // squares).  The angle only steers convergence; the rotation is orthogonal because cs and sn are
                // derived from the same t.
                const double g2 = gre * gre + gim * gim;
                if (valid && g2 != 0.0 && g2 > tol2 * alpha * beta) {
                    if (sl == 0) s_rot = 1;
                    const double rg = rsqrt(g2);                        // 1 / |gamma|
                    const double zeta = (beta - alpha) * 0.5 * rg;
                    const double w1 = 1.0 + zeta * zeta;
                    const double t = copysign(1.0, zeta) * __drcp_rn(fabs(zeta) + w1 * rsqrt(w1));
                    const double cs = rsqrt(1.0 + t * t), sn = cs * t;
                    // e^{-i theta} = conj(gamma)/|gamma|
                    const double er = gre * rg, ei = -gim * rg;
                    T ph = rc_make<T>(er, ei);
                    T csT = rc_make<T>(cs, 0.0), snT = rc_make<T>(sn, 0.0);
                    if (tiled) {
#pragma unroll
                        for (int u = 0; u < RPL; ++u) {
                            const int r = sl + u * LPP;
                            if (r < rows) {
                                const T y = ph * yg[u];
                                gp[r] = csT * xg[u] - snT * y;
                                gq[r] = snT * xg[u] + csT * y;
                            }
                            if (r < n) {
                                const T y = ph * yv[u];
                                vp[r] = csT * xv[u] - snT * y;
                                vq[r] = snT * xv[u] + csT * y;
                            }
                        }
                    } else {
                        for (int r = sl; r < rows; r += LPP) {
                            T x = gp[r], y = ph * gq[r];
                            gp[r] = csT * x - snT * y;
                            gq[r] = snT * x + csT * y;
                        }
                        for (int r = sl; r < n; r += LPP) {
                            T x = vp[r], y = ph * vq[r];
                            vp[r] = csT * x - snT * y;
                            vq[r] = snT * x + csT * y;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (s_rot == 0) break;
        __syncthreads();
    }
    if (tid == 0) { info[0] = (sweep >= max_sweeps) ? 1 : 0; info[1] = sweep; }
    // singular values
    for (int c = warp; c < n; c += JW) {
        const T* gc = G + (int64_t)c * ldgg;
        double a = 0.0;
        for (int r = lane; r < rows; r += 32) a += rc_abs2(gc[r]);
        a = rc_warp_sum(a);
        if (lane == 0) sig_scratch[c] = sqrt(a);
    }
    __syncthreads();
    // rank sort (descending, stable) and scatter
    for (int c = warp; c < n; c += JW) {
        double sc = sig_scratch[c];
        int rank = 0;
        for (int o = lane; o < n; o += 32) {
            double so = sig_scratch[o];
            if (so > sc || (so == sc && o < c)) ++rank;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, m);
        const T* gc = G + (int64_t)c * ldgg;
        const T* vc = V + (int64_t)c * n;
        double inv = (sc > 0.0) ? 1.0 / sc : 0.0;
        T invT = rc_make<T>(inv, 0.0);
        for (int r = lane; r < rows; r += 32) u[(int64_t)r * ldu + rank] = gc[r] * invT;
        for (int r = lane; r < n; r += 32) w[(int64_t)r * ldw + rank] = vc[r];
        if (lane == 0) s_out[rank] = sc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Large factors (the w x w triangle of a dense matrix that does not fit one CTA's shared memory: SVD::compute_from on
// a general matrix, src/svd.rs:165-169): the same one-sided Jacobi iteration as a persistent COOPERATIVE kernel over
// the whole GPU.  G and V stay in global memory (L2-resident up to n ~ 2 800 in f64), one warp per column pair, all
// pairs of a round-robin round in parallel across the grid, one grid-wide sync per round.  Columns move between SMs
// from round to round, so every load bypasses L1 (ld.global.cg).
__device__ __forceinline__ float  jld(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double jld(const double* p) { return __ldcg(p); }
__device__ __forceinline__ c32 jld(const c32* p) { float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return c32(v.x, v.y); }
__device__ __forceinline__ c64 jld(const c64* p) { double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return c64(v.x, v.y); }

constexpr int JG_NT = 256;
constexpr int JG_RPT = 8;         // rows of a column pair a thread keeps in registers between the Gram pass and the rotation

// One CTA per column pair: 256 threads share the two columns (a warp per pair left the round latency-bound: 1 024 warps
// with one L2 round trip per 32 rows each -- 190 us per round at n = 2 048), the pair's Gram entries are reduced
// through shared memory, and a thread's rows stay in registers between the Gram pass and the rotation.
template <class T>
__global__ void __launch_bounds__(JG_NT)
jacobi_grid_kernel(T* __restrict__ G, T* __restrict__ V, int rows, int n, int max_sweeps, double tol, int* __restrict__ flags,
                   T* __restrict__ u, int64_t ldu, double* __restrict__ s_out, T* __restrict__ w, int64_t ldw,
                   double* __restrict__ sig, int* __restrict__ info) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_part[JG_NT / 32][4];
    __shared__ double s_tot[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (int64_t)gridDim.x * blockDim.x;
    const int gw = (int)(gtid >> 5), nwarps = (int)(gthreads >> 5);
    const int npad = n + (n & 1), half = npad / 2;
    for (int64_t e = gtid; e < (int64_t)n * n; e += gthreads) V[e] = (e / n == e % n) ? rc_one<T>() : rc_zero<T>();
    grid.sync();
    const double tol2 = tol * tol;
    const bool cached = rows <= JG_NT * JG_RPT;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        bool rotated = false;
        for (int round = 0; round < npad - 1; ++round) {
            for (int k = blockIdx.x; k < half; k += gridDim.x) {
                int a = npad - 1, b = round;
                if (k > 0) {
                    a = round + k; if (a >= npad - 1) a -= npad - 1;
                    b = round - k; if (b < 0) b += npad - 1;
                }
                const int p = min(a, b), q = max(a, b);
                if (q >= n) continue;                                   // padding column of an odd n (block-uniform)
                T* gp = G + (int64_t)p * rows;
                T* gq = G + (int64_t)q * rows;
                double alpha = 0.0, beta = 0.0, gre = 0.0, gim = 0.0;
                T xr_[JG_RPT], yr_[JG_RPT];
                if (cached) {
#pragma unroll
                    for (int t = 0; t < JG_RPT; ++t) {
                        const int r = tid + t * JG_NT;
                        xr_[t] = (r < rows) ? jld(gp + r) : rc_zero<T>();
                        yr_[t] = (r < rows) ? jld(gq + r) : rc_zero<T>();
                    }
#pragma unroll
                    for (int t = 0; t < JG_RPT; ++t) {
                        alpha += rc_abs2(xr_[t]);
                        beta += rc_abs2(yr_[t]);
                        const double xr = (double)rc_real(xr_[t]), xi = (double)rc_imag(xr_[t]);
                        const double yr = (double)rc_real(yr_[t]), yi = (double)rc_imag(yr_[t]);
                        gre += xr * yr + xi * yi;
                        gim += xr * yi - xi * yr;
                    }
                } else {
                    for (int r = tid; r < rows; r += JG_NT) {
                        const T x = jld(gp + r), y = jld(gq + r);
                        alpha += rc_abs2(x);
                        beta += rc_abs2(y);
                        const double xr = (double)rc_real(x), xi = (double)rc_imag(x);
                        const double yr = (double)rc_real(y), yi = (double)rc_imag(y);
                        gre += xr * yr + xi * yi;
                        gim += xr * yi - xi * yr;
                    }
                }
                alpha = rc_warp_sum(alpha); beta = rc_warp_sum(beta); gre = rc_warp_sum(gre); gim = rc_warp_sum(gim);
                if (lane == 0) { s_part[warp][0] = alpha; s_part[warp][1] = beta; s_part[warp][2] = gre; s_part[warp][3] = gim; }
                __syncthreads();
                if (tid < 4) {
                    double v = 0.0;
#pragma unroll
                    for (int ww = 0; ww < JG_NT / 32; ++ww) v += s_part[ww][tid];
                    s_tot[tid] = v;
                }
                __syncthreads();
                alpha = s_tot[0]; beta = s_tot[1]; gre = s_tot[2]; gim = s_tot[3];
                const double g2 = gre * gre + gim * gim;
                if (g2 != 0.0 && g2 > tol2 * alpha * beta) {
                    rotated = true;
                    const double rg = rsqrt(g2);
                    const double zeta = (beta - alpha) * 0.5 * rg;
                    const double w1 = 1.0 + zeta * zeta;
                    const double t = copysign(1.0, zeta) * __drcp_rn(fabs(zeta) + w1 * rsqrt(w1));
                    const double cs = rsqrt(1.0 + t * t), sn = cs * t;
                    const T ph = rc_make<T>(gre * rg, -gim * rg);
                    const T csT = rc_make<T>(cs, 0.0), snT = rc_make<T>(sn, 0.0);
                    if (cached) {
#pragma unroll
                        for (int t2 = 0; t2 < JG_RPT; ++t2) {
                            const int r = tid + t2 * JG_NT;
                            if (r < rows) {
                                const T y = ph * yr_[t2];
                                gp[r] = csT * xr_[t2] - snT * y;
                                gq[r] = snT * xr_[t2] + csT * y;
                            }
                        }
                    } else {
                        for (int r = tid; r < rows; r += JG_NT) {
                            const T x = jld(gp + r), y = ph * jld(gq + r);
                            gp[r] = csT * x - snT * y;
                            gq[r] = snT * x + csT * y;
                        }
                    }
                    T* vp = V + (int64_t)p * n;
                    T* vq = V + (int64_t)q * n;
                    for (int r = tid; r < n; r += JG_NT) {
                        const T x = jld(vp + r), y = ph * jld(vq + r);
                        vp[r] = csT * x - snT * y;
                        vq[r] = snT * x + csT * y;
                    }
                }
                // (s_part / s_tot are rewritten by the next pair only after the two barriers above)
            }
            grid.sync();
        }
        if (rotated && tid == 0) flags[sweep] = 1;
        grid.sync();
        if (__ldcg(flags + sweep) == 0) break;
    }
    if (gtid == 0) { info[0] = (sweep >= max_sweeps) ? 1 : 0; info[1] = sweep; }
    for (int c = gw; c < n; c += nwarps) {
        const T* gc = G + (int64_t)c * rows;
        double a = 0.0;
        for (int r = lane; r < rows; r += 32) a += rc_abs2(jld(gc + r));
        a = rc_warp_sum(a);
        if (lane == 0) sig[c] = sqrt(a);
    }
    grid.sync();
    for (int c = gw; c < n; c += nwarps) {
        const double sc = __ldcg(sig + c);
        int rank = 0;
        for (int o = lane; o < n; o += 32) {
            const double so = __ldcg(sig + o);
            if (so > sc || (so == sc && o < c)) ++rank;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, m);
        const T* gc = G + (int64_t)c * rows;
        const T* vc = V + (int64_t)c * n;
        const T invT = rc_make<T>((sc > 0.0) ? 1.0 / sc : 0.0, 0.0);
        for (int r = lane; r < rows; r += 32) u[(int64_t)r * ldu + rank] = jld(gc + r) * invT;
        for (int r = lane; r < n; r += 32) w[(int64_t)r * ldw + rank] = jld(vc + r);
        if (lane == 0) s_out[rank] = sc;
    }
}

}  // namespace

template <class T>
void jacobi_svd(rc_ctx* c, const T* g, int64_t ldg, int64_t rows, int64_t n, T* u, int64_t ldu, double* s, T* w, int64_t ldw, int* info_dev) {
    RC_REQUIRE(n > 0 && n <= 16384 && rows >= n && rows < ((int64_t)1 << 31), "jacobi_svd: unsupported size %lld x %lld", (long long)rows, (long long)n);
    DevBuf<T> G(c, (size_t)rows * n), V(c, (size_t)n * n);
    DevBuf<double> sig(c, (size_t)n);
    DevBuf<int> info_own;
    if (!info_dev) { info_own.alloc(c, 2); info_dev = info_own.p; }
    struct { int* p; } info{info_dev};
    // column-major copy of g == transpose of the row-major matrix
    k_transpose<T>(c, G.p, rows, g, ldg, rows, n, false);
    double eps = (sizeof(RealOf<T>) == 4) ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
    // rotation threshold on the cosine of a column pair (?gesvj: sqrt(m) eps).  For very short columns the floor of 8 eps
    // keeps the threshold above the rounding error of the computed Gram entries themselves (a 2 x 2 factor could rotate
    // for ever on an eps-sized cosine and report non-convergence).
    double tol = eps * sqrt((double)std::max<int64_t>(rows, 64));
    // column stride of the shared-memory copy: an odd multiple of 64 bytes when the rows would otherwise put every
    // column on the same banks (the lane groups of a warp work on different columns)
    const int ldc = (int)((rows * sizeof(T)) % 128 == 0 ? rows + (int64_t)(64 / sizeof(T)) : rows);
    size_t smem = ((size_t)ldc * n + (size_t)n * n) * sizeof(T);
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    const bool fits = smem + 4096 <= lim;
    // lanes per column pair: the fewest that still hold a pair's share of G and V in the register tile of the kernel
    // (8 rows per lane, 4 for complex scalars); a warp rotates 32 / lpp pairs at once
    const int rpl = ScalarTraits<T>::is_complex ? 4 : 8;
    const int lpp = (rows <= 8 * rpl) ? 8 : ((rows <= 16 * rpl) ? 16 : 32);
    const int half = (int)((n + 1) / 2);
    const int nthreads = std::min(JT, std::max(64, 32 * ((half + 32 / lpp - 1) / (32 / lpp))));
#define RC_JACOBI(SM, L)                                                                                              \
    do {                                                                                                              \
        if (SM) RC_CUDA(cudaFuncSetAttribute(jacobi_kernel<T, SM, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        jacobi_kernel<T, SM, L><<<1, nthreads, SM ? smem : 0, c->stream>>>(G.p, V.p, (int)rows, (int)n, ldc, 60, tol, u, ldu, s, w, ldw, sig.p, info.p); \
    } while (0)
    if (fits) {
        if (lpp == 8) RC_JACOBI(true, 8); else if (lpp == 16) RC_JACOBI(true, 16); else RC_JACOBI(true, 32);
        RC_CHECK_LAUNCH(c);
    } else {
        // does not fit one CTA: the cooperative whole-GPU kernel (one warp per column pair, one grid sync per round)
        int per_sm = 0;
        RC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_grid_kernel<T>, JG_NT, 0));
        const int want = half;                       // one CTA per column pair of a round
        int grid = std::max(1, std::min(want, per_sm * c->sm_count));
        const int max_sweeps = 60;
        DevBuf<int> flags(c, (size_t)max_sweeps + 1);
        RC_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int) * (max_sweeps + 1), c->stream));
        T* Gp = G.p; T* Vp = V.p;
        int rows_i = (int)rows, n_i = (int)n, ms = max_sweeps;
        double* sigp = sig.p; int* flp = flags.p; int* infp = info.p;
        void* args[] = {&Gp, &Vp, &rows_i, &n_i, &ms, &tol, &flp, &u, &ldu, &s, &w, &ldw, &sigp, &infp};
        RC_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_grid_kernel<T>, dim3(grid), dim3(JG_NT), args, 0, c->stream));
        RC_COUNT_LAUNCH(c);
    }
#undef RC_JACOBI
    if (c->trace) {
        int h[2] = {0, 0};
        RC_CUDA(cudaMemcpyAsync(h, info.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        RC_CUDA(cudaStreamSynchronize(c->stream));
        fprintf(stderr, "[rc trace]     jacobi %lld x %lld: %d sweeps%s\n", (long long)rows, (long long)n, h[1], h[0] ? " (NOT converged)" : "");
    }
}

template void jacobi_svd<float>(rc_ctx*, const float*, int64_t, int64_t, int64_t, float*, int64_t, double*, float*, int64_t, int*);
template void jacobi_svd<double>(rc_ctx*, const double*, int64_t, int64_t, int64_t, double*, int64_t, double*, double*, int64_t, int*);
template void jacobi_svd<c32>(rc_ctx*, const c32*, int64_t, int64_t, int64_t, c32*, int64_t, double*, c32*, int64_t, int*);
template void jacobi_svd<c64>(rc_ctx*, const c64*, int64_t, int64_t, int64_t, c64*, int64_t, double*, c64*, int64_t, int*);
