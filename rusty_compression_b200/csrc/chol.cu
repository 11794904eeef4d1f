// Small Cholesky factor + triangular inverse for the Cholesky-QR2 fast path of the tall sketches
// (host_api.cu: cholqr2).  One CTA, the w x w Gram matrix resident in shared memory.
//   G = R^H R (R upper triangular, real positive diagonal);  Rinv = R^{-1}.
// status[0] = 0 ok / 1 breakdown (non-positive pivot), status[1] = min diag(R), status[2] = max diag(R),
// status[3] = max |G - I| of the input (orthogonality defect when G is a Gram matrix of a Q factor).
#include "rc_internal.cuh"
#include "host_linalg.cuh"

namespace {

constexpr int CT = 1024;

constexpr int NB = 16;        // block size of the factorisation and of the triangular inverse

// Blocked right-looking Cholesky G = L L^H (L = R^H lower triangular) and blocked triangular inverse X = L^{-1}.
// The unblocked kernel this replaces paid one block barrier and one dependent scalar chain per COLUMN (74 of each for
// the 74 x 74 Gram matrix of the config-2 sketch: ~90 us, on the critical path of every tall QR); here the serial part
// is one 16 x 16 diagonal block per 16 columns, factored and inverted by one warp, and everything else is
// GEMM-shaped work for the whole block:
//   per block column:  diagonal block (warp 0: Cholesky, then its inverse)            | barrier
//                      panel   L21 = A21 L11^{-H}   (one output per thread)           | 2 barriers (in place)
//                      update  A22 -= L21 L21^H                                       | barrier
//   inverse:           level d = 1 .. nblk-1 (block sub-diagonal d, all blocks of a level at once):
//                      T = sum_k L_ik X_kj ; X_ij = -X_ii T                           | 2 barriers per level
template <class T>
__global__ void __launch_bounds__(CT)
chol_inv_kernel(const T* __restrict__ g, int64_t ldg, int w, T* __restrict__ r, T* __restrict__ rinv, int64_t ldo, double* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = w | 1;                            // odd pitch: a column of L / X touches every bank once
    T* L = reinterpret_cast<T*>(smem_raw);           // lower triangle: the factor
    T* X = L + (size_t)w * ld;                       // lower triangle: its inverse; the upper triangle is scratch
    __shared__ double s_red[CT / 32];
    __shared__ double s_dinv[256];                   // 1 / L_jj (w <= 256 is guaranteed by the shared-memory limit)
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    using R = RealOf<T>;
    double defect = 0.0;
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        T v = g[(int64_t)i * ldg + j];
        if (j <= i) L[i * ld + j] = v;
        T d = (i == j) ? v - rc_one<T>() : v;
        defect = fmax(defect, rc_abs(d));
    }
    if (tid == 0) s_bad = 0;
    for (int m = 16; m > 0; m >>= 1) defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, m));
    if (lane == 0) s_red[warp] = defect;
    __syncthreads();
    if (tid == 0) { double d = 0.0; for (int i = 0; i < CT / 32; ++i) d = fmax(d, s_red[i]); status[3] = d; }

    const int nblk = (w + NB - 1) / NB;
    for (int kb = 0; kb < nblk; ++kb) {
        const int j0 = kb * NB, jb = min(NB, w - j0);
        // ---- diagonal block: unblocked Cholesky (lane = row), then the inverse of the block (lane = column)
        if (warp == 0) {
            // The block lives in REGISTERS (lane = row, a[c] = L[row][c]) and rows talk through shuffles: a
            // read-modify-write loop over shared memory serialises on every store (the compiler must assume that the
            // next load aliases it), which made this 16-step chain the most expensive part of the kernel.
            T a[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) a[c] = (lane < jb && c <= lane) ? L[(j0 + lane) * ld + j0 + c] : rc_zero<T>();
            double dinv_mine = 1.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < jb) {                                       // warp-uniform
                    double d = __shfl_sync(0xffffffffu, (double)rc_real(a[c]), c);
                    const bool bad = !(d > 0.0);
                    if (bad) d = 1.0;
                    const double rd = rsqrt(d);
                    if (lane == c) { a[c] = rc_make<T>(d * rd, 0.0); dinv_mine = rd; if (bad) s_bad = 1; }
                    else if (lane > c) a[c] = a[c] * (R)rd;
#pragma unroll
                    for (int c2 = c + 1; c2 < NB; ++c2) {
                        if (c2 < jb) {
                            const T l2c = rc_shfl(a[c], c2);        // L[c2][c]
                            if (lane >= c2) a[c2] = a[c2] - a[c] * rc_conj(l2c);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) if (lane < jb && c <= lane) L[(j0 + lane) * ld + j0 + c] = a[c];
            if (lane < jb) s_dinv[j0 + lane] = dinv_mine;
            __syncwarp();
            // inverse of the block, lane = column: x[rr] = -(1 / L_rr) sum_{k < rr} L[rr][k] x[k]  (x[k] = 0 for k < lane)
            T x[NB];
#pragma unroll
            for (int rr = 0; rr < NB; ++rr) {
                x[rr] = rc_zero<T>();
                if (rr < jb) {
                    T acc = rc_zero<T>();
#pragma unroll
                    for (int k = 0; k < rr; ++k) acc = rc_fma(L[(j0 + rr) * ld + j0 + k], x[k], acc);
                    const R dr = (R)s_dinv[j0 + rr];
                    x[rr] = (rr == lane) ? rc_make<T>((double)dr, 0.0) : ((rr > lane) ? -(acc * dr) : rc_zero<T>());
                }
            }
#pragma unroll
            for (int rr = 0; rr < NB; ++rr) if (lane < jb && rr >= lane && rr < jb) X[(j0 + rr) * ld + j0 + lane] = x[rr];
        }
        __syncthreads();
        const int i0 = j0 + jb, nrow = w - i0;
        if (nrow <= 0) break;
        // ---- panel: L21 = A21 X11^H, out[i][c] = sum_{c' <= c} A[i][c'] conj(X11[c][c']); one output per thread, held in
        //      a register across the barrier because the panel is overwritten in place
        {
            T outv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = tid + u * CT;
                outv[u] = rc_zero<T>();
                if (e < nrow * jb) {
                    const int i = i0 + e / jb, c = e % jb;
                    T acc = rc_zero<T>();
                    for (int c1 = 0; c1 <= c; ++c1) acc = rc_fma(L[i * ld + j0 + c1], rc_conj(X[(j0 + c) * ld + j0 + c1]), acc);
                    outv[u] = acc;
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = tid + u * CT;
                if (e < nrow * jb) L[(i0 + e / jb) * ld + j0 + e % jb] = outv[u];
            }
            __syncthreads();
        }
        // ---- trailing update (lower triangle): A22[i][j] -= sum_c L21[i][c] conj(L21[j][c])
        for (int i = i0 + warp; i < w; i += CT / 32)
            for (int j = i0 + lane; j <= i; j += 32) {
                T acc = L[i * ld + j];
                for (int c = 0; c < jb; ++c) acc = acc - L[i * ld + j0 + c] * rc_conj(L[j * ld + j0 + c]);
                L[i * ld + j] = acc;
            }
        __syncthreads();
    }
    __syncthreads();
    // ---- X = L^{-1}: the diagonal blocks are in place; block sub-diagonal d needs only the sub-diagonals < d
    for (int d = 1; d < nblk; ++d) {
        const int nb_lvl = nblk - d;                                 // blocks (i = j + d, j), j = 0 .. nblk-1-d
        // T = sum_k L_ik X_kj, stored transposed in the (unused) upper triangle of X
        for (int e = tid; e < nb_lvl * NB * NB; e += CT) {
            const int j = e / (NB * NB), rr = (e / NB) % NB, c = e % NB;
            const int row = (j + d) * NB + rr, col = j * NB + c;
            if (row < w) {
                T acc = rc_zero<T>();
                for (int kk = col; kk < (j + d) * NB; ++kk) acc = rc_fma(L[row * ld + kk], X[kk * ld + col], acc);
                X[col * ld + row] = acc;
            }
        }
        __syncthreads();
        // X_ij = -X_ii T
        for (int e = tid; e < nb_lvl * NB * NB; e += CT) {
            const int j = e / (NB * NB), rr = (e / NB) % NB, c = e % NB;
            const int ib = (j + d) * NB, row = ib + rr, col = j * NB + c;
            if (row < w) {
                T acc = rc_zero<T>();
                for (int r1 = 0; r1 <= rr; ++r1) acc = rc_fma(X[row * ld + ib + r1], X[col * ld + ib + r1], acc);
                X[row * ld + col] = -acc;
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        double dmin = 1e300, dmax = 0.0;
        for (int j = lane; j < w; j += 32) { double pv = (double)rc_real(L[j * ld + j]); dmin = fmin(dmin, pv); dmax = fmax(dmax, pv); }
        for (int m = 16; m > 0; m >>= 1) { dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, m)); dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, m)); }
        if (lane == 0) { status[0] = (double)s_bad; status[1] = dmin; status[2] = dmax; }
    }
    // R = L^H (upper triangular, zeros below), R^{-1} = X^H
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        r[(int64_t)i * ldo + j] = (j >= i) ? rc_conj(L[j * ld + i]) : rc_zero<T>();
        rinv[(int64_t)i * ldo + j] = (j >= i) ? rc_conj(X[j * ld + i]) : rc_zero<T>();
    }
}

}  // namespace

// r, rinv: w x w row-major (ld = ldo).  Returns false if w does not fit in shared memory.
template <class T>
bool chol_inv(rc_ctx* c, const T* g, int64_t ldg, int64_t w, T* r, T* rinv, int64_t ldo, double* status_dev) {
    size_t smem = 2 * (size_t)w * (size_t)(w | 1) * sizeof(T);
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    if (smem + 8192 > lim || w > 256) return false;
    RC_CUDA(cudaFuncSetAttribute(chol_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_inv_kernel<T><<<1, CT, smem, c->stream>>>(g, ldg, (int)w, r, rinv, ldo, status_dev);
    RC_CHECK_LAUNCH(c);
    return true;
}
namespace {
// status words of the 2 x 2 blocked factorisation from those of its two diagonal blocks and the full Gram matrix:
// {breakdown in either block, min diag R, max diag R, max |G - I| over the whole w x w input}
template <class T>
__global__ void __launch_bounds__(256)
chol_blocked_status_kernel(const T* __restrict__ g, int64_t ldg, int w, const double* __restrict__ sa, const double* __restrict__ sb,
                           double* __restrict__ status) {
    __shared__ double s_red[8];
    double defect = 0.0;
    for (int e = threadIdx.x; e < w * w; e += 256) {
        const int i = e / w, j = e - i * w;
        const T v = g[(int64_t)i * ldg + j];
        const T d = (i == j) ? v - rc_one<T>() : v;
        defect = fmax(defect, rc_abs(d));
    }
    for (int m = 16; m > 0; m >>= 1) defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, m));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = defect;
    __syncthreads();
    if (threadIdx.x == 0) {
        double d = 0.0;
        for (int i = 0; i < 8; ++i) d = fmax(d, s_red[i]);
        status[0] = (sa[0] != 0.0 || sb[0] != 0.0) ? 1.0 : 0.0;
        status[1] = fmin(sa[1], sb[1]);
        status[2] = fmax(sa[2], sb[2]);
        status[3] = d;
    }
}
}  // namespace

// Gram matrices wider than one CTA's shared memory (w <= 2 x chol_max_width: the 266-column f32 sketch of config 4, the
// 138-column c64 sketch of config 5): one level of 2 x 2 blocking around the one-CTA kernel,
//   G = [G11 G12; G12^H G22]:  R11 = chol(G11),  R12 = R11^{-H} G12,  R22 = chol(G22 - R12^H R12),
//   R^{-1} = [R11^{-1}, -R11^{-1} R12 R22^{-1}; 0, R22^{-1}],
// all the off-diagonal work in (small) GEMMs.  Same status words as chol_inv (breakdown of either block is a breakdown).
template <class T>
bool chol_inv_blocked(rc_ctx* c, const T* g, int64_t ldg, int64_t w, T* r, T* rinv, int64_t ldo, double* status_dev) {
    if (chol_inv<T>(c, g, ldg, w, r, rinv, ldo, status_dev)) return true;
    const int dtype = ScalarTraits<T>::code;       // rc_dtype: RC_F32 = 0, RC_F64 = 1, RC_C32 = 2, RC_C64 = 3
    const int64_t wmax = chol_max_width(c, dtype);
    if (w > 2 * wmax || w > 512) return false;
    const int64_t e = std::max<int64_t>(1, (int64_t)(16 / sizeof(T)));
    const int64_t w1 = std::min<int64_t>(wmax / e * e, ((w + 1) / 2 + e - 1) / e * e), w2 = w - w1;
    if (w2 <= 0 || w2 > wmax) return false;
    const int64_t l1 = rc_pad_ld(dtype, w1), l2 = rc_pad_ld(dtype, w2);
    DevBuf<T> r11(c, (size_t)w1 * l1), ri11(c, (size_t)w1 * l1), r12(c, (size_t)w1 * l2), s22(c, (size_t)w2 * l2),
        r22(c, (size_t)w2 * l2), ri22(c, (size_t)w2 * l2), t12(c, (size_t)w1 * l2), x12(c, (size_t)w1 * l2);
    DevBuf<double> st(c, 8);
    if (!chol_inv<T>(c, g, ldg, w1, r11.p, ri11.p, l1, st.p)) return false;
    // R12 = R11^{-H} G12
    gemm<T>(c, RC_OP_H, RC_OP_N, w1, w2, w1, ri11.p, l1, g + w1, ldg, r12.p, l2, rc_one<T>(), rc_zero<T>());
    // S = G22 - R12^H R12
    gemm<T>(c, RC_OP_H, RC_OP_N, w2, w2, w1, r12.p, l2, r12.p, l2, s22.p, l2, rc_one<T>(), rc_zero<T>());
    k_sub<T>(c, s22.p, l2, g + w1 * ldg + w1, ldg, s22.p, l2, w2, w2);
    if (!chol_inv<T>(c, s22.p, l2, w2, r22.p, ri22.p, l2, st.p + 4)) return false;
    // X12 = -R11^{-1} R12 R22^{-1}
    gemm<T>(c, RC_OP_N, RC_OP_N, w1, w2, w1, ri11.p, l1, r12.p, l2, t12.p, l2, rc_one<T>(), rc_zero<T>());
    gemm<T>(c, RC_OP_N, RC_OP_N, w1, w2, w2, t12.p, l2, ri22.p, l2, x12.p, l2, rc_zero<T>() - rc_one<T>(), rc_zero<T>());
    // assemble R and R^{-1} (upper block-triangular, zeros below)
    k_fill<T>(c, r, w, w, ldo, rc_zero<T>());
    k_fill<T>(c, rinv, w, w, ldo, rc_zero<T>());
    k_copy<T>(c, r, ldo, r11.p, l1, w1, w1);
    k_copy<T>(c, r + w1, ldo, r12.p, l2, w1, w2);
    k_copy<T>(c, r + w1 * ldo + w1, ldo, r22.p, l2, w2, w2);
    k_copy<T>(c, rinv, ldo, ri11.p, l1, w1, w1);
    k_copy<T>(c, rinv + w1, ldo, x12.p, l2, w1, w2);
    k_copy<T>(c, rinv + w1 * ldo + w1, ldo, ri22.p, l2, w2, w2);
    chol_blocked_status_kernel<T><<<1, 256, 0, c->stream>>>(g, ldg, (int)w, st.p, st.p + 4, status_dev);
    RC_CHECK_LAUNCH(c);
    return true;
}

// G += factor * max_j Re(G_jj) * I  (the shift of the first round of shifted Cholesky-QR3, host_api.cu: cholqr2)
namespace {
template <class T>
__global__ void __launch_bounds__(256) chol_shift_kernel(T* __restrict__ g, int64_t ldg, int w, double factor) {
    __shared__ double s_red[8];
    double mx = 0.0;
    for (int j = threadIdx.x; j < w; j += 256) mx = fmax(mx, (double)rc_real(g[(int64_t)j * ldg + j]));
    for (int m = 16; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int i = 0; i < 8; ++i) mx = fmax(mx, s_red[i]);
    const T shift = rc_make<T>(factor * mx, 0.0);
    for (int j = threadIdx.x; j < w; j += 256) g[(int64_t)j * ldg + j] = g[(int64_t)j * ldg + j] + shift;
}
}  // namespace
template <class T>
void chol_shift(rc_ctx* c, T* g, int64_t ldg, int64_t w, double factor) {
    chol_shift_kernel<T><<<1, 256, 0, c->stream>>>(g, ldg, (int)w, factor);
    RC_CHECK_LAUNCH(c);
}
template void chol_shift<float>(rc_ctx*, float*, int64_t, int64_t, double);
template void chol_shift<double>(rc_ctx*, double*, int64_t, int64_t, double);
template void chol_shift<c32>(rc_ctx*, c32*, int64_t, int64_t, double);
template void chol_shift<c64>(rc_ctx*, c64*, int64_t, int64_t, double);

int64_t chol_max_width(rc_ctx* c, int dtype) {
    size_t lim = (c->smem_optin ? c->smem_optin : (size_t)227 * 1024) - 8192;
    int64_t w = 1;
    while (2 * (size_t)(w + 1) * (w + 1) * rc_dtype_size(dtype) <= lim) ++w;
    return w;
}

template bool chol_inv<float>(rc_ctx*, const float*, int64_t, int64_t, float*, float*, int64_t, double*);
template bool chol_inv<double>(rc_ctx*, const double*, int64_t, int64_t, double*, double*, int64_t, double*);
template bool chol_inv<c32>(rc_ctx*, const c32*, int64_t, int64_t, c32*, c32*, int64_t, double*);
template bool chol_inv<c64>(rc_ctx*, const c64*, int64_t, int64_t, c64*, c64*, int64_t, double*);
template bool chol_inv_blocked<float>(rc_ctx*, const float*, int64_t, int64_t, float*, float*, int64_t, double*);
template bool chol_inv_blocked<double>(rc_ctx*, const double*, int64_t, int64_t, double*, double*, int64_t, double*);
template bool chol_inv_blocked<c32>(rc_ctx*, const c32*, int64_t, int64_t, c32*, c32*, int64_t, double*);
template bool chol_inv_blocked<c64>(rc_ctx*, const c64*, int64_t, int64_t, c64*, c64*, int64_t, double*);
