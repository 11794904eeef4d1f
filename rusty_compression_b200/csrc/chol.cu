// Small Cholesky factor + triangular inverse for the Cholesky-QR2 fast path of the tall sketches
// (host_api.cu: cholqr2).  One CTA, the w x w Gram matrix resident in shared memory.
//   G = R^H R (R upper triangular, real positive diagonal);  Rinv = R^{-1}.
// status[0] = 0 ok / 1 breakdown (non-positive pivot), status[1] = min diag(R), status[2] = max diag(R),
// status[3] = max |G - I| of the input (orthogonality defect when G is a Gram matrix of a Q factor).
#include "rc_internal.cuh"

namespace {

constexpr int CT = 1024;

template <class T>
__global__ void __launch_bounds__(CT)
chol_inv_kernel(const T* __restrict__ g, int64_t ldg, int w, T* __restrict__ r, T* __restrict__ rinv, int64_t ldo, double* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);          // w x w row-major working copy (upper part used)
    T* X = S + (size_t)w * w;                        // inverse
    __shared__ double s_red[CT / 32];
    __shared__ double s_piv[257];                    // pivots d_j = R_jj^2 (w <= 256 is guaranteed by the smem limit)
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double defect = 0.0;
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        T v = g[(int64_t)i * ldg + j];
        S[e] = v;
        T d = (i == j) ? v - rc_one<T>() : v;
        defect = fmax(defect, rc_abs(d));
    }
    if (tid == 0) s_bad = 0;
    // block max of the defect
    for (int m = 16; m > 0; m >>= 1) defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, m));
    if (lane == 0) s_red[warp] = defect;
    __syncthreads();
    if (tid == 0) { double d = 0.0; for (int i = 0; i < CT / 32; ++i) d = fmax(d, s_red[i]); status[3] = d; }
    // Right-looking (outer-product) Cholesky with ONE barrier per step: step j updates the trailing upper
    // triangle with the UNSCALED row j, S[r][c] -= conj(S[j][r]) S[j][c] / d_j, so no thread has to wait for the
    // scaled pivot row; the rows are scaled by 1 / sqrt(d_j) once at the end.
    // The reciprocal of the NEXT pivot is computed by the thread that finishes S[j+1][j+1] (its first element
    // of the step), so the ~200-cycle double division overlaps that thread's remaining updates instead of
    // sitting at the head of every step for everybody.
    __shared__ double s_invd[2];
    if (tid == 0) {
        const double d0 = (double)rc_real(S[0]);
        s_piv[0] = d0; if (!(d0 > 0.0)) s_bad = 1;
        s_invd[0] = 1.0 / ((d0 > 0.0) ? d0 : 1.0);
    }
    __syncthreads();
    {
        const int tx = lane, ty = warp;              // 32 x 32 thread tile (measured: a 16 x 16 tile with a
        for (int j = 0; j < w; ++j) {                // 256-thread named barrier is 30 % slower at w = 74)
            const RealOf<T> id = (RealOf<T>)s_invd[j & 1];
            for (int rr = j + 1 + ty; rr < w; rr += 32) {
                const T f = rc_conj(S[j * w + rr]) * id;
                for (int cc = j + 1 + tx; cc < w; cc += 32)
                    if (cc >= rr) {
                        const T v = S[rr * w + cc] - f * S[j * w + cc];
                        S[rr * w + cc] = v;
                        if (rr == j + 1 && cc == j + 1) {        // thread 0, first element: the next pivot
                            const double d = (double)rc_real(v);
                            s_piv[j + 1] = d; if (!(d > 0.0)) s_bad = 1;
                            s_invd[(j + 1) & 1] = 1.0 / ((d > 0.0) ? d : 1.0);
                        }
                    }
            }
            __syncthreads();
        }
    }
    // scale the rows: R[j][c] = S[j][c] / sqrt(d_j); diagonal real positive
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        if (j >= i) {
            const double d = s_piv[i];
            const double piv = sqrt((d > 0.0) ? d : 1.0);
            S[e] = (i == j) ? rc_make<T>(piv, 0.0) : S[e] * (RealOf<T>)(1.0 / piv);
        }
    }
    __syncthreads();
    if (warp == 0) {
        double dmin = 1e300, dmax = 0.0;
        for (int j = lane; j < w; j += 32) { double pv = (double)rc_real(S[j * w + j]); dmin = fmin(dmin, pv); dmax = fmax(dmax, pv); }
        for (int m = 16; m > 0; m >>= 1) { dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, m)); dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, m)); }
        if (lane == 0) { status[0] = (double)s_bad; status[1] = dmin; status[2] = dmax; }
    }
    // R out (upper triangular, zeros below)
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        r[(int64_t)i * ldo + j] = (j >= i) ? S[e] : rc_zero<T>();
    }
    // X = R^{-1}: columns are independent (no block barriers); four lanes share a column and split each
    // inner product, so the dependent chain of a column is c steps of (c - i) / 4 FMAs + two shuffles.  The
    // eight columns of a warp run a common trip count (the shuffles need the whole warp).
    {
        const int sub = lane & 3;
        for (int cbase = warp * 8; cbase < w; cbase += (CT / 32) * 8) {
            const int c = cbase + (lane >> 2);
            const bool active = c < w;
            const int cmax = min(cbase + 7, w - 1);
            if (active) {
                for (int i = w - 1 - sub; i > c; i -= 4) X[i * w + c] = rc_zero<T>();
                if (sub == 0) X[c * w + c] = rc_one<T>() / S[c * w + c];
            }
            __syncwarp();
            for (int i = cmax - 1; i >= 0; --i) {
                const bool work = active && i < c;
                T acc = rc_zero<T>();
                if (work)
                    for (int l = i + 1 + sub; l <= c; l += 4) acc = rc_fma(S[i * w + l], X[l * w + c], acc);
                acc = acc + rc_shfl_xor(acc, 1);
                acc = acc + rc_shfl_xor(acc, 2);
                if (work && sub == 0) X[i * w + c] = -(acc / S[i * w + i]);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < w * w; e += CT) { int i = e / w, j = e - i * w; rinv[(int64_t)i * ldo + j] = X[e]; }
}

}  // namespace

// r, rinv: w x w row-major (ld = ldo).  Returns false if w does not fit in shared memory.
template <class T>
bool chol_inv(rc_ctx* c, const T* g, int64_t ldg, int64_t w, T* r, T* rinv, int64_t ldo, double* status_dev) {
    size_t smem = 2 * (size_t)w * w * sizeof(T);
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    if (smem + 8192 > lim || w > 256) return false;
    RC_CUDA(cudaFuncSetAttribute(chol_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_inv_kernel<T><<<1, CT, smem, c->stream>>>(g, ldg, (int)w, r, rinv, ldo, status_dev);
    RC_CHECK_LAUNCH(c);
    return true;
}
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
