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
    const int tid = threadIdx.x;
    double defect = 0.0;
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        T v = g[(int64_t)i * ldg + j];
        S[e] = v;
        T d = (i == j) ? v - rc_one<T>() : v;
        defect = fmax(defect, rc_abs(d));
    }
    // block max of the defect
    for (int m = 16; m > 0; m >>= 1) defect = fmax(defect, __shfl_xor_sync(0xffffffffu, defect, m));
    if ((tid & 31) == 0) s_red[tid >> 5] = defect;
    __syncthreads();
    if (tid == 0) { double d = 0.0; for (int i = 0; i < CT / 32; ++i) d = fmax(d, s_red[i]); status[3] = d; }
    // Row-by-row (left-looking) Cholesky by the first 128 threads, one thread per column:
    //   R[j][c] = (G[j][c] - sum_{l<j} conj(R[l][j]) R[l][c]) / R[j][j]
    // The l-loop reads R[l][j] as a broadcast and R[l][c] conflict-free; the two barriers per row are
    // 128-thread named barriers (the other 28 warps wait at the block barrier below).
    __shared__ double s_d;
    if (tid < 128) {
        double dmin = 1e300, dmax = 0.0;
        int bad = 0;
        for (int j = 0; j < w; ++j) {
            for (int c = j + tid; c < w; c += 128) {
                T a0 = S[j * w + c], a1 = rc_zero<T>();
                int l = 0;
                for (; l + 1 < j; l += 2) {
                    a0 = a0 - rc_conj(S[l * w + j]) * S[l * w + c];
                    a1 = a1 - rc_conj(S[(l + 1) * w + j]) * S[(l + 1) * w + c];
                }
                if (l < j) a0 = a0 - rc_conj(S[l * w + j]) * S[l * w + c];
                a0 = a0 + a1;
                S[j * w + c] = a0;
                if (c == j) s_d = (double)rc_real(a0);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            double d = s_d;
            if (!(d > 0.0)) { bad = 1; d = 1.0; }
            const double piv = sqrt(d);
            dmin = fmin(dmin, piv); dmax = fmax(dmax, piv);
            const RealOf<T> ip = (RealOf<T>)(1.0 / piv);
            for (int c = j + tid; c < w; c += 128) S[j * w + c] = (c == j) ? rc_make<T>(piv, 0.0) : S[j * w + c] * ip;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (tid == 0) { status[0] = (double)bad; status[1] = dmin; status[2] = dmax; }
    }
    __syncthreads();
    // R out (upper triangular, zeros below)
    for (int e = tid; e < w * w; e += CT) {
        int i = e / w, j = e - i * w;
        r[(int64_t)i * ldo + j] = (j >= i) ? S[e] : rc_zero<T>();
    }
    // X = R^{-1}: columns are independent; one warp per column walks its rows upwards, the lanes
    // split each inner product (a thread-per-column loop was a 2.7K-long dependent chain)
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int c = warp; c < w; c += CT / 32) {
            for (int i = w - 1 - lane; i > c; i -= 32) X[i * w + c] = rc_zero<T>();
            if (lane == 0) X[c * w + c] = rc_one<T>() / S[c * w + c];
            __syncwarp();
            for (int i = c - 1; i >= 0; --i) {
                T acc = rc_zero<T>();
                for (int l = i + 1 + lane; l <= c; l += 32) acc = rc_fma(S[i * w + l], X[l * w + c], acc);
                acc = rc_warp_sum(acc);
                if (lane == 0) X[i * w + c] = -(acc / S[i * w + i]);
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
    if (smem + 8192 > lim) return false;
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
