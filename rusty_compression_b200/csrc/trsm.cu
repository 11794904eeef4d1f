// Batched upper-triangular solve with all right-hand sides at once: the replacement for the
// crate's loop of one ?trtrs call per column/row (reference N4: src/qr.rs:290-301, 384-395).
// X = U^{-1} B, one thread per right-hand side, the current row of U staged in shared memory.
#include "rc_internal.cuh"

namespace {

constexpr int TT = 128;

template <class T, bool UT>
__global__ void __launch_bounds__(TT)
trsm_upper_kernel(const T* __restrict__ u, int64_t ldu, int k, const T* __restrict__ b, int64_t ldb,
                  int64_t nrhs, T* __restrict__ x, int64_t ldx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* urow = reinterpret_cast<T*>(smem_raw);
    const int64_t j = (int64_t)blockIdx.x * TT + threadIdx.x;
    for (int i = k - 1; i >= 0; --i) {
        for (int jj = i + threadIdx.x; jj < k; jj += TT)
            urow[jj] = UT ? u[(int64_t)jj * ldu + i] : u[(int64_t)i * ldu + jj];
        __syncthreads();
        if (j < nrhs) {
            T acc = b[(int64_t)i * ldb + j];
            for (int jj = i + 1; jj < k; ++jj) acc = acc - urow[jj] * x[(int64_t)jj * ldx + j];
            x[(int64_t)i * ldx + j] = acc / urow[i];
        }
        __syncthreads();
    }
}


// Same solve with the solution tile resident in shared memory: the global version above re-reads every
// computed x from L2 once per remaining row (k^2 / 2 * nrhs elements: 5.2 GB at config 3), this one reads and
// writes global memory once.  TTS right-hand sides per CTA, xs[k][TTS] + the current row of U in shared memory;
// four partial sums per thread break the k-long dependent FMA chain.
template <class T, bool UT, int TTS>
__global__ void __launch_bounds__(TTS)
trsm_upper_smem_kernel(const T* __restrict__ u, int64_t ldu, int k, const T* __restrict__ b, int64_t ldb,
                       int64_t nrhs, T* __restrict__ x, int64_t ldx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* urow = reinterpret_cast<T*>(smem_raw);               // k entries
    T* xs = urow + k;                                        // k x TTS
    const int t = threadIdx.x;
    const int64_t j = (int64_t)blockIdx.x * TTS + t;
    for (int i = k - 1; i >= 0; --i) {
        for (int jj = i + t; jj < k; jj += TTS)
            urow[jj] = UT ? u[(int64_t)jj * ldu + i] : u[(int64_t)i * ldu + jj];
        __syncthreads();
        if (j < nrhs) {
            T a0 = b[(int64_t)i * ldb + j], a1 = rc_zero<T>(), a2 = rc_zero<T>(), a3 = rc_zero<T>();
            int jj = i + 1;
            for (; jj + 3 < k; jj += 4) {
                a0 = a0 - urow[jj] * xs[(size_t)jj * TTS + t];
                a1 = a1 - urow[jj + 1] * xs[(size_t)(jj + 1) * TTS + t];
                a2 = a2 - urow[jj + 2] * xs[(size_t)(jj + 2) * TTS + t];
                a3 = a3 - urow[jj + 3] * xs[(size_t)(jj + 3) * TTS + t];
            }
            for (; jj < k; ++jj) a0 = a0 - urow[jj] * xs[(size_t)jj * TTS + t];
            const T v = ((a0 + a1) + (a2 + a3)) / urow[i];
            xs[(size_t)i * TTS + t] = v;
            x[(int64_t)i * ldx + j] = v;
        }
        __syncthreads();
    }
}

}  // namespace

template <class T>
void trsm_upper(rc_ctx* c, const T* u, int64_t ldu, bool u_transposed, int64_t k, const T* b, int64_t ldb,
                int64_t nrhs, T* x, int64_t ldx) {
    if (k == 0 || nrhs == 0) return;
    const size_t lim = (c->smem_optin ? c->smem_optin : (size_t)227 * 1024) - 4096;
#define RC_TRSM_SMEM(TTS)                                                                                         \
    if ((size_t)k * (TTS + 1) * sizeof(T) <= lim) {                                                               \
        size_t sm = (size_t)k * (TTS + 1) * sizeof(T);                                                            \
        unsigned nbs = (unsigned)((nrhs + TTS - 1) / TTS);                                                        \
        if (u_transposed) {                                                                                       \
            RC_CUDA(cudaFuncSetAttribute(trsm_upper_smem_kernel<T, true, TTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            trsm_upper_smem_kernel<T, true, TTS><<<nbs, TTS, sm, c->stream>>>(u, ldu, (int)k, b, ldb, nrhs, x, ldx); \
        } else {                                                                                                  \
            RC_CUDA(cudaFuncSetAttribute(trsm_upper_smem_kernel<T, false, TTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            trsm_upper_smem_kernel<T, false, TTS><<<nbs, TTS, sm, c->stream>>>(u, ldu, (int)k, b, ldb, nrhs, x, ldx); \
        }                                                                                                         \
        RC_CHECK_LAUNCH(c);                                                                                       \
        return;                                                                                                   \
    }
    // widest tile of right-hand sides whose solution block fits into shared memory (and still gives every SM work)
    if (nrhs >= 128 * (int64_t)c->sm_count) { RC_TRSM_SMEM(128) }
    RC_TRSM_SMEM(64)
    RC_TRSM_SMEM(32)
#undef RC_TRSM_SMEM
    size_t smem = (size_t)k * sizeof(T);
    unsigned nb = (unsigned)((nrhs + TT - 1) / TT);
    if (u_transposed) {
        RC_CUDA(cudaFuncSetAttribute(trsm_upper_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trsm_upper_kernel<T, true><<<nb, TT, smem, c->stream>>>(u, ldu, (int)k, b, ldb, nrhs, x, ldx);
    } else {
        RC_CUDA(cudaFuncSetAttribute(trsm_upper_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trsm_upper_kernel<T, false><<<nb, TT, smem, c->stream>>>(u, ldu, (int)k, b, ldb, nrhs, x, ldx);
    }
    RC_CHECK_LAUNCH(c);
}

template void trsm_upper<float>(rc_ctx*, const float*, int64_t, bool, int64_t, const float*, int64_t, int64_t, float*, int64_t);
template void trsm_upper<double>(rc_ctx*, const double*, int64_t, bool, int64_t, const double*, int64_t, int64_t, double*, int64_t);
template void trsm_upper<c32>(rc_ctx*, const c32*, int64_t, bool, int64_t, const c32*, int64_t, int64_t, c32*, int64_t);
template void trsm_upper<c64>(rc_ctx*, const c64*, int64_t, bool, int64_t, const c64*, int64_t, int64_t, c64*, int64_t);
