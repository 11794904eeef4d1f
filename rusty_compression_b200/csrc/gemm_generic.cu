// Generic SIMT tiled GEMM for all four scalars and all operand modes (N / T / H), with a
// deterministic two-phase split-K for skinny outputs with a long contraction
// (Z = A^H X, B = Q^H A: reduction over the m rows).  This is the correctness workhorse that
// replaces ndarray's `.dot` (reference N5: src/types.rs:119,129-131; src/qr.rs:76,161,277,288,...).
// The FP64 tensor-pipe kernels in gemm_dmma.cu take over the large f64/c64 contractions.
#include "rc_internal.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <class T, int OPA, int OPB>
__global__ void __launch_bounds__(256)
gemm_kernel(int64_t M, int64_t N, int64_t K, const T* __restrict__ A, int64_t lda,
            const T* __restrict__ B, int64_t ldb, T* __restrict__ C, int64_t ldc, T alpha, T beta,
            int64_t kchunk, T* __restrict__ partial) {
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
    const int64_t kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = rc_zero<T>();

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage op(A) tile: As[k][i] = op(A)[m0+i][k0+k]
#pragma unroll
        for (int r = 0; r < (BM * BK) / 256; ++r) {
            int e = tid + r * 256;
            int i, k;
            if (OPA == RC_OP_N) { k = e % BK; i = e / BK; } else { i = e % BM; k = e / BM; }
            int64_t gi = m0 + i, gk = k0 + k;
            T v = rc_zero<T>();
            if (gi < M && gk < kend) {
                if (OPA == RC_OP_N) v = A[gi * lda + gk];
                else { v = A[gk * lda + gi]; if (OPA == RC_OP_H) v = rc_conj(v); }
            }
            As[k][i] = v;
        }
        // ---- stage op(B) tile: Bs[k][j] = op(B)[k0+k][n0+j]
#pragma unroll
        for (int r = 0; r < (BN * BK) / 256; ++r) {
            int e = tid + r * 256;
            int j, k;
            if (OPB == RC_OP_N) { j = e % BN; k = e / BN; } else { k = e % BK; j = e / BK; }
            int64_t gj = n0 + j, gk = k0 + k;
            T v = rc_zero<T>();
            if (gj < N && gk < kend) {
                if (OPB == RC_OP_N) v = B[gk * ldb + gj];
                else { v = B[gj * ldb + gk]; if (OPB == RC_OP_H) v = rc_conj(v); }
            }
            Bs[k][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = rc_fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t gi = m0 + ty + 16 * i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t gj = n0 + tx + 16 * j;
            if (gj >= N) continue;
            if (partial) {
                partial[((int64_t)blockIdx.z * M + gi) * N + gj] = acc[i][j];
            } else {
                T v = alpha * acc[i][j];
                if (!(rc_real(beta) == RealOf<T>(0) && rc_imag(beta) == RealOf<T>(0))) v = v + beta * C[gi * ldc + gj];
                C[gi * ldc + gj] = v;
            }
        }
    }
}

template <class T>
__global__ void splitk_reduce_kernel(int64_t M, int64_t N, int splits, const T* __restrict__ partial,
                                     T* __restrict__ C, int64_t ldc, T alpha, T beta) {
    int64_t n = M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        T s = rc_zero<T>();
        for (int z = 0; z < splits; ++z) s = s + partial[(int64_t)z * n + e];   // fixed order: deterministic
        int64_t i = e / N, j = e - i * N;
        T v = alpha * s;
        if (!(rc_real(beta) == RealOf<T>(0) && rc_imag(beta) == RealOf<T>(0))) v = v + beta * C[i * ldc + j];
        C[i * ldc + j] = v;
    }
}

template <class T, int OPA, int OPB>
void launch(rc_ctx* c, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda, const T* B, int64_t ldb,
            T* C, int64_t ldc, T alpha, T beta) {
    int64_t gm = (M + BM - 1) / BM, gn = (N + BN - 1) / BN;
    int64_t tiles = gm * gn;
    int splits = 1;
    if (K >= 2048 && tiles < 2 * c->sm_count) {
        int64_t want = (4 * (int64_t)c->sm_count + tiles - 1) / tiles;
        int64_t maxs = K / 512;
        splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxs), 64));
    }
    int64_t kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    if (kchunk <= 0) kchunk = BK;
    splits = (int)((K + kchunk - 1) / kchunk);
    if (splits < 1) splits = 1;
    RC_REQUIRE(gn <= 65535, "gemm_generic: N too large for grid.y (%lld)", (long long)N);
    dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
    if (splits == 1) {
        gemm_kernel<T, OPA, OPB><<<grid, 256, 0, c->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kchunk, nullptr);
        RC_CHECK_LAUNCH(c);
    } else {
        DevBuf<T> part(c, (size_t)splits * M * N);
        gemm_kernel<T, OPA, OPB><<<grid, 256, 0, c->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kchunk, part.p);
        RC_CHECK_LAUNCH(c);
        int64_t n = M * N;
        int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
        splitk_reduce_kernel<T><<<nb, 256, 0, c->stream>>>(M, N, splits, part.p, C, ldc, alpha, beta);
        RC_CHECK_LAUNCH(c);
    }
}

}  // namespace

template <class T>
void gemm_generic(rc_ctx* c, RcOp opa, RcOp opb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda,
                  const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta) {
    if (M == 0 || N == 0) return;
    if (!ScalarTraits<T>::is_complex) {   // H == T for real scalars: halve the instantiations
        if (opa == RC_OP_H) opa = RC_OP_T;
        if (opb == RC_OP_H) opb = RC_OP_T;
    }
#define RC_CASE(a, b)                                                                               \
    if (opa == a && opb == b) { launch<T, a, b>(c, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta); return; }
    RC_CASE(RC_OP_N, RC_OP_N)
    RC_CASE(RC_OP_N, RC_OP_T)
    RC_CASE(RC_OP_T, RC_OP_N)
    RC_CASE(RC_OP_T, RC_OP_T)
    if constexpr (ScalarTraits<T>::is_complex) {
        RC_CASE(RC_OP_N, RC_OP_H)
        RC_CASE(RC_OP_H, RC_OP_N)
        RC_CASE(RC_OP_H, RC_OP_H)
        RC_CASE(RC_OP_T, RC_OP_H)
        RC_CASE(RC_OP_H, RC_OP_T)
    }
#undef RC_CASE
    RC_THROW(RC_INVALID_ARGUMENT, "gemm_generic: bad op combination");
}

template void gemm_generic<float>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float, float);
template void gemm_generic<double>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, double, double);
template void gemm_generic<c32>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const c32*, int64_t, const c32*, int64_t, c32*, int64_t, c32, c32);
template void gemm_generic<c64>(rc_ctx*, RcOp, RcOp, int64_t, int64_t, int64_t, const c64*, int64_t, const c64*, int64_t, c64*, int64_t, c64, c64);
