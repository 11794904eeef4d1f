// One-sided (Hestenes) Jacobi SVD of a small square matrix: the B200 replacement for LAPACK
// ?gesdd jobz='S' (reference N3: src/compute_svd.rs:19) after the k x n factor has been reduced
// to k x k by a TSQR of its conjugate transpose (host_linalg.cu: svd_impl).
//
// G (n x n, column-major) is rotated from the right until its columns are mutually orthogonal:
// G J = U diag(s), J accumulated in V.  One warp per column pair, round-robin ordering, all pairs
// of a round in parallel; Gram entries accumulated in double.  Singular values come out with
// high relative accuracy, sorted descending like ?gesdd's.
#include "rc_internal.cuh"

namespace {

constexpr int JT = 1024;
constexpr int JW = JT / 32;

template <class T, bool SMEM>
__global__ void __launch_bounds__(JT)
jacobi_kernel(T* __restrict__ Gg, T* __restrict__ Vg, int rows, int n, int max_sweeps, double tol,
              T* __restrict__ u, int64_t ldu, double* __restrict__ s_out, T* __restrict__ w, int64_t ldw,
              double* __restrict__ sig_scratch, int* __restrict__ info) {
    __shared__ int s_rot;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // SMEM: G (rows x n) and V (n x n) are resident in shared memory for the whole iteration
    T* G = SMEM ? reinterpret_cast<T*>(smem_raw) : Gg;
    T* V = SMEM ? G + (size_t)rows * n : Vg;
    if (SMEM) {
        for (int e = tid; e < rows * n; e += JT) G[e] = Gg[e];
    }
    const int npad = n + (n & 1);
    const int half = npad / 2;
    // V = I
    for (int e = tid; e < n * n; e += JT) {
        int c = e / n, r = e - c * n;
        V[e] = (r == c) ? rc_one<T>() : rc_zero<T>();
    }
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < npad - 1; ++round) {
            for (int k = warp; k < half; k += JW) {
                int a, b;
                if (k == 0) { a = npad - 1; b = round; }
                else { a = (round + k) % (npad - 1); b = (round - k + (npad - 1)) % (npad - 1); }
                int p = min(a, b), q = max(a, b);
                if (q >= n) continue;                      // padding column
                T* gp = G + (int64_t)p * rows;
                T* gq = G + (int64_t)q * rows;
                double alpha = 0.0, beta = 0.0, gre = 0.0, gim = 0.0;
                for (int r = lane; r < rows; r += 32) {
                    T x = gp[r], y = gq[r];
                    alpha += rc_abs2(x);
                    beta += rc_abs2(y);
                    // gamma = conj(x) * y
                    double xr = (double)rc_real(x), xi = (double)rc_imag(x);
                    double yr = (double)rc_real(y), yi = (double)rc_imag(y);
                    gre += xr * yr + xi * yi;
                    gim += xr * yi - xi * yr;
                }
                alpha = rc_warp_sum(alpha); beta = rc_warp_sum(beta);
                gre = rc_warp_sum(gre); gim = rc_warp_sum(gim);
                // real scalars: gamma is real, no hypot (a ~200-cycle call at the head of the rotation chain)
                const double gabs = ScalarTraits<T>::is_complex ? hypot(gre, gim) : fabs(gre);
                if (gabs == 0.0 || gabs <= tol * sqrt(alpha * beta)) continue;
                if (lane == 0) s_rot = 1;
                double zeta = (beta - alpha) / (2.0 * gabs);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = rsqrt(1.0 + t * t), sn = cs * t;
                // e^{-i theta} = conj(gamma)/|gamma|
                double er = gre / gabs, ei = -gim / gabs;
                T ph = rc_make<T>(er, ei);
                T csT = rc_make<T>(cs, 0.0), snT = rc_make<T>(sn, 0.0);
                for (int r = lane; r < rows; r += 32) {
                    T x = gp[r], y = ph * gq[r];
                    gp[r] = csT * x - snT * y;
                    gq[r] = snT * x + csT * y;
                }
                T* vp = V + (int64_t)p * n;
                T* vq = V + (int64_t)q * n;
                for (int r = lane; r < n; r += 32) {
                    T x = vp[r], y = ph * vq[r];
                    vp[r] = csT * x - snT * y;
                    vq[r] = snT * x + csT * y;
                }
            }
            __syncthreads();
        }
        if (s_rot == 0) break;
        __syncthreads();
    }
    if (tid == 0) info[0] = (sweep >= max_sweeps) ? 1 : 0;
    // singular values
    for (int c = warp; c < n; c += JW) {
        const T* gc = G + (int64_t)c * rows;
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
        const T* gc = G + (int64_t)c * rows;
        const T* vc = V + (int64_t)c * n;
        double inv = (sc > 0.0) ? 1.0 / sc : 0.0;
        T invT = rc_make<T>(inv, 0.0);
        for (int r = lane; r < rows; r += 32) u[(int64_t)r * ldu + rank] = gc[r] * invT;
        for (int r = lane; r < n; r += 32) w[(int64_t)r * ldw + rank] = vc[r];
        if (lane == 0) s_out[rank] = sc;
    }
}

}  // namespace

template <class T>
void jacobi_svd(rc_ctx* c, const T* g, int64_t ldg, int64_t rows, int64_t n, T* u, int64_t ldu, double* s, T* w, int64_t ldw) {
    RC_REQUIRE(n > 0 && n <= 8192 && rows >= n, "jacobi_svd: unsupported size %lld x %lld", (long long)rows, (long long)n);
    DevBuf<T> G(c, (size_t)rows * n), V(c, (size_t)n * n);
    DevBuf<double> sig(c, (size_t)n);
    DevBuf<int> info(c, 1);
    // column-major copy of g == transpose of the row-major matrix
    k_transpose<T>(c, G.p, rows, g, ldg, rows, n, false);
    double eps = (sizeof(RealOf<T>) == 4) ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
    double tol = eps * sqrt((double)rows);
    size_t smem = ((size_t)rows * n + (size_t)n * n) * sizeof(T);
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    if (smem + 4096 <= lim) {
        RC_CUDA(cudaFuncSetAttribute(jacobi_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_kernel<T, true><<<1, JT, smem, c->stream>>>(G.p, V.p, (int)rows, (int)n, 60, tol, u, ldu, s, w, ldw, sig.p, info.p);
    } else {
        jacobi_kernel<T, false><<<1, JT, 0, c->stream>>>(G.p, V.p, (int)rows, (int)n, 60, tol, u, ldu, s, w, ldw, sig.p, info.p);
    }
    RC_CHECK_LAUNCH(c);
}

template void jacobi_svd<float>(rc_ctx*, const float*, int64_t, int64_t, int64_t, float*, int64_t, double*, float*, int64_t);
template void jacobi_svd<double>(rc_ctx*, const double*, int64_t, int64_t, int64_t, double*, int64_t, double*, double*, int64_t);
template void jacobi_svd<c32>(rc_ctx*, const c32*, int64_t, int64_t, int64_t, c32*, int64_t, double*, c32*, int64_t);
template void jacobi_svd<c64>(rc_ctx*, const c64*, int64_t, int64_t, int64_t, c64*, int64_t, double*, c64*, int64_t);
