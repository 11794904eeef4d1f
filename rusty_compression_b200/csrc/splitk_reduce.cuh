// Fixed-order reduction of split-K partials, shared by the DMMA (f64) and tcgen05 (f32) GEMM front ends:
//   c[i][j] = sum_z part[z * part_stride + i * ldp + j],   z = 0 .. splits-1.
// The summation order depends only on (splits, shape), never on scheduling, so results are deterministic.
#pragma once
#include "rc_internal.cuh"

namespace rc_splitk {

// Few splits, many elements (Z = A^T Y with a long output): one thread per element, splits summed serially.
template <class T>
__global__ void reduce_serial_kernel(int64_t M, int64_t N, int splits, const T* __restrict__ part, int64_t ldp,
                                     int64_t part_stride, T* __restrict__ c, int64_t ldc) {
    const int64_t n = M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / N, j = e - i * N;
        T s = T(0);
        for (int z = 0; z < splits; ++z) s += part[(int64_t)z * part_stride + i * ldp + j];
        c[i * ldc + j] = s;
    }
}

// Many splits, few elements (the l x l Gram matrices of the Cholesky-QR path: 256-296 splits of a 74 x 74
// output): the serial kernel above keeps only ~20 SMs busy with a dependent chain of `splits` adds per thread
// (62 us for 256 x 74 x 74 f64).  Here a CTA owns 32 consecutive elements; split lane y sums the splits
// z = y, y + 32, ... and the 32 lane sums are added in lane order.
template <class T>
__global__ void __launch_bounds__(1024)
reduce_wide_kernel(int64_t M, int64_t N, int splits, const T* __restrict__ part, int64_t ldp,
                   int64_t part_stride, T* __restrict__ c, int64_t ldc) {
    __shared__ T sh[32][33];
    const int64_t n = M * N;
    const int64_t e = blockIdx.x * 32LL + threadIdx.x;
    int64_t i = 0, j = 0;
    T s = T(0);
    if (e < n) {
        i = e / N; j = e - i * N;
        const T* p = part + i * ldp + j;
        for (int z = threadIdx.y; z < splits; z += 32) s += p[(int64_t)z * part_stride];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < n) {
        T t = T(0);
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sh[y][threadIdx.x];
        c[i * ldc + j] = t;
    }
}

template <class T>
inline void reduce(rc_ctx* c, int64_t M, int64_t N, int splits, const T* part, int64_t ldp, int64_t part_stride,
                   T* out, int64_t ldc) {
    const int64_t n = M * N;
    if (splits >= 16 && n <= (1LL << 22)) {
        reduce_wide_kernel<T><<<(unsigned)((n + 31) / 32), dim3(32, 32), 0, c->stream>>>(M, N, splits, part, ldp, part_stride, out, ldc);
    } else {
        const int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
        reduce_serial_kernel<T><<<nb, 256, 0, c->stream>>>(M, N, splits, part, ldp, part_stride, out, ldc);
    }
    RC_CHECK_LAUNCH(c);
}

}  // namespace rc_splitk
