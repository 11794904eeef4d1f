// Stand-alone microbenchmark: own-measured FP64 peaks of this B200 (the driver's
// MEASURED_PEAKS.json has only HBM copy and bf16 tensor).  Prints one JSON line:
//   {"dfma_tflops": .., "dmma_tflops": .., "copy_gbs": .., "sm_count": .., "clock_mhz": ..}
// Used by bench.py as the denominator of the FP64 roofline ("of own-measured").
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    double x = 1.0000001, y = 1e-9 * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double c[12][2];
#pragma unroll
    for (int i = 0; i < 12; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    const int iters = 20000, ctas = sms * 8;
    float t1 = time_ms([&] { dfma_kernel<<<ctas, 256>>>(out, iters, 1.0); }, 5);
    double dfma = 2.0 * 16 * (double)iters * 256.0 * ctas / (t1 * 1e-3) / 1e12;
    float t2 = time_ms([&] { dmma_kernel<<<ctas, 256>>>(out, iters, 1.0); }, 5);
    double dmma = 2.0 * 256.0 * 12 * (double)iters * 8.0 * ctas / (t2 * 1e-3) / 1e12;
    size_t n = (size_t)1 << 27;   // 4 GiB each way
    double4 *a, *b; CK(cudaMalloc(&a, n * 32)); CK(cudaMalloc(&b, n * 32));
    CK(cudaMemset(a, 1, n * 32));
    float t3 = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
    double copy = 2.0 * n * 32 / (t3 * 1e-3) / 1e9;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"copy_gbs\": %.1f, \"sm_count\": %d, \"clock_mhz\": %.0f}\n",
           dfma, dmma, copy, sms, clk / 1000.0);
    return 0;
}
