// Microbenchmark: does the FP64 pipe of sm_100a run DFMA and DMMA back to back without a switching penalty?
// Each warp iteration issues ND independent DMMA.8x8x4 and NF independent DFMA; prints the measured pipe cycles
// per warp-iteration and sub-partition next to the model 16 * ND + 2 * NF.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rusty_compression_b200/build/mix_fp64 tools/micro/mix_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ND, int NF, bool SPREAD>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, double seed) {
    double c[ND > 0 ? ND : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = seed + i + threadIdx.x;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 * seed, x = 1.0000001, y = 1e-9 * seed;
    for (int it = 0; it < iters; ++it) {
        if (SPREAD) {
#pragma unroll
            for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
                if (i < ND) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
                if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(x), "d"(y));
            }
        } else {
#pragma unroll
            for (int i = 0; i < ND; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(x), "d"(y));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}

template <int ND, int NF, bool SPREAD>
void run(double* out, int sms, double clk_ghz, int warps_per_sm) {
    const int iters = 4000;
    const int ctas = sms * warps_per_sm / 8;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mix_kernel<ND, NF, SPREAD><<<ctas, 256>>>(out, iters, 1.0);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        mix_kernel<ND, NF, SPREAD><<<ctas, 256>>>(out, iters, 1.0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    // warps per sub-partition = warps_per_sm / 4; pipe cycles per warp-iteration = t * clk / (iters * warps per sub-partition)
    double cyc = best * 1e-3 * clk_ghz * 1e9 / ((double)iters * warps_per_sm / 4.0);
    printf("ND=%2d NF=%2d %s warps/SM=%2d : %7.1f cycles per warp-iteration (model %d)\n", ND, NF, SPREAD ? "spread " : "grouped",
           warps_per_sm, cyc, 16 * ND + 2 * NF);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount, clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ghz = clk / 1e6;
    double* out; CK(cudaMalloc(&out, 64));
    for (int w : {16, 64}) {
        run<20, 0, false>(out, sms, ghz, w);
        run<16, 0, false>(out, sms, ghz, w);
        run<0, 8, false>(out, sms, ghz, w);
        run<0, 16, false>(out, sms, ghz, w);
        run<16, 8, false>(out, sms, ghz, w);
        run<16, 8, true>(out, sms, ghz, w);
        run<16, 16, true>(out, sms, ghz, w);
        run<18, 4, false>(out, sms, ghz, w);
    }
    return 0;
}
