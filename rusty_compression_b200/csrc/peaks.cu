// Stand-alone microbenchmark: own-measured FP64 and tcgen05 peaks of this B200 (the driver's
// MEASURED_PEAKS.json has only HBM copy and cuBLAS bf16).  Prints one JSON line:
//   {"dfma_tflops": .., "dmma_tflops": .., "tf32_umma_tflops": .., "bf16_umma_tflops": .., "copy_gbs": .., "sm_count": .., "clock_mhz": ..}
// Used by bench.py as the denominator of the FP64 and TF32 rooflines ("of own-measured").  The tcgen05 figures
// are issue-rate peaks: one CTA per SM streams 128 x 256 x K MMAs (operands in shared memory, accumulators in
// TMEM) with no data movement at all, so they bound what any kind::tf32 / kind::f16 kernel can reach here.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

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

// ---- tcgen05 issue-rate peak: kind 0 = kind::tf32 (K = 8 per MMA), kind 1 = kind::f16 with bf16 operands (K = 16)
__device__ __forceinline__ uint32_t pk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool pk_elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
    return pred != 0;
}
// n = UMMA N (multiple of 16, <= 256); ts = 1: A operand from tensor memory (the TS form the TF32-split kernel uses);
// one_acc = 1: every MMA accumulates into the same TMEM tile (a GEMM main loop), 0: two tiles alternate
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters, int kind, unsigned* out, int n = 256, int ts = 0, int one_acc = 0) {
    extern __shared__ unsigned char pk_smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (pk_smem_u32(pk_smem) + 1023u) & ~1023u;            // A: 128 rows x 128 B, B: 256 rows x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (48 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(pk_smem + (base - pk_smem_u32(pk_smem)))[i] = 0u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        const uint64_t hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;     // K-major, SWIZZLE_128B, SBO 1024
        const uint64_t adesc = hi | (uint64_t)((base & 0x3FFFFu) >> 4), bdesc = hi | (uint64_t)(((base + 16384u) & 0x3FFFFu) >> 4);
        const uint32_t fmt = kind == 0 ? 2u : 1u;                                         // TF32 / BF16
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
        for (int it = 0; it < iters; ++it) {
            if (pk_elect_one()) {
                const uint32_t d = tmem + (uint32_t)((one_acc || ts) ? 0 : (it & 1) * 256);
                const uint32_t a_tm = tmem + 256u;          // TS form: 32 columns of (zero) A at column 256
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t acc = (it > 1 || k > 0) ? 1u : 0u;
                    if (kind == 0 && ts)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                                     ::"r"(d), "r"(a_tm + 8u * k), "l"(bdesc + 2u * k), "r"(idesc), "r"(acc) : "memory");
                    else if (kind == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d), "l"(adesc + 2u * k), "l"(bdesc + 2u * k), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(d), "l"(adesc + 2u * k), "l"(bdesc + 2u * k), "r"(idesc), "r"(acc) : "memory");
                }
            }
            __syncwarp();
        }
        if (pk_elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pk_smem_u32(&bar)) : "memory");
        __syncwarp();
        asm volatile("{\n.reg .pred p;\nPK_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra PK_DONE;\nbra PK_WAIT;\nPK_DONE:\n}\n"
                     ::"r"(pk_smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    if (out && tid == 0 && iters < 0) out[0] = tmem;
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

int main(int argc, char** argv) {
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
    // tcgen05 issue-rate peaks (sm_100 only)
    double tf32 = 0.0, bf16 = 0.0;
    if (prop.major == 10) {
        const int umma_iters = 8192;
        const size_t pk_bytes = 48 * 1024 + 1024;
        CK(cudaFuncSetAttribute(umma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pk_bytes));
        float t4 = time_ms([&] { umma_peak_kernel<<<sms, 128, pk_bytes>>>(umma_iters, 0, nullptr); }, 5);
        tf32 = 2.0 * 128 * 256 * 8 * 4.0 * umma_iters * sms / (t4 * 1e-3) / 1e12;
        float t5 = time_ms([&] { umma_peak_kernel<<<sms, 128, pk_bytes>>>(umma_iters, 1, nullptr); }, 5);
        bf16 = 2.0 * 128 * 256 * 16 * 4.0 * umma_iters * sms / (t5 * 1e-3) / 1e12;
    }
    if (argc > 1 && prop.major == 10) {
        // rc_peaks shapes: kind::tf32 MMA time by N and operand form (what bounds the TF32-split kernel at small l)
        const int umma_iters = 8192;
        const size_t pk_bytes = 48 * 1024 + 1024;
        printf("kind::tf32 M = 128, K = 8 per MMA; cycles per MMA at %.0f MHz and TFLOP/s over %d SMs\n", clk / 1000.0, sms);
        for (int form = 0; form < 3; ++form)
            for (int nn : {32, 64, 96, 128, 192, 256}) {
                const int ts_ = form == 2, one = form >= 1;
                float t = time_ms([&] { umma_peak_kernel<<<sms, 128, pk_bytes>>>(umma_iters, 0, nullptr, nn, ts_, one); }, 5);
                const double per = (double)t * 1e-3 * (clk * 1e3) / (4.0 * umma_iters);
                printf("  %-34s N = %3d: %6.1f cycles per MMA, %7.1f TFLOP/s\n",
                       form == 0 ? "SS, two accumulators alternating" : (form == 1 ? "SS, one accumulator" : "TS (A from TMEM), one accumulator"),
                       nn, per, 2.0 * 128 * nn * 8 * 4.0 * umma_iters * sms / (t * 1e-3) / 1e12);
            }
    }
    printf("{\"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"tf32_umma_tflops\": %.1f, \"bf16_umma_tflops\": %.1f, \"copy_gbs\": %.1f, "
           "\"sm_count\": %d, \"clock_mhz\": %.0f}\n", dfma, dmma, tf32, bf16, copy, sms, clk / 1000.0);
    return 0;
}
