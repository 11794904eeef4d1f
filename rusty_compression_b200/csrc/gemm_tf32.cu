// f32 contraction Y = A X on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// 3-product split  A X ~= A_hi X_hi + A_lo X_hi + A_hi X_lo  (hi = value rounded to TF32, lo = the
// exact f32 remainder), FP32 accumulation in TMEM.  This is what carries f32 parity (1e-4 in
// north_star; measured ~1e-6) on the tensor pipe instead of the FP32 SIMT pipe.
//
// Reference role: MatMat::matmat for f32 operators (src/types.rs:58-71), the dominant cost of
// configs 3 and 4.
//
// Structure (one CTA per SM, persistent over 128-row tiles of A):
//   warp 0      TMA producer: raw f32 A tile [128 x 32] + pre-split X^T tiles (hi, lo) [N x 32],
//               SWIZZLE_128B, 2-4-deep mbarrier ring
//   warps 2-5   splitter: A tile -> A_hi (in place) and A_lo (second buffer), element positions
//               preserved so the TMA-written canonical K-major layout stays valid for UMMA;
//               fence.proxy.async, then signal the MMA warp.  Splitting A on the fly avoids a
//               pre-split copy that would double the HBM traffic of the pass.
//   warp 1      MMA issuer: one elected thread issues 12 tcgen05.mma (3 products x 4 K-steps of 8)
//               per stage, tcgen05.commit frees the stage; accumulator double-buffered in TMEM
//   warps 6-9   epilogue: tcgen05.ld (32 lanes x 16 columns) -> registers -> global Y
// X is tiny (n x l): it is transposed and split once by a prologue kernel so that both B operands
// are K-major TMA tiles.
//
// TRANS mode (Z = A^T Y, reduction over the rows of A, split-K with a fixed-order reduction): the
// raw tile arrives as four [32 k][32 i] boxes and the splitter transposes it while splitting, so
// the MMA still sees K-major operands (an MN-major descriptor variant produced all-zero products
// on this toolchain and was dropped).
#include <cuda.h>
#include <cstdlib>
#include "rc_internal.cuh"

namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1)
constexpr int BK = 32;            // floats per stage = one 128-byte swizzle row
constexpr int NTHREADS = 320;     // 10 warps
constexpr int MAX_STAGES = 4;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
            RC_THROW(RC_CUDA_ERROR, "cuTensorMapEncodeTiled entry point not available");
        fn = (EncodeFn)p;
    }
    return fn;
}
// Row-major [rows][cols] f32 matrix, box = 32 cols (128 B) x box_rows, SWIZZLE_128B.
CUtensorMap make_map_f32(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    RC_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map box rows out of range");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RC_THROW(RC_CUDA_ERROR, "cuTensorMapEncodeTiled (f32) failed (%d)", (int)r);
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 128-byte rows, 8-row atoms of 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);           // start address  [0,14)
    d |= (uint64_t)0 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset [32,46): 8 rows x 128 B
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32 (both K-major), M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t n) {
    uint32_t d = 0;
    d |= 1u << 4;            // c_format = F32
    d |= 2u << 7;            // a_format = TF32
    d |= 2u << 10;           // b_format = TF32
    d |= (n >> 3) << 17;     // n_dim
    d |= (uint32_t)(BM >> 4) << 24;   // m_dim
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Tf32Params {
    float* y;
    int64_t ldy;
    int M, N, K;          // N = real number of columns
    int npad;             // UMMA N (multiple of 16, <= 256)
    int m_tiles;
    uint32_t tmem_cols;   // power of two >= 2 * npad
    // TRANS (Z = A^T Y): split-K over the rows of A; partial results at y + split * part_stride
    int splits, kb_per_split;
    int64_t part_stride;
};

// KC k-blocks (KC * 32 values of K) are accumulated inside the tensor core before the partial sum is
// promoted to an FP32 register accumulator (round-to-nearest adds on the CUDA cores): the tensor
// core's own FP32 accumulation truncates, which costs ~K * 2^-24 relative accuracy on long chains
// (measured 2.9e-5 at K = 4096 without promotion).
constexpr int KC = 8;

template <int STAGES, int NPADC, bool TRANS>
__global__ void __launch_bounds__(NTHREADS, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, Tf32Params prm) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long bars[3 * MAX_STAGES + 4];
    __shared__ uint32_t tmem_base_smem;
    const uint32_t full0 = smem_u32(&bars[0]);                   // TMA landed (raw A + B hi/lo)
    const uint32_t split0 = smem_u32(&bars[MAX_STAGES]);         // A split into hi/lo
    const uint32_t empty0 = smem_u32(&bars[2 * MAX_STAGES]);     // MMAs that read the stage are done
    const uint32_t accf0 = smem_u32(&bars[3 * MAX_STAGES]);      // accumulator buffer full (2)
    const uint32_t acce0 = smem_u32(&bars[3 * MAX_STAGES + 2]);  // accumulator buffer drained (2)

    constexpr int npad = NPADC;
    const uint32_t A_BYTES = BM * BK * 4;                        // 16 KB
    const uint32_t B_BYTES = (uint32_t)npad * BK * 4;
    // stage layout: [raw A (TRANS only)] [A_hi] [A_lo] [B_hi] [B_lo]; in NN mode the raw tile is split in place
    const uint32_t RAW_BYTES = TRANS ? A_BYTES : 0u;
    const uint32_t STAGE_BYTES = RAW_BYTES + 2 * A_BYTES + 2 * B_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 4);        // one arrival per splitter warp
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {     // TMEM allocation (whole warp), address published through shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(prm.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    const int kblocks_all = (prm.K + BK - 1) / BK;
    const int total_items = prm.m_tiles * prm.splits;      // NN: splits == 1
    // work item -> (tile along M, k-block range)
#define RC_ITEM(t)                                                                         \
    const int mt_ = (t) % prm.m_tiles, sp_ = (t) / prm.m_tiles;                            \
    const int kb0_ = sp_ * prm.kb_per_split;                                               \
    const int kb1_ = min(kblocks_all, kb0_ + prm.kb_per_split);                            \
    const int kblocks = kb1_ - kb0_;                                                       \
    (void)mt_; (void)sp_; (void)kblocks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                RC_ITEM(t)
                const int m0 = mt_ * BM;
                for (int kb = kb0_; kb < kb1_; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, A_BYTES + 2 * B_BYTES);
                    if (!TRANS) {
                        tma_load_2d(sa, &tmA, kb * BK, m0, fb);                                   // [128 rows][32 k]
                    } else {
#pragma unroll
                        for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &tmA, m0 + 32 * b, kb * BK, fb);   // raw [32 k][32 i] x 4
                    }
                    tma_load_2d(sa + RAW_BYTES + 2 * A_BYTES, &tmBhi, kb * BK, 0, fb);            // [npad rows][32 k]
                    tma_load_2d(sa + RAW_BYTES + 2 * A_BYTES + B_BYTES, &tmBlo, kb * BK, 0, fb);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32((uint32_t)npad);
            int stage = 0; uint32_t phase = 0;
            int it = 0;                                              // counts K-chunks (TMEM buffer hand-offs)
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                RC_ITEM(t)
                for (int kc0 = 0; kc0 < kblocks; kc0 += KC, ++it) {
                    const int buf = it & 1;
                    const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
                    mbar_wait(acce0 + 8 * buf, acc_phase ^ 1u);      // epilogue drained this buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * npad);
                    const int kc1 = min(kblocks, kc0 + KC);
                    for (int kb = kc0; kb < kc1; ++kb) {
                        mbar_wait(split0 + 8 * stage, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_base + stage * STAGE_BYTES;
                        const uint32_t a_hi = sa + RAW_BYTES, a_lo = a_hi + A_BYTES, b_hi = a_lo + A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            // K-major: 8 floats along K inside the swizzle row; MN-major: next group of 8 k-rows
                            const uint32_t koff = (uint32_t)k * 32u;
                            const uint64_t dah = umma_desc_sw128(a_hi + koff), dal = umma_desc_sw128(a_lo + koff);
                            const uint64_t dbh = umma_desc_sw128(b_hi + koff), dbl = umma_desc_sw128(b_lo + koff);
                            // small terms first, then the dominant one
                            umma_tf32(d_tmem, dal, dbh, idesc, (kb != kc0 || k != 0) ? 1u : 0u);
                            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                            umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                        }
                        umma_commit(empty0 + 8 * stage);             // frees the stage when the MMAs retire
                        if (kb == kc1 - 1) umma_commit(accf0 + 8 * buf);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ===================== splitter: A -> (A_hi in place, A_lo) =====================
        const int st = tid - 64;                                     // 0..127
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
            RC_ITEM(t)
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                unsigned char* base = smem_dyn + (smem_base - smem_u32(smem_dyn)) + (size_t)stage * STAGE_BYTES;
                if (!TRANS) {
                    float4* raw = reinterpret_cast<float4*>(base);
                    float4* lo = reinterpret_cast<float4*>(base + A_BYTES);
#pragma unroll
                    for (int i = 0; i < (int)(A_BYTES / 16 / 128); ++i) {
                        const int idx = st + 128 * i;
                        float4 v = raw[idx], h, l;
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
                        raw[idx] = h;
                        lo[idx] = l;
                    }
                } else {
                    // Transposing split: the raw tile is 4 boxes of [32 k][32 i] (the contraction index is the
                    // row index of A).  Thread `st` owns output row i = st of the K-major [128 i][32 k] tiles:
                    // 32 conflict-free LDS.32 (a warp reads one 128-byte row per k), split, then 2 x 8 STS.128
                    // into the SWIZZLE_128B positions UMMA expects for a K-major operand.
                    const int bx = st >> 5, ci = st & 31;
                    const unsigned char* rawb = base + bx * 4096;
                    float4* hi = reinterpret_cast<float4*>(base + A_BYTES);
                    float4* lo = reinterpret_cast<float4*>(base + 2 * A_BYTES);
                    float vals[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        vals[k] = *reinterpret_cast<const float*>(rawb + k * 128 + ((((ci >> 2) ^ (k & 7)) << 4) | ((ci & 3) << 2)));
#pragma unroll
                    for (int cch = 0; cch < 8; ++cch) {
                        float4 h, l;
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vals[4 * cch + 0])); h.x = __uint_as_float(u); l.x = vals[4 * cch + 0] - h.x;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vals[4 * cch + 1])); h.y = __uint_as_float(u); l.y = vals[4 * cch + 1] - h.y;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vals[4 * cch + 2])); h.z = __uint_as_float(u); l.z = vals[4 * cch + 2] - h.z;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vals[4 * cch + 3])); h.w = __uint_as_float(u); l.w = vals[4 * cch + 3] - h.w;
                        const int dst = st * 8 + (cch ^ (st & 7));       // float4 index: row st, swizzled 16-byte chunk
                        hi[dst] = h;
                        lo[dst] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) mbar_arrive(split0 + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int lg = warp & 3;                                     // TMEM lane group this warp may access
        int it = 0;
        for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
            RC_ITEM(t)
            float acc[NPADC];
#pragma unroll
            for (int j = 0; j < NPADC; ++j) acc[j] = 0.f;
            for (int kc0 = 0; kc0 < kblocks; kc0 += KC, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
                mbar_wait(accf0 + 8 * buf, acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < NPADC; c0 += 16) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * npad + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(acce0 + 8 * buf);
            }
            const int row = mt_ * BM + lg * 32 + lane;
            if (row < prm.M) {
                float* yrow = prm.y + (int64_t)sp_ * prm.part_stride + (int64_t)row * prm.ldy;
#pragma unroll
                for (int j = 0; j < NPADC; ++j)
                    if (j < prm.N) yrow[j] = acc[j];
            }
        }
    }
#undef RC_ITEM
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols) : "memory");
    }
}

// X (K x N row-major) -> XhiT, XloT (npad x K row-major, K contiguous), rows >= N zero.
__global__ void split_transpose_kernel(const float* __restrict__ x, int64_t ldx, int K, int N, int npad,
                                       float* __restrict__ hiT, float* __restrict__ loT, int64_t ldt) {
    int64_t total = (int64_t)npad * K;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(e / K), k = (int)(e - (int64_t)j * K);
        float v = (j < N) ? x[(int64_t)k * ldx + j] : 0.f;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        float h = __uint_as_float(u);
        hiT[(int64_t)j * ldt + k] = h;
        loT[(int64_t)j * ldt + k] = v - h;
    }
}

template <int STAGES, int NPADC, bool TRANS>
void launch_tf32(rc_ctx* c, const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
                 const Tf32Params& prm) {
    constexpr size_t stage_bytes = (TRANS ? 3 : 2) * (size_t)BM * BK * 4 + 2 * (size_t)NPADC * BK * 4;
    constexpr size_t smem = STAGES * stage_bytes + 1024;
    static_assert(smem <= 227 * 1024, "stage configuration exceeds shared memory");
    RC_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<STAGES, NPADC, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = std::min(prm.m_tiles * prm.splits, c->sm_count);
    tf32x3_gemm_kernel<STAGES, NPADC, TRANS><<<grid, NTHREADS, smem, c->stream>>>(tmA, tmBhi, tmBlo, prm);
    RC_CHECK_LAUNCH(c);
}
template <bool TRANS>
void dispatch_tf32(rc_ctx* c, int npad, const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
                   const Tf32Params& prm) {
    switch (npad) {
        case 32: launch_tf32<(TRANS ? 3 : 4), 32, TRANS>(c, tmA, tmBhi, tmBlo, prm); break;
        case 64: launch_tf32<(TRANS ? 3 : 4), 64, TRANS>(c, tmA, tmBhi, tmBlo, prm); break;
        case 96: launch_tf32<(TRANS ? 3 : 4), 96, TRANS>(c, tmA, tmBhi, tmBlo, prm); break;
        default: launch_tf32<(TRANS ? 2 : 3), 128, TRANS>(c, tmA, tmBhi, tmBlo, prm); break;
    }
}

__global__ void tf32_reduce_kernel(int64_t M, int N, int splits, const float* __restrict__ part, int64_t ldp, int64_t part_stride,
                                   float* __restrict__ z, int64_t ldz) {
    int64_t n = M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / N; int j = (int)(e - i * N);
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += part[(int64_t)sp * part_stride + i * ldp + j];   // fixed order
        z[i * ldz + j] = s;
    }
}

}  // namespace

// Y (M x N, ldy) = A (M x K, lda) * X (K x N, ldx), f32, 3xTF32 on tcgen05.  Returns false when the
// shape / alignment is not supported (caller falls back to the SIMT kernel).
bool gemm_tf32x3_f32(rc_ctx* c, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                     const float* X, int64_t ldx, float* Y, int64_t ldy) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 3)) return false;      // TMA: 16-byte base and pitch
    if (M > (1LL << 30) || K > (1LL << 30)) return false;
    // columns are processed in chunks of at most 128 (the promoted FP32 accumulator of a row lives in
    // the registers of one epilogue thread)
    int64_t nchunks = (N + 127) / 128;
    int64_t per = ((N + nchunks - 1) / nchunks + 31) / 32 * 32;
    int64_t ldt = (K + 3) / 4 * 4;
    for (int64_t n0 = 0; n0 < N; n0 += per) {
        const int ncols = (int)std::min<int64_t>(per, N - n0);
        const int npad = (ncols + 31) / 32 * 32;          // 32, 64, 96 or 128
        DevBuf<float> hiT(c, (size_t)npad * ldt), loT(c, (size_t)npad * ldt);
        {
            int64_t total = (int64_t)npad * K;
            int nb = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
            split_transpose_kernel<<<nb, 256, 0, c->stream>>>(X + n0, ldx, (int)K, ncols, npad, hiT.p, loT.p, ldt);
            RC_CHECK_LAUNCH(c);
        }
        CUtensorMap tmA = make_map_f32(A, M, K, lda, BM);
        CUtensorMap tmBhi = make_map_f32(hiT.p, npad, K, ldt, npad);
        CUtensorMap tmBlo = make_map_f32(loT.p, npad, K, ldt, npad);
        Tf32Params prm;
        prm.y = Y + n0; prm.ldy = ldy; prm.M = (int)M; prm.N = ncols; prm.K = (int)K; prm.npad = npad;
        prm.m_tiles = (int)((M + BM - 1) / BM);
        uint32_t cols = 32;
        while (cols < (uint32_t)(2 * npad)) cols <<= 1;
        prm.tmem_cols = cols;
        prm.splits = 1; prm.kb_per_split = (int)((K + BK - 1) / BK); prm.part_stride = 0;
        dispatch_tf32<false>(c, npad, tmA, tmBhi, tmBlo, prm);
    }
    c->gemm_flops += 2 * M * N * K;
    return true;
}

// Z (M x N, ldz) = A^T Y with A stored K x M (row-major, lda) and Y stored K x N (row-major, ldy): the
// reduction runs over the ROWS of both (MN-major UMMA operands), split-K over the rows, partials
// reduced in a fixed order.  f32, 3xTF32 on tcgen05.  Returns false when unsupported.
bool gemm_tf32x3_f32_tn(rc_ctx* c, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                        const float* Y, int64_t ldy, float* Z, int64_t ldz) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 3)) return false;
    if (M > (1LL << 30) || K > (1LL << 30)) return false;
    int64_t nchunks = (N + 127) / 128;
    int64_t per = ((N + nchunks - 1) / nchunks + 31) / 32 * 32;
    const int64_t kblocks = (K + BK - 1) / BK;
    for (int64_t n0 = 0; n0 < N; n0 += per) {
        const int ncols = (int)std::min<int64_t>(per, N - n0);
        const int npad = (ncols + 31) / 32 * 32;
        // B operand K-major: Y^T split into hi / lo once (Y is the small m x l matrix)
        const int64_t ldt = (K + 3) / 4 * 4;
        DevBuf<float> hi(c, (size_t)npad * ldt), lo(c, (size_t)npad * ldt);
        {
            int64_t total = (int64_t)npad * K;
            int nb = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
            split_transpose_kernel<<<nb, 256, 0, c->stream>>>(Y + n0, ldy, (int)K, ncols, npad, hi.p, lo.p, ldt);
            RC_CHECK_LAUNCH(c);
        }
        CUtensorMap tmA = make_map_f32(A, K, M, lda, 32);           // boxes [32 k rows][32 cols]
        CUtensorMap tmBhi = make_map_f32(hi.p, npad, K, ldt, npad);
        CUtensorMap tmBlo = make_map_f32(lo.p, npad, K, ldt, npad);
        Tf32Params prm;
        prm.M = (int)M; prm.N = ncols; prm.K = (int)K; prm.npad = npad;
        prm.m_tiles = (int)((M + BM - 1) / BM);
        uint32_t cols = 32;
        while (cols < (uint32_t)(2 * npad)) cols <<= 1;
        prm.tmem_cols = cols;
        // split-K: about 2 work items per SM, each at least KC k-blocks
        int64_t want = std::max<int64_t>(1, (2LL * c->sm_count + prm.m_tiles - 1) / prm.m_tiles);
        int64_t maxs = std::max<int64_t>(1, kblocks / KC);
        int splits = (int)std::min(want, maxs);
        int64_t kbps = ((kblocks + splits - 1) / splits + KC - 1) / KC * KC;
        splits = (int)((kblocks + kbps - 1) / kbps);
        prm.splits = splits; prm.kb_per_split = (int)kbps;
        DevBuf<float> part;
        if (splits == 1) { prm.y = Z + n0; prm.ldy = ldz; prm.part_stride = 0; }
        else {
            part.alloc(c, (size_t)splits * M * npad);
            prm.y = part.p; prm.ldy = npad; prm.part_stride = M * (int64_t)npad;
        }
        dispatch_tf32<true>(c, npad, tmA, tmBhi, tmBlo, prm);
        if (splits > 1) {
            int64_t n = M * (int64_t)ncols;
            int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
            tf32_reduce_kernel<<<nb, 256, 0, c->stream>>>(M, ncols, splits, part.p, npad, prm.part_stride, Z + n0, ldz);
            RC_CHECK_LAUNCH(c);
        }
    }
    c->gemm_flops += 2 * M * N * K;
    return true;
}

// ---- c32 through the same tcgen05 kernels via the exact real expansion (see gemm_dmma_c64):
//   NN:  (A viewed as real m x 2n) * X' = (A X viewed as real m x 2l),  X' rows 2k / 2k+1 = X_k / i X_k
//   TN:  W = (A viewed as real)^T * (X viewed as real) is 2n x 2l;  A^H X is recombined from W.
namespace {
__global__ void expand_rhs_c32_kernel(float* dst, int64_t ldd, const c32* x, int64_t ldx, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = e / cols, j = e - k * cols;
        c32 v = x[k * ldx + j];
        float* r0 = dst + (2 * k) * ldd + 2 * j;
        float* r1 = dst + (2 * k + 1) * ldd + 2 * j;
        r0[0] = v.re; r0[1] = v.im;
        r1[0] = -v.im; r1[1] = v.re;
    }
}
__global__ void combine_conj_c32_kernel(int64_t n, int64_t l, const float* __restrict__ w, int64_t ldw, c32* __restrict__ z, int64_t ldz) {
    int64_t tot = n * l;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = e / l, c = e - j * l;
        const float* r0 = w + (2 * j) * ldw + 2 * c;
        const float* r1 = w + (2 * j + 1) * ldw + 2 * c;
        z[j * ldz + c] = c32(r0[0] + r1[1], r0[1] - r1[0]);
    }
}
}  // namespace

bool gemm_tf32x3_c32(rc_ctx* c, bool a_conj_transposed, int64_t M, int64_t N, int64_t K, const c32* A, int64_t lda,
                     const c32* B, int64_t ldb, c32* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 1)) return false;
    if (!a_conj_transposed) {
        int64_t ldx = 2 * N;
        DevBuf<float> xp(c, (size_t)(2 * K) * ldx);
        int nb = (int)std::min<int64_t>((K * N + 255) / 256, 148 * 8);
        expand_rhs_c32_kernel<<<nb, 256, 0, c->stream>>>(xp.p, ldx, B, ldb, K, N);
        RC_CHECK_LAUNCH(c);
        return gemm_tf32x3_f32(c, M, 2 * N, 2 * K, reinterpret_cast<const float*>(A), 2 * lda, xp.p, ldx,
                               reinterpret_cast<float*>(C), 2 * ldc);
    }
    int64_t ldw = 2 * N;
    DevBuf<float> w(c, (size_t)(2 * M) * ldw);
    if (!gemm_tf32x3_f32_tn(c, 2 * M, 2 * N, K, reinterpret_cast<const float*>(A), 2 * lda,
                            reinterpret_cast<const float*>(B), 2 * ldb, w.p, ldw)) return false;
    int nb = (int)std::min<int64_t>((M * N + 255) / 256, 148 * 8);
    combine_conj_c32_kernel<<<nb, 256, 0, c->stream>>>(M, N, w.p, ldw, C, ldc);
    RC_CHECK_LAUNCH(c);
    return true;
}
