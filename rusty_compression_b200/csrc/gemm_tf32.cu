// f32 contraction Y = A X on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// 3-product split  A X ~= A_hi X_hi + A_lo X_hi + A_hi X_lo  (hi = value rounded to TF32, lo = the
// exact f32 remainder), FP32 accumulation in TMEM.  This is what carries f32 parity (1e-4 in
// north_star; measured ~1e-6) on the tensor pipe instead of the FP32 SIMT pipe.
//
// Reference role: MatMat::matmat for f32 operators (src/types.rs:58-71), the dominant cost of
// configs 3 and 4.
//
// Structure (one persistent CTA of 15 warps per SM; work item = 128 rows of A x one chunk of <= 96 columns):
//   warp 0        TMA producer of the raw f32 A tiles [128 rows x 32 k], SWIZZLE_128B, 7-8-deep ring
//   warp 10       TMA producer of the pre-split X^T tiles (hi, lo) [N x 32 k], own 4-deep ring (L2-resident)
//   warps 2-5,    splitters, two groups taking alternate k-blocks: raw A tile (shared memory) -> A_hi, A_lo
//   warps 11-14   written straight into TENSOR MEMORY (tcgen05.st, one thread per tile row); the MMAs take A
//                 from TMEM, so the split tiles never go back through shared memory.  Splitting A on the fly
//                 avoids a pre-split copy that would double the HBM traffic of the pass.
//   warp 1        MMA issuer: the warp runs converged and one elected lane issues 12 tcgen05.mma per k-block
//                 (3 products x 4 K-steps of 8; A from TMEM, B from shared memory); tcgen05.commit frees the
//                 TMEM stage and the X slot; accumulator double-buffered in TMEM
//   warps 6-9     epilogue: tcgen05.ld (32 lanes x 16 columns) -> FP32 promotion in registers -> global Y
// X is tiny (n x l): it is transposed and split once by a prologue kernel so that both B operands
// are K-major TMA tiles.
//
// PREC = 1 (context option "f32_precision" = 1, opt-in): ONE product in bf16 -- the splitter rounds the raw f32 tile to
// bf16 pairs (cvt.rn.bf16x2.f32) on its way into tensor memory, X^T is pre-rounded to bf16 tiles (64-byte rows,
// SWIZZLE_64B), the MMAs are tcgen05.mma kind::f16 (two K = 16 steps per k-block instead of twelve kind::tf32 ones),
// FP32 accumulation and promotion as before.  Same rings, same pass over A: it pays where the TF32 split is
// tensor-bound (l > ~100); relative accuracy ~3e-3 instead of ~1e-6, which is why it is never the default.
//
// TRANS mode (Z = A^T Y, reduction over the rows of A, split-K with a fixed-order reduction): the
// raw tile arrives as four [32 k][32 i] boxes and the splitter transposes it through registers while
// splitting, so the MMA still sees a K-major A (an MN-major descriptor variant produced all-zero
// products on this toolchain and was dropped).
#include <cuda.h>
#include <cstdlib>
#include "rc_internal.cuh"
#include "splitk_reduce.cuh"

namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1)
constexpr int BK = 32;            // floats per stage = one 128-byte swizzle row
constexpr int NTHREADS = 480;     // 15 warps: A-TMA, MMA, 4 splitters, 4 epilogue, B-TMA, 4 more splitters
constexpr int MAX_RS = 8, MAX_MS = 6, MAX_BS = 6;   // ring depths: raw A tiles (smem), split A (TMEM), B tiles (smem)
constexpr int MAX_CHUNK = 96;     // columns per work item: the promoted FP32 accumulator row lives in one epilogue thread's registers
constexpr int A_TMEM_COLS = 64;   // one split stage in TMEM: A_hi (32 columns = 32 k) + A_lo (32 columns)

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

// Row-major [rows][cols] bf16 matrix (the pre-rounded X^T), box = 32 cols (64 B) x box_rows, SWIZZLE_64B.
CUtensorMap make_map_bf16(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    RC_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map box rows out of range");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RC_THROW(RC_CUDA_ERROR, "cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
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
// A operand from tensor memory (128 lanes = tile rows, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: D = F32, A = B = BF16 (both K-major), M = 128, N = n   (kind::f16)
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t n) {
    uint32_t d = 0;
    d |= 1u << 4;            // c_format = F32
    d |= 1u << 7;            // a_format = BF16
    d |= 1u << 10;           // b_format = BF16
    d |= (n >> 3) << 17;     // n_dim
    d |= (uint32_t)(BM >> 4) << 24;   // m_dim
    return d;
}
// kind::f16, A from tensor memory (two bf16 values of consecutive k per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive TMEM columns of this thread's lane <- 16 registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 32 consecutive TMEM columns of this thread's lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// One lane of a converged warp (the CUTLASS elect_one_sync idiom): lets the compiler issue the
// uniform-datapath tcgen05 instructions without a per-active-thread serialisation loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n" : "+r"(pred));
    return pred != 0;
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
    // N > 128: the columns are processed as `nchunks` chunks of npad columns each (the promoted FP32
    // accumulator of a row has to fit the registers of one epilogue thread).  The chunk index is the
    // FASTEST-varying part of the work-item order, so the CTAs that work on the same rows of A run
    // side by side and A is fetched from HBM once and served to the other chunks out of L2.
    int nchunks;
    int vec_store;        // y and ldy are 16-byte aligned: the epilogue may use float4 stores
    int agroup;           // raw A tiles requested per group (<= RS / 2)
    uint32_t zero;        // always 0 at run time (opaque to ptxas): ties the raw-slot release to the loaded data
};

// KC k-blocks (KC * 32 values of K) are accumulated inside the tensor core before the partial sum is
// promoted to an FP32 register accumulator (round-to-nearest adds on the CUDA cores): the tensor
// core's own FP32 accumulation truncates, which costs ~K * 2^-24 relative accuracy on long chains
// (measured 2.9e-5 at K = 4096 without promotion).
constexpr int KC = 8;

// Three rings decouple the HBM latency from the tensor pipe:
//   raw  (RS deep, 16 KB each)  TMA landing zone for the raw f32 A tiles -- deep, because at l <= 100 the
//                               pass is HBM-bound and ~64 KB per SM must be in flight to cover the latency
//   ahl  (MS deep, 64 TMEM cols) A_hi / A_lo in tensor memory (lane = tile row, column = k), written by the splitter
//   bt   (BS deep)              X_hi^T / X_lo^T tiles (L2-resident), own TMA producer warp
template <int RS, int MS, int BS, int NPADC, bool TRANS, int PREC>
__global__ void __launch_bounds__(NTHREADS, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, Tf32Params prm) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long bars[2 * MAX_RS + 2 * MAX_MS + MAX_BS + 4];
    __shared__ uint32_t tmem_base_smem;
    const uint32_t rawfull0 = smem_u32(&bars[0]);                          // raw A tile landed
    const uint32_t rawempty0 = smem_u32(&bars[MAX_RS]);                    // raw A tile consumed by the splitter
    const uint32_t split0 = smem_u32(&bars[2 * MAX_RS]);                   // A_hi / A_lo written
    const uint32_t aempty0 = smem_u32(&bars[2 * MAX_RS + MAX_MS]);         // MMAs that read A_hi / A_lo are done
    const uint32_t bfull0 = smem_u32(&bars[2 * MAX_RS + 2 * MAX_MS]);      // B tiles landed
    const uint32_t accf0 = smem_u32(&bars[2 * MAX_RS + 2 * MAX_MS + MAX_BS]);          // accumulator buffer full (2)
    const uint32_t acce0 = accf0 + 16;                                                 // accumulator buffer drained (2)

    constexpr int npad = NPADC;
    // PREC: 0 = 3xTF32 with A_hi and A_lo in tensor memory; 1 = bf16 single product; 2 = 3xTF32 with the HIGH part taken
    // straight from the raw tile in shared memory (SS-form MMAs: the tensor core reads only the top 19 bits of a TF32
    // operand, i.e. hi = a truncated, so only lo = a - trunc(a) has to be produced and stored to tensor memory -- half
    // the splitter's tcgen05.st traffic and half the TMEM per stage; NN mode only)
    constexpr bool TF32 = (PREC != 1);
    constexpr bool HS = (PREC == 2);
    static_assert(!(HS && TRANS), "the shared-memory high part needs the K-major raw tile of the NN mode");
    constexpr int ACOLS = HS ? 32 : A_TMEM_COLS;                 // TMEM columns of one split stage
    constexpr uint32_t A_BYTES = BM * BK * 4;                    // 16 KB
    constexpr uint32_t B_BYTES = (uint32_t)NPADC * BK * (TF32 ? 4 : 2);      // one X^T tile: f32 hi (and lo) / bf16
    constexpr uint32_t B_STAGE = TF32 ? 2 * B_BYTES : B_BYTES;
    const uint32_t raw_base = smem_base;
    const uint32_t bt_base = raw_base + RS * A_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < RS; ++s) { mbar_init(rawfull0 + 8 * s, 1); mbar_init(rawempty0 + 8 * s, HS ? 5 : 4); }   // HS: + the MMAs' commit
        for (int s = 0; s < MS; ++s) { mbar_init(split0 + 8 * s, 4); mbar_init(aempty0 + 8 * s, 1); }
        for (int s = 0; s < BS; ++s) mbar_init(bfull0 + 8 * s, 1);
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
    const int total_items = prm.m_tiles * prm.splits * prm.nchunks;      // NN: splits == 1
    // work item -> (column chunk, tile along M, k-block range)
#define RC_ITEM(t)                                                                         \
    const int ch_ = (t) % prm.nchunks, rest_ = (t) / prm.nchunks;                          \
    const int mt_ = rest_ % prm.m_tiles, sp_ = rest_ / prm.m_tiles;                        \
    const int kb0_ = sp_ * prm.kb_per_split;                                               \
    const int kb1_ = min(kblocks_all, kb0_ + prm.kb_per_split);                            \
    const int kblocks = kb1_ - kb0_;                                                       \
    (void)mt_; (void)sp_; (void)kblocks; (void)ch_;

    if (warp == 0) {
        // ===================== TMA producer: raw A tiles =====================
        // Tiles can be requested in groups of `agroup` consecutive k-blocks (agroup x 128 B contiguous per row
        // of A, issued back to back) -- an experiment on DRAM page locality that did not pay off on B200.
        if (lane == 0) {
            int rs = 0; uint32_t rph = 0;
            const int G = prm.agroup;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                RC_ITEM(t)
                const int m0 = mt_ * BM;
                for (int kbg = kb0_; kbg < kb1_; kbg += G) {
                    const int ng = min(G, kb1_ - kbg);
                    {   // all slots of the group must be free before the first request goes out
                        int s2 = rs; uint32_t p2 = rph;
                        for (int g = 0; g < ng; ++g) {
                            mbar_wait(rawempty0 + 8 * s2, p2 ^ 1u);
                            if (++s2 == RS) { s2 = 0; p2 ^= 1u; }
                        }
                    }
                    for (int g = 0; g < ng; ++g) {
                        const int kb = kbg + g;
                        const uint32_t sa = raw_base + rs * A_BYTES;
                        const uint32_t fb = rawfull0 + 8 * rs;
                        mbar_expect_tx(fb, A_BYTES);
                        if (!TRANS) {
                            tma_load_2d(sa, &tmA, kb * BK, m0, fb);                                   // [128 rows][32 k]
                        } else {
#pragma unroll
                            for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &tmA, m0 + 32 * b, kb * BK, fb);   // raw [32 k][32 i] x 4
                        }
                        if (++rs == RS) { rs = 0; rph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 10) {
        // ===================== TMA producer: X_hi^T / X_lo^T tiles =====================
        if (lane == 0) {
            int bs = 0; uint32_t bph = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                RC_ITEM(t)
                for (int kb = kb0_; kb < kb1_; ++kb) {
                    mbar_wait(aempty0 + 8 * bs, bph ^ 1u);           // MS == BS: one commit frees the A (TMEM) and B slot
                    const uint32_t sb = bt_base + bs * B_STAGE;
                    const uint32_t fb = bfull0 + 8 * bs;
                    mbar_expect_tx(fb, B_STAGE);
                    tma_load_2d(sb, &tmBhi, kb * BK, ch_ * npad, fb);                             // [npad rows][32 k]
                    if (TF32) tma_load_2d(sb + B_BYTES, &tmBlo, kb * BK, ch_ * npad, fb);
                    if (++bs == BS) { bs = 0; bph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the control flow (converged, so the tcgen05 instructions of the elected lane
        // compile to plain uniform-datapath issues); the single-thread form of this loop spent ~100 cycles
        // per MMA on elect/branch scaffolding and capped the kernel at 12 MMAs per 1380 cycles.
        {
            const uint32_t idesc = TF32 ? umma_idesc_tf32((uint32_t)npad) : umma_idesc_bf16((uint32_t)npad);
            // UMMA shared-memory descriptor of the X^T tiles: constant high word, low word = address >> 4
            // (f32 tiles: 128-byte rows, SWIZZLE_128B, 8-row atoms of 1024 B; bf16 tiles: 64-byte rows, SWIZZLE_64B, 512 B)
            const uint64_t desc_hi = TF32 ? (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32
                                               : (uint64_t)((512u >> 4) | (1u << 14) | (4u << 29)) << 32;
            const uint32_t bt_lo = (bt_base & 0x3FFFFu) >> 4;
            const uint32_t raw_lo = (raw_base & 0x3FFFFu) >> 4;
            int rs = 0;                                              // HS: raw slot of the current k-block
            int ms = 0; uint32_t mph = 0;
            int bs = 0; uint32_t bph = 0;
            int it = 0;                                              // counts K-chunks (TMEM buffer hand-offs)
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                RC_ITEM(t)
                for (int kc0 = 0; kc0 < kblocks; kc0 += KC, ++it) {
                    const int buf = it & 1;
                    const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
                    mbar_wait(acce0 + 8 * buf, acc_phase ^ 1u);      // epilogue drained this buffer
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * npad);
                    const int kc1 = min(kblocks, kc0 + KC);
                    for (int kb = kc0; kb < kc1; ++kb) {
                        mbar_wait(bfull0 + 8 * bs, bph);
                        mbar_wait(split0 + 8 * ms, mph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            const uint32_t a_hi = tmem_base + (uint32_t)(2 * npad + ms * ACOLS), a_lo = HS ? a_hi : a_hi + 32u;
                            const uint32_t bh = bt_lo + (uint32_t)bs * (B_STAGE >> 4), bl = bh + (B_BYTES >> 4);
                            if (PREC == 1) {
                                // one product, bf16: K = 16 per MMA = 32 bytes along the 64-byte row = 8 TMEM columns of A
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_bf16_ts(d_tmem, a_hi + 8u * k, desc_hi | (uint64_t)(bh + 2u * k), idesc, (kb != kc0 || k != 0) ? 1u : 0u);
                            } else if (HS) {
                                // raw tile [128 rows][32 k] f32, SWIZZLE_128B: the same K-major layout as the X^T tiles
                                const uint32_t ar = raw_lo + (uint32_t)rs * (A_BYTES >> 4);
#pragma unroll
                                for (int k = 0; k < BK / 8; ++k) {
                                    const uint64_t dbh = desc_hi | (uint64_t)(bh + 2u * k), dbl = desc_hi | (uint64_t)(bl + 2u * k);
                                    const uint64_t da = desc_hi | (uint64_t)(ar + 2u * k);
                                    umma_tf32_ts(d_tmem, a_lo + 8u * k, dbh, idesc, (kb != kc0 || k != 0) ? 1u : 0u);
                                    umma_tf32(d_tmem, da, dbl, idesc, 1u);
                                    umma_tf32(d_tmem, da, dbh, idesc, 1u);
                                }
                                umma_commit(rawempty0 + 8 * rs);             // the raw slot is free when the SS MMAs have read it
                            } else
#pragma unroll
                            for (int k = 0; k < BK / 8; ++k) {
                                // B K-major: 8 floats (32 B) along K inside the swizzle row; A: 8 TMEM columns per K-step
                                const uint64_t dbh = desc_hi | (uint64_t)(bh + 2u * k), dbl = desc_hi | (uint64_t)(bl + 2u * k);
                                // small terms first, then the dominant one
                                umma_tf32_ts(d_tmem, a_lo + 8u * k, dbh, idesc, (kb != kc0 || k != 0) ? 1u : 0u);
                                umma_tf32_ts(d_tmem, a_hi + 8u * k, dbl, idesc, 1u);
                                umma_tf32_ts(d_tmem, a_hi + 8u * k, dbh, idesc, 1u);
                            }
                            umma_commit(aempty0 + 8 * ms);               // frees the TMEM stage and the B slot when the MMAs retire
                            if (kb == kc1 - 1) umma_commit(accf0 + 8 * buf);
                        }
                        __syncwarp();
                        if (++rs == RS) rs = 0;
                        if (++ms == MS) { ms = 0; mph ^= 1u; }
                        if (++bs == BS) { bs = 0; bph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp < 6 || warp >= 11) {
        // ===================== splitter: raw A tile (smem) -> A_hi, A_lo (TMEM) =====================
        // Two groups of four warps (warps 2-5 and 11-14) take alternate k-blocks: one block is a serial chain
        // (wait, LDS, convert, wait, tcgen05.st, wait::st, arrive) of ~500 cycles, which a single group
        // cannot hide.  thread <-> tile row: a warp may only touch the TMEM lane quadrant (warp % 4).
        const int grp = (warp >= 11) ? 1 : 0;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int rs = grp; uint32_t rph = 0;            // ring positions of this group's first block (RS, MS >= 2)
        int ms = grp; uint32_t mph = 0;
        int seq = 0;                               // k-block sequence number over all work items of this CTA
        const unsigned char* const smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
        for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
            RC_ITEM(t)
            for (int kb = 0; kb < kblocks; ++kb, ++seq) {
                if ((seq & 1) != grp) continue;
                mbar_wait(rawfull0 + 8 * rs, rph);
                const unsigned char* rawp = smem_al + (size_t)rs * A_BYTES;
                float vals[32];
                if (!TRANS) {
                    // row `row` of the [128 rows][32 k] tile: 16-byte chunk j sits at chunk position j ^ (row & 7)
                    // (SWIZZLE_128B), so the eight LDS.128 of a quarter-warp hit all 32 banks
                    const unsigned char* rp = rawp + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(rp + ((j ^ (row & 7)) << 4));
                        vals[4 * j + 0] = v.x; vals[4 * j + 1] = v.y; vals[4 * j + 2] = v.z; vals[4 * j + 3] = v.w;
                    }
                } else {
                    // the raw tile is 4 boxes of [32 k][32 i] (the contraction index is the row index of A);
                    // this thread owns tile row i = row: 32 conflict-free LDS.32 (a warp reads one 128-byte row per k)
                    const unsigned char* rawb = rawp + quad * 4096;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        vals[k] = *reinterpret_cast<const float*>(rawb + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
                }
                uint32_t hi[32], lo[32];
                if (HS) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const uint32_t u = __float_as_uint(vals[k]) & 0xFFFFE000u;      // what the tensor core reads of a
                        hi[k] = u;
                        lo[k] = __float_as_uint(vals[k] - __uint_as_float(u));
                    }
                } else if (PREC == 0) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vals[k]));
                        hi[k] = u;
                        lo[k] = __float_as_uint(vals[k] - __uint_as_float(u));
                    }
                } else {
                    // bf16 pairs: column j of the TMEM stage holds k = 2j (low half) and k = 2j + 1 (high half)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t u;
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(vals[2 * j + 1]), "f"(vals[2 * j]));
                        hi[j] = u;
                    }
                    lo[31] = 0u;
                }
                mbar_wait(aempty0 + 8 * ms, mph ^ 1u);                   // MMAs that read this TMEM stage are done
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * npad + ms * ACOLS);
                if (HS) {
                    tmem_st32(ta, lo);
                } else if (PREC == 0) {
                    tmem_st32(ta, hi);
                    tmem_st32(ta + 32u, lo);
                } else {
                    uint32_t h16[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) h16[j] = hi[j];
                    tmem_st16(ta, h16);
                }
                // The raw slot is released only BEHIND the tcgen05.st: they cannot issue before the LDS results
                // are in registers.  (An arrive placed right after the loads issues behind the LDS *issue*, and
                // when the load/store pipe is backed up by the epilogue's store burst the TMA refill of the slot
                // overtook the loads: whole lane-quadrants of wrong rows on CTAs with more than one work item.)
                // (and the barrier address carries a data dependence on the loaded values, so no scheduler may
                // hoist the arrive above the loads' completion whatever it does with the tcgen05.st)
                __syncwarp();
                if (lane == 0) mbar_arrive(rawempty0 + 8 * rs + ((hi[0] ^ lo[31]) & prm.zero));
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(split0 + 8 * ms);
                rs += 2; if (rs >= RS) { rs -= RS; rph ^= 1u; }
                ms += 2; if (ms >= MS) { ms -= MS; mph ^= 1u; }
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
                float* yrow = prm.y + (int64_t)sp_ * prm.part_stride + (int64_t)row * prm.ldy + ch_ * npad;
                const int nvalid = prm.N - ch_ * npad;
                if (prm.vec_store) {          // 16-byte stores: a quarter of the LSU transactions of scalar stores
#pragma unroll
                    for (int j = 0; j < NPADC; j += 4) {
                        if (j + 3 < nvalid) *reinterpret_cast<float4*>(yrow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (j + q < nvalid) yrow[j + q] = acc[j + q];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NPADC; ++j)
                        if (j < nvalid) yrow[j] = acc[j];
                }
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

// X (K x N row-major) -> XhiT, XloT (nrows x K row-major, K contiguous), rows >= N zero.  32 x 32 tiles
// through shared memory: reads run along the rows of X, writes along K (both coalesced); for the Gram
// matrices of the Cholesky-QR path X is the m x l sketch itself, so this is a bandwidth kernel.
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ x, int64_t ldx, int K, int N, int nrows,
                       float* __restrict__ hiT, float* __restrict__ loT, int64_t ldt) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j0 = blockIdx.y * 32;
    for (int64_t kt = blockIdx.x; kt * 32 < K; kt += gridDim.x) {
        const int64_t k0 = kt * 32;
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t k = k0 + r; const int j = j0 + tx;
            tile[r][tx] = (k < K && j < N) ? x[k * ldx + j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int j = j0 + r; const int64_t k = k0 + tx;
            if (k < K && j < nrows) {
                const float v = tile[tx][r];
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
                const float h = __uint_as_float(u);
                hiT[(int64_t)j * ldt + k] = h;
                loT[(int64_t)j * ldt + k] = v - h;
            }
        }
        __syncthreads();
    }
}
// X (K x N row-major) -> X^T rounded to bf16 (nrows x K row-major, K contiguous), rows >= N zero.
__global__ void __launch_bounds__(256)
bf16_transpose_kernel(const float* __restrict__ x, int64_t ldx, int K, int N, int nrows, unsigned short* __restrict__ xT, int64_t ldt) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j0 = blockIdx.y * 32;
    for (int64_t kt = blockIdx.x; kt * 32 < K; kt += gridDim.x) {
        const int64_t k0 = kt * 32;
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t k = k0 + r; const int j = j0 + tx;
            tile[r][tx] = (k < K && j < N) ? x[k * ldx + j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int j = j0 + r; const int64_t k = k0 + tx;
            if (k < K && j < nrows) {
                unsigned short h;
                asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(tile[tx][r]));
                xT[(int64_t)j * ldt + k] = h;
            }
        }
        __syncthreads();
    }
}
void launch_bf16_transpose(rc_ctx* c, const float* x, int64_t ldx, int64_t K, int N, int nrows, unsigned short* xT, int64_t ldt) {
    dim3 block(32, 8);
    dim3 grid((unsigned)std::min<int64_t>((K + 31) / 32, 148 * 64), (unsigned)((nrows + 31) / 32));
    bf16_transpose_kernel<<<grid, block, 0, c->stream>>>(x, ldx, (int)K, N, nrows, xT, ldt);
    RC_CHECK_LAUNCH(c);
}
void launch_split_transpose(rc_ctx* c, const float* x, int64_t ldx, int64_t K, int N, int nrows, float* hiT, float* loT, int64_t ldt) {
    dim3 block(32, 8);
    dim3 grid((unsigned)std::min<int64_t>((K + 31) / 32, 148 * 64), (unsigned)((nrows + 31) / 32));
    split_transpose_kernel<<<grid, block, 0, c->stream>>>(x, ldx, (int)K, N, nrows, hiT, loT, ldt);
    RC_CHECK_LAUNCH(c);
}

template <int RS, int MS, int BS, int NPADC, bool TRANS, int PREC>
void launch_tf32(rc_ctx* c, const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
                 Tf32Params prm) {
    static_assert(RS <= MAX_RS && MS <= MAX_MS && BS <= MAX_BS && MS == BS, "ring depth (A-TMEM and B rings share their release barrier)");
    constexpr int ACOLS = PREC == 2 ? 32 : A_TMEM_COLS;
    static_assert(2 * NPADC + MS * ACOLS <= 512, "tensor memory columns");
    constexpr size_t smem = (size_t)RS * BM * BK * 4 + (size_t)BS * (PREC != 1 ? 2 * NPADC * BK * 4 : NPADC * BK * 2) + 1024;
    static_assert(smem <= 226 * 1024, "ring configuration exceeds shared memory");
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * NPADC + MS * ACOLS)) cols <<= 1;
    prm.tmem_cols = cols;
    prm.agroup = 1;     // measured: 1 is best on B200 (2, 4, 8 are 0-4 % slower at 32768^2 x 64)
    prm.zero = 0;
    prm.vec_store = ((reinterpret_cast<uintptr_t>(prm.y) & 15) == 0 && (prm.ldy & 3) == 0 && (prm.part_stride & 3) == 0) ? 1 : 0;
    RC_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<RS, MS, BS, NPADC, TRANS, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = std::min(prm.m_tiles * prm.splits * prm.nchunks, c->sm_count);
    tf32x3_gemm_kernel<RS, MS, BS, NPADC, TRANS, PREC><<<grid, NTHREADS, smem, c->stream>>>(tmA, tmBhi, tmBlo, prm);
    RC_CHECK_LAUNCH(c);
}
template <bool TRANS>
void dispatch_tf32(rc_ctx* c, int npad, const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo,
                   const Tf32Params& prm, int prec) {
    if (prec == 1) {
        switch (npad) {  // bf16: the X^T ring is a quarter of the size, the raw ring stays as deep as it can be
            case 32: launch_tf32<8, 4, 4, 32, TRANS, 1>(c, tmA, tmBhi, tmBlo, prm); break;
            case 64: launch_tf32<8, 4, 4, 64, TRANS, 1>(c, tmA, tmBhi, tmBlo, prm); break;
            default: launch_tf32<8, 4, 4, 96, TRANS, 1>(c, tmA, tmBhi, tmBlo, prm); break;
        }
        return;
    }
    if constexpr (!TRANS) {
        if (c->tf32_ring == 2) {
            // high part from shared memory (PREC = 2), six split stages of 32 TMEM columns
            switch (npad) {
                case 32: launch_tf32<8, 6, 6, 32, TRANS, 2>(c, tmA, tmBhi, tmBlo, prm); break;
                case 64: launch_tf32<7, 6, 6, 64, TRANS, 2>(c, tmA, tmBhi, tmBlo, prm); break;
                default: launch_tf32<6, 5, 5, 96, TRANS, 2>(c, tmA, tmBhi, tmBlo, prm); break;
            }
            return;
        }
    }
    if (c->tf32_ring >= 1) {
        // deeper split ring (option "tf32_ring" = 1): the splitters spend ~30 % of their samples waiting for the MMAs of
        // the stage four k-blocks back to retire (ncu, round 1) -- six TMEM stages of A_hi / A_lo next to the two accumulators
        switch (npad) {
            case 32: launch_tf32<8, 6, 6, 32, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     // 128 + 48 KB, 64 + 384 TMEM columns
            case 64: launch_tf32<7, 6, 6, 64, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     // 112 + 96 KB, 128 + 384
            default: launch_tf32<6, 5, 5, 96, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     //  96 + 120 KB, 192 + 320
        }
        return;
    }
    switch (npad) {      //      raw (smem)  A hi/lo (TMEM)  X^T (smem)
        case 32: launch_tf32<8, 4, 4, 32, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     // 128 + 32 KB
        case 64: launch_tf32<8, 4, 4, 64, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     // 128 + 64 KB
        default: launch_tf32<7, 4, 4, 96, TRANS, 0>(c, tmA, tmBhi, tmBlo, prm); break;     // 112 + 96 KB
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
    // columns are processed in chunks of at most MAX_CHUNK (the promoted FP32 accumulator of a row lives in
    // the registers of one epilogue thread); all chunks run in ONE launch, chunk-fastest (Tf32Params)
    const int nchunks = (int)((N + MAX_CHUNK - 1) / MAX_CHUNK);
    const int npad = (int)(((N + nchunks - 1) / nchunks + 31) / 32 * 32);          // 32, 64 or 96
    const int prec = c->f32_precision == 1 ? 1 : 0;
    const int64_t ldt = (K + 7) / 8 * 8;
    const int nrows = nchunks * npad;                   // stacked X^T: column j of X is row j
    DevBuf<float> hiT, loT;
    DevBuf<unsigned short> bfT;
    CUtensorMap tmA = make_map_f32(A, M, K, lda, BM);
    CUtensorMap tmBhi, tmBlo;
    if (prec == 1) {
        bfT.alloc(c, (size_t)nrows * ldt);
        launch_bf16_transpose(c, X, ldx, K, (int)N, nrows, bfT.p, ldt);
        tmBhi = make_map_bf16(bfT.p, nrows, K, ldt, npad);
        tmBlo = tmBhi;
    } else {
        // only the N real rows exist: the TMA zero-fills the out-of-bounds rows of the last chunk's box
        hiT.alloc(c, (size_t)N * ldt); loT.alloc(c, (size_t)N * ldt);
        launch_split_transpose(c, X, ldx, K, (int)N, (int)N, hiT.p, loT.p, ldt);
        tmBhi = make_map_f32(hiT.p, N, K, ldt, npad);
        tmBlo = make_map_f32(loT.p, N, K, ldt, npad);
    }
    Tf32Params prm;
    prm.y = Y; prm.ldy = ldy; prm.M = (int)M; prm.N = (int)N; prm.K = (int)K; prm.npad = npad;
    prm.m_tiles = (int)((M + BM - 1) / BM);
    prm.nchunks = nchunks;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * npad)) cols <<= 1;
    prm.tmem_cols = cols;
    prm.splits = 1; prm.kb_per_split = (int)((K + BK - 1) / BK); prm.part_stride = 0;
    dispatch_tf32<false>(c, npad, tmA, tmBhi, tmBlo, prm, prec);
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
    const int nchunks = (int)((N + MAX_CHUNK - 1) / MAX_CHUNK);
    const int npad = (int)(((N + nchunks - 1) / nchunks + 31) / 32 * 32);
    const int nrows = nchunks * npad;
    const int64_t kblocks = (K + BK - 1) / BK;
    // B operand K-major: Y^T split into hi / lo once (Y is the small m x l matrix), chunks stacked
    const int prec = c->f32_precision == 1 ? 1 : 0;
    const int64_t ldt = (K + 7) / 8 * 8;
    DevBuf<float> hi, lo;
    DevBuf<unsigned short> bfT;
    CUtensorMap tmA = make_map_f32(A, K, M, lda, 32);           // boxes [32 k rows][32 cols]
    CUtensorMap tmBhi, tmBlo;
    if (prec == 1) {
        bfT.alloc(c, (size_t)nrows * ldt);
        launch_bf16_transpose(c, Y, ldy, K, (int)N, nrows, bfT.p, ldt);
        tmBhi = make_map_bf16(bfT.p, nrows, K, ldt, npad);
        tmBlo = tmBhi;
    } else {
        hi.alloc(c, (size_t)N * ldt); lo.alloc(c, (size_t)N * ldt);
        launch_split_transpose(c, Y, ldy, K, (int)N, (int)N, hi.p, lo.p, ldt);
        tmBhi = make_map_f32(hi.p, N, K, ldt, npad);
        tmBlo = make_map_f32(lo.p, N, K, ldt, npad);
    }
    Tf32Params prm;
    prm.M = (int)M; prm.N = (int)N; prm.K = (int)K; prm.npad = npad;
    prm.m_tiles = (int)((M + BM - 1) / BM);
    prm.nchunks = nchunks;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * npad)) cols <<= 1;
    prm.tmem_cols = cols;
    // split-K: about 2 work items per SM, each at least KC k-blocks
    const int64_t base_items = (int64_t)prm.m_tiles * nchunks;
    int64_t want = std::max<int64_t>(1, (2LL * c->sm_count + base_items - 1) / base_items);
    int64_t maxs = std::max<int64_t>(1, kblocks / KC);
    int splits = (int)std::min(want, maxs);
    int64_t kbps = ((kblocks + splits - 1) / splits + KC - 1) / KC * KC;
    splits = (int)((kblocks + kbps - 1) / kbps);
    prm.splits = splits; prm.kb_per_split = (int)kbps;
    DevBuf<float> part;
    if (splits == 1) { prm.y = Z; prm.ldy = ldz; prm.part_stride = 0; }
    else {
        part.alloc(c, (size_t)splits * M * nrows);
        prm.y = part.p; prm.ldy = nrows; prm.part_stride = M * (int64_t)nrows;
    }
    dispatch_tf32<true>(c, npad, tmA, tmBhi, tmBlo, prm, prec);
    if (splits > 1) rc_splitk::reduce<float>(c, M, N, splits, part.p, nrows, prm.part_stride, Z, ldz);
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
