// FP64 tensor-pipe GEMM for the dominant contractions of the f64/c64 path:
//   NN:  Y = A X        (A m x n row-major, X n x l)          reference a2, src/types.rs:58-71
//   TN:  Z = A^T X      (reduction over the m rows, split-K)   reference a3, src/types.rs:88-101
// Blackwell has no FP64 kind of tcgen05.mma, so the FP64 path uses the DMMA pipe
// (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) fed by TMA:
//   * one producer warp issues cp.async.bulk.tensor (SWIZZLE_128B boxes of 16 doubles) into a
//     6-deep mbarrier ring; 8 consumer warps (4 x 2) own a 64 x BN tile, BN in {16..96};
//   * fragment addressing is chosen so that every shared-memory read is bank-conflict free
//     under the 128B swizzle: k-slot t of MMA h reads k = 8*kg + 2t + h (A and B agree), and in
//     NN mode MMA row g is tile row 4*(g&1) + (g>>1); every fragment address is one of six per-lane
//     constants plus an immediate (no spills at 96 registers: 2.53 -> 2.43 ms NN, 2.66 -> 2.48 ms TN);
//   * persistent CTAs (2 per SM) walk (split, n-chunk, m-tile) work items; split-K partials are
//     reduced in a fixed order (deterministic).
// c64 reuses the same kernels through an exact real expansion (see gemm_dmma_c64).
#include <cuda.h>
#include <atomic>
#include "rc_internal.cuh"
#include "splitk_reduce.cuh"

namespace {

constexpr int BM = 64;
constexpr int BK = 16;
constexpr int NCW = 8;                       // consumer warps (4 along M x 2 along N)
constexpr int NTHREADS = (NCW + 1) * 32;     // + 1 producer warp
constexpr int MAX_STAGES = 6;
// stages: as many as fit in ~112 KB so two CTAs share an SM
__host__ __device__ constexpr int stages_for(int bn) {
    int sb = BM * BK * 8 + BK * bn * 8;
    int s = (112 * 1024 - 1024) / sb;
    return s > MAX_STAGES ? MAX_STAGES : s;
}

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

// Row-major [rows][cols] f64 matrix with pitch ld (elements); box = 16 cols x box_rows.
CUtensorMap make_map(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RC_THROW(RC_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
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
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
}

struct DmmaParams {
    double* d;            // output (or partial buffer)
    int64_t ldd;
    int64_t part_stride;  // elements between split partials (0: direct output)
    int M, N, K;
    int m_tiles, n_chunks, splits;
    int k_chunk;          // multiple of BK
    int* tile_counter;    // dynamic tile scheduler (zeroed before the launch)
    uint32_t zero;        // always 0 at run time (opaque to ptxas): ties the stage release to the loaded data
};

// ---- DMMA consumer warps -------------------------------------------------------------------------------
// Shared-memory addressing.  A box is [16 k][16 doubles] with 128-byte rows, 16-byte chunk index XOR (k & 7).
// For lane (g, t4), slot h of MMA kg reads k = 8 kg + kk_h with kk_h = 2 t4 + h, and column 8 gj + g of
// 8-column group gj sits at   (gj >> 1) * 2048 + kg * 1024 + (L_h ^ ((gj & 1) << 6)),
//   L_h = kk_h * 128 + (((g >> 1) ^ kk_h) << 4) + ((g & 1) << 3),
// so all the B (and transposed-A) fragment addresses of a warp are four lane constants plus immediates.
//
// TAILW (0, 2 or 4) = number of valid columns in the LAST 8-column group of this warp's tile when the output
// is ragged (N = BN - 8 + TAILW, e.g. l = 74 -> BN = 80, TAILW = 2); 0 for a warp whose groups are all full.
// An m8n8k4 DMMA on that group would spend a full 16 pipe cycles per sub-partition on 2 (4) useful columns;
// the warps that own it (column half 1) instead form those columns with plain DFMAs on the A fragments they
// already hold: lane (g, t4) keeps the partial sum over its own k slots and the four t4 lanes are summed once
// in the epilogue.  2 * TAILW DFMAs (2 cycles each) replace 2 DMMAs (16 cycles each) per (8 rows x 8 k); the
// FP64 pipe runs the two back to back without a switching penalty (tools/micro/mix_fp64.cu: 16 DMMA + 8 DFMA
// = 277 cycles against 321 for 20 DMMA).  The column half of a warp is (warp ^ (warp >> 2)) & 1, so that every
// SM sub-partition (warp & 3) hosts one warp of each kind per CTA: with half = warp & 1 the lighter warps all
// sat on sub-partitions 1 and 3 and the kernel was exactly as fast as without the tail path.
// The two kinds of warp run separate instantiations of this function, so neither carries the other's registers.
template <int BN, bool TRANS_A, int TAILW>
__device__ __forceinline__ void dmma_consumer(const DmmaParams& prm, const uint32_t smem_base, const uint32_t full0,
                                              const uint32_t empty0, const int* tile_ring, const int warp, const int lane) {
    constexpr int WM = BM / 4, WN = BN / 2;          // warp tile
    constexpr int RG = WM / 8, CG = WN / 8;
    constexpr int CGD = (TAILW > 0) ? CG - 1 : CG;   // column groups on the DMMA pipe
    constexpr int TW = (TAILW > 0) ? TAILW : 2;
    constexpr int A_BYTES = BM * BK * 8;
    constexpr int STAGE_BYTES = A_BYTES + BK * BN * 8;
    constexpr int STAGES = stages_for(BN);
    static_assert(RG == 2, "the address tables below assume two row groups per warp");

    const int g = lane >> 2, t4 = lane & 3;
    const int w1 = (warp ^ (warp >> 2)) & 1;         // column half (see above)
    const int wm0 = (warp >> 1) * WM, wn0 = w1 * WN;
    const int rho = 4 * (g & 1) + (g >> 1);          // NN: MMA row g <-> tile row rho (conflict-free LDS.128)
    uint32_t L[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int kk = 2 * t4 + h;
        L[h] = (uint32_t)(kk * 128 + (((g >> 1) ^ kk) << 4) + ((g & 1) << 3));
    }
    // A fragments.  NN: row = wm0 + 8 i + rho, one LDS.128 at row * 128 + (((4 kg + t4) ^ rho) << 4).
    //               TN: column group gj = wm0 / 8 + i (wm0 / 8 is even) of the [16 k][16 i] boxes.
    uint32_t aoff[4];
    if (!TRANS_A) {
        aoff[0] = (uint32_t)((wm0 + rho) * 128 + ((t4 ^ rho) << 4));
        aoff[1] = aoff[0] ^ 64u;                      // kg = 1
        aoff[2] = aoff[3] = 0;
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 2; ++i) aoff[2 * h + i] = (uint32_t)((warp >> 1) * 2048) + (L[h] ^ (uint32_t)(i << 6));
    }
    // B fragments: group gj = w1 * CG + j; offset = boff[h][j & 1] + (j >> 1) * 2048 + kg * 1024 (see above).
    uint32_t boff[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
            const int gj = w1 * CG + jp;              // group of j = jp; j = jp + 2 q adds q * 2048
            boff[h][jp] = (uint32_t)((gj >> 1) * 2048) + (L[h] ^ (uint32_t)((gj & 1) << 6));
        }
    // tail columns BN - 8 + 2 c2 (+1): box BN / 16 - 1, chunk (4 + c2) ^ kk_h
    uint32_t toff[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int kk = 2 * t4 + h;
        toff[h] = (uint32_t)((BN / 16 - 1) * 2048 + kk * 128 + ((4 ^ kk) << 4));
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int seq = 0;; ++seq) {
        mbar_wait(full0 + 8 * stage, phase);             // first stage of the next tile (or the sentinel)
        const int t = tile_ring[seq & 7];
        if (t < 0) break;
        // work-item order: n-chunk fastest, so the CTAs that share a tile of A run side by side and A comes from HBM
        // once and from L2 for the other chunks (m-tile fastest streamed the whole of A once per chunk)
        const int nc = t % prm.n_chunks;
        const int rest = t / prm.n_chunks;
        const int mt = rest % prm.m_tiles, sp = rest / prm.m_tiles;
        const int m0 = mt * BM, n0 = nc * BN;
        const int kbeg = sp * prm.k_chunk;
        const int kend = min(prm.K, kbeg + prm.k_chunk);
        double acc[RG][CGD > 0 ? CGD : 1][2];
        double tacc[RG][TW];
#pragma unroll
        for (int i = 0; i < RG; ++i) {
#pragma unroll
            for (int j = 0; j < CGD; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
            for (int c = 0; c < TW; ++c) tacc[i][c] = 0.0;
        }

        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            if (k0 != kbeg) mbar_wait(full0 + 8 * stage, phase);
            const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
            uint32_t dep = 0;
#pragma unroll
            for (int kg = 0; kg < BK / 8; ++kg) {
                double a[RG][2], b[CGD > 0 ? CGD : 1][2], bt[2][TW];
                if (!TRANS_A) {
#pragma unroll
                    for (int i = 0; i < RG; ++i) lds128(sa + aoff[kg] + i * 1024, a[i][0], a[i][1]);
                } else {
#pragma unroll
                    for (int i = 0; i < RG; ++i)
#pragma unroll
                        for (int h = 0; h < 2; ++h) a[i][h] = lds64(sa + aoff[2 * h + i] + kg * 1024);
                }
#pragma unroll
                for (int j = 0; j < CGD; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) b[j][h] = lds64(sb + boff[h][j & 1] + (j >> 1) * 2048 + kg * 1024);
                if (TAILW > 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int c2 = 0; c2 < TW / 2; ++c2)
                            lds128(sb + (toff[h] ^ (uint32_t)(c2 << 4)) + kg * 1024, bt[h][2 * c2], bt[h][2 * c2 + 1]);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < RG; ++i)
#pragma unroll
                        for (int j = 0; j < CGD; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i][h], b[j][h]);
                if (TAILW > 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < RG; ++i)
#pragma unroll
                            for (int c = 0; c < TW; ++c) tacc[i][c] = fma(a[i][h], bt[h][c], tacc[i][c]);
                    dep |= (uint32_t)__double2loint(bt[1][TW - 1]);
                }
                dep |= (uint32_t)__double2loint(a[RG - 1][1]);
                if (CGD > 0) dep |= (uint32_t)__double2loint(b[CGD > 0 ? CGD - 1 : 0][1]);
            }
            // Release the stage only once the fragments are in registers: the barrier address carries a
            // data dependence on the last loads (dep & 0 at run time).  Without it the arrive issues right
            // behind the LDS *issue* and a TMA refill could, in principle, overtake loads that are still
            // queued in a backed-up load/store pipe.
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * stage + (dep & prm.zero));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // ---- epilogue: accumulators -> global (direct or split partial)
        double* out = prm.d + (int64_t)sp * prm.part_stride;
        if (TAILW > 0) {          // sum the four k-slot partials of the tail columns (fixed order), whole warp
#pragma unroll
            for (int i = 0; i < RG; ++i)
#pragma unroll
                for (int c = 0; c < TW; ++c) {
                    double v = tacc[i][c];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    tacc[i][c] = v;
                }
        }
#pragma unroll
        for (int i = 0; i < RG; ++i) {
            const int row = m0 + wm0 + i * 8 + (TRANS_A ? g : rho);
            if (row >= prm.M) continue;
#pragma unroll
            for (int j = 0; j < CGD; ++j) {
                const int col = n0 + wn0 + j * 8 + 2 * t4;
                double* p = out + (int64_t)row * prm.ldd + col;
                if (col + 1 < prm.N) {
                    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
                    else { p[0] = acc[i][j][0]; p[1] = acc[i][j][1]; }
                } else if (col < prm.N) {
                    p[0] = acc[i][j][0];
                }
            }
            if (TAILW > 0) {      // every lane of the quad holds all the sums: lane t4 writes tail column t4
                const int col = n0 + BN - 8 + t4;
                if (t4 < TW && col < prm.N) {
                    double v = tacc[i][0];
#pragma unroll
                    for (int c = 1; c < TW; ++c) v = (t4 == c) ? tacc[i][c] : v;
                    out[(int64_t)row * prm.ldd + col] = v;
                }
            }
        }
    }
}

template <int BN, bool TRANS_A, int TAIL>
__global__ void __launch_bounds__(NTHREADS, 2)
dmma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, DmmaParams prm) {
    constexpr int A_BYTES = BM * BK * 8;             // 8 KB
    constexpr int B_BYTES = BK * BN * 8;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int STAGES = stages_for(BN);
    static_assert(BN % 16 == 0, "BN must be a multiple of 16");

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES];
    __shared__ int tile_ring[8];      // tile ids handed from the producer to the consumers
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int total = prm.m_tiles * prm.n_chunks * prm.splits;

    if (warp == NCW) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // Dynamic tile scheduler: an atomic counter hands out work items, so SMs that finish early
            // take more (a static round-robin over 2 CTAs/SM left ~13% of the SM-time idle in the tail).
            for (int seq = 0;; ++seq) {
                const int t = atomicAdd(prm.tile_counter, 1);
                if (t >= total) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    tile_ring[seq & 7] = -1;
                    mbar_arrive(full0 + 8 * stage);            // wake the consumers with the sentinel
                    break;
                }
                const int nc = t % prm.n_chunks;
                const int rest = t / prm.n_chunks;
                const int mt = rest % prm.m_tiles, sp = rest / prm.m_tiles;
                const int m0 = mt * BM, n0 = nc * BN;
                const int kbeg = sp * prm.k_chunk;
                const int kend = min(prm.K, kbeg + prm.k_chunk);
                for (int k0 = kbeg; k0 < kend; k0 += BK) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    if (k0 == kbeg) tile_ring[seq & 7] = t;    // published by the release of the arrive below
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint32_t fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    if (!TRANS_A) {
                        tma_load_2d(sa, &tmA, k0, m0, fb);                       // [BM rows][16 k]
                    } else {
#pragma unroll
                        for (int b = 0; b < BM / 16; ++b) tma_load_2d(sa + b * 2048, &tmA, m0 + 16 * b, k0, fb);   // [16 k][16 i]
                    }
#pragma unroll
                    for (int b = 0; b < BN / 16; ++b) tma_load_2d(sb + b * 2048, &tmB, n0 + 16 * b, k0, fb);       // [16 k][16 n]
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (TAIL > 0 && ((warp ^ (warp >> 2)) & 1)) {
        dmma_consumer<BN, TRANS_A, TAIL>(prm, smem_base, full0, empty0, tile_ring, warp, lane);   // owns the ragged last group
    } else {
        dmma_consumer<BN, TRANS_A, 0>(prm, smem_base, full0, empty0, tile_ring, warp, lane);
    }
}

// W (2n x 2l real, ldw) -> Z (n x l complex): Z = (W[2j][2c] + W[2j+1][2c+1]) + i (W[2j][2c+1] - W[2j+1][2c])
__global__ void combine_conj_kernel(int64_t n, int64_t l, const double* __restrict__ w, int64_t ldw, c64* __restrict__ z, int64_t ldz) {
    int64_t tot = n * l;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = e / l, c = e - j * l;
        const double* r0 = w + (2 * j) * ldw + 2 * c;
        const double* r1 = w + (2 * j + 1) * ldw + 2 * c;
        z[j * ldz + c] = c64(r0[0] + r1[1], r0[1] - r1[0]);
    }
}

template <int BN, bool TRANS_A, int TAIL>
void launch_dmma(rc_ctx* c, const CUtensorMap& tmA, const CUtensorMap& tmB, const DmmaParams& prm) {
    constexpr int STAGE_BYTES = BM * BK * 8 + BK * BN * 8;
    size_t smem = (size_t)stages_for(BN) * STAGE_BYTES + 1024;
    // the opt-in shared-memory size is a per-device function attribute: remember it per device, not per process
    static std::atomic<unsigned long long> configured{0};
    const unsigned long long bit = 1ull << (c->device & 63);
    if (!(configured.load(std::memory_order_relaxed) & bit)) {
        RC_CUDA(cudaFuncSetAttribute(dmma_gemm_kernel<BN, TRANS_A, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(bit, std::memory_order_relaxed);
    }
    int total = prm.m_tiles * prm.n_chunks * prm.splits;
    int grid = std::min(total, 2 * c->sm_count);
    RC_CUDA(cudaMemsetAsync(prm.tile_counter, 0, sizeof(int), c->stream));
    dmma_gemm_kernel<BN, TRANS_A, TAIL><<<grid, NTHREADS, smem, c->stream>>>(tmA, tmB, prm);
    RC_CHECK_LAUNCH(c);
}

template <bool TRANS_A, int TAIL>
void dispatch_bn(rc_ctx* c, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const DmmaParams& prm) {
    switch (bn) {
        case 16: launch_dmma<16, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
        case 32: launch_dmma<32, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
        case 48: launch_dmma<48, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
        case 64: launch_dmma<64, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
        case 80: launch_dmma<80, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
        default: launch_dmma<96, TRANS_A, TAIL>(c, tmA, tmB, prm); break;
    }
}

template <bool TRANS_A>
void dispatch_tail(rc_ctx* c, int bn, int tail, const CUtensorMap& tmA, const CUtensorMap& tmB, const DmmaParams& prm) {
    if (tail == 2) dispatch_bn<TRANS_A, 2>(c, bn, tmA, tmB, prm);
    else if (tail == 4) dispatch_bn<TRANS_A, 4>(c, bn, tmA, tmB, prm);
    else dispatch_bn<TRANS_A, 0>(c, bn, tmA, tmB, prm);
}

}  // namespace

bool gemm_dmma_f64(rc_ctx* c, bool a_transposed, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
    if ((lda & 1) || (ldb & 1)) return false;                 // TMA: 16-byte row pitch
    if (M > (1LL << 30) || N > (1LL << 30) || K > (1LL << 30)) return false;
    DmmaParams prm;
    prm.M = (int)M; prm.N = (int)N; prm.K = (int)K;
    prm.m_tiles = (int)((M + BM - 1) / BM);
    const int64_t grid_cap = 2LL * c->sm_count;
    int bn = 96, n_chunks = 1, splits = 1;
    if (N <= 96) {
        // one chunk: the narrowest multiple of 16 that holds N
        bn = (int)((N + 15) / 16 * 16);
        const int64_t tiles = prm.m_tiles;
        if (tiles < 2 * grid_cap) {
            // enough work items for ~4 rounds of the persistent grid, picked so the last round is full
            int64_t want = (3 * grid_cap + tiles - 1) / tiles;
            int64_t maxs = std::max<int64_t>(1, K / (BK * 16));
            double best_eff = -1.0;
            for (int64_t s = std::max<int64_t>(1, want); s <= std::min<int64_t>(maxs, 2 * want + 2); ++s) {
                int64_t tot = tiles * s, rounds = (tot + grid_cap - 1) / grid_cap;
                double eff = (double)tot / (double)(rounds * grid_cap);
                if (eff > best_eff + 0.02) { best_eff = eff; splits = (int)s; }
            }
            if (best_eff < 0) splits = (int)std::max<int64_t>(1, std::min<int64_t>(want, maxs));
        }
    } else {
        // Several column chunks (the real expansion of the c64 products, wide f64 operands): tile width and split-K
        // factor by a cost model -- time ~ padded work / (pipe efficiency of the tile shape x how evenly the work
        // items fill the SMs).  The two CTAs of an SM share one FP64 pipe, so the quantisation that matters is items
        // per SM.  Narrower tiles can waste fewer padded columns (N = 256 real columns: 4 x 64 instead of 3 x 96 with
        // 12 % padding), and a split-K of 2-4 fills the last round even when there are plenty of tiles (config 5,
        // Y = A Omega: 768 items on 148 SMs leave the last of 6 rounds 19 % full; 3 splits fill 15.6 of 16 rounds and
        // the partial sums cost ~1 % of the GEMM).
        double best = -1.0;
        const int64_t maxs_k = std::max<int64_t>(1, K / (BK * 16));
        for (int cand = 96; cand >= 48; cand -= 16) {
            const int64_t nch = (N + cand - 1) / cand;
            const double pad_eff = (double)N / (double)(nch * cand);
            const double tile_eff = cand >= 80 ? 1.0 : (cand == 64 ? 0.97 : 0.93);
            const int64_t tiles = (int64_t)prm.m_tiles * nch;
            const int64_t smax = std::min<int64_t>(maxs_k, tiles < 2 * grid_cap ? 4 * grid_cap / std::max<int64_t>(1, tiles) + 2 : 4);
            for (int64_t sp = 1; sp <= std::max<int64_t>(1, smax); ++sp) {
                const int64_t items = tiles * sp;
                const int64_t rounds = (items + c->sm_count - 1) / c->sm_count;
                const double fill = (double)items / (double)(rounds * c->sm_count);
                const double occ = items >= grid_cap ? 1.0 : 0.5 + 0.5 * (double)items / (double)grid_cap;
                const double red = sp > 1 ? 1.0 / (1.0 + 0.004 * (double)sp * 16384.0 / (double)std::max<int64_t>(K, 1)) : 1.0;
                const double score = pad_eff * tile_eff * fill * occ * red;
                if (score > best + 1e-3) { best = score; bn = cand; n_chunks = (int)nch; splits = (int)sp; }
            }
        }
    }
    prm.n_chunks = n_chunks;
    int64_t k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (int)((K + k_chunk - 1) / k_chunk);
    prm.splits = splits;
    prm.k_chunk = (int)k_chunk;

    CUtensorMap tmA = a_transposed ? make_map(A, K, M, lda, BK) : make_map(A, M, K, lda, BM);
    CUtensorMap tmB = make_map(B, K, N, ldb, BK);

    if (!c->tile_counter) RC_CUDA(cudaMalloc((void**)&c->tile_counter, 256));
    prm.tile_counter = c->tile_counter;
    prm.zero = 0;
    DevBuf<double> part;
    if (splits == 1) {
        prm.d = C; prm.ldd = ldc; prm.part_stride = 0;
    } else {
        int64_t ldp = (N + 1) & ~1LL;
        part.alloc(c, (size_t)splits * M * ldp);
        prm.d = part.p; prm.ldd = ldp; prm.part_stride = M * ldp;
    }
    // ragged output (N = bn - 8 + 2 or + 4, single chunk): the last column group goes to the DFMA tail path
    const int tail = (n_chunks == 1 && c->dmma_tail && (bn - N == 6 || bn - N == 4)) ? (int)(N - (bn - 8)) : 0;
    if (a_transposed) dispatch_tail<true>(c, bn, tail, tmA, tmB, prm);
    else dispatch_tail<false>(c, bn, tail, tmA, tmB, prm);
    if (splits > 1) rc_splitk::reduce<double>(c, M, N, splits, part.p, prm.ldd, prm.part_stride, C, ldc);
    c->gemm_flops += 2 * M * N * K;
    return true;
}

// c64 through an exact real expansion (same flop count as the 4-real-product formulation):
//   NN:  (A viewed as real m x 2n) * X' = (A X viewed as real m x 2l),  X' rows 2k / 2k+1 = X_k / i X_k
//   TN:  W = (A viewed as real)^T * (X viewed as real) is 2n x 2l;  Z = A^H X is recombined from W.
bool gemm_dmma_c64(rc_ctx* c, bool a_conj_transposed, int64_t M, int64_t N, int64_t K, const c64* A, int64_t lda,
                   const c64* B, int64_t ldb, c64* C, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    if (!a_conj_transposed) {
        // A: M x K complex ; B: K x N complex
        int64_t ldx = 2 * N;
        DevBuf<double> xp(c, (size_t)(2 * K) * ldx);
        k_expand_rhs_c64(c, xp.p, ldx, B, ldb, K, N);
        return gemm_dmma_f64(c, false, M, 2 * N, 2 * K, reinterpret_cast<const double*>(A), 2 * lda, xp.p, ldx,
                             reinterpret_cast<double*>(C), 2 * ldc);
    }
    // A stored K x M complex (we want A^H B with A: K x M) ; B: K x N complex ; C: M x N
    int64_t ldw = 2 * N;
    DevBuf<double> w(c, (size_t)(2 * M) * ldw);
    bool ok = gemm_dmma_f64(c, true, 2 * M, 2 * N, K, reinterpret_cast<const double*>(A), 2 * lda,
                            reinterpret_cast<const double*>(B), 2 * ldb, w.p, ldw);
    if (!ok) return false;
    int64_t tot = M * N;
    int nb = (int)std::min<int64_t>((tot + 255) / 256, 148 * 8);
    combine_conj_kernel<<<nb, 256, 0, c->stream>>>(M, N, w.p, ldw, C, ldc);
    RC_CHECK_LAUNCH(c);
    return true;
}
