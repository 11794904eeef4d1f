// Tall-skinny Householder QR (TSQR) and application of its implicit Q.
//
// Replaces LAPACK ?geqp3's factorisation work and ?orgqr/?ungqr (reference N1/N2:
// src/pivoted_qr.rs:104-111, 139-173) for the tall sketches Y = A*Omega: an unpivoted
// Householder TSQR gives Y = Q0 R0 in one pass over Y; the column pivoting is then done on the
// small w x w factor R0 (pivqr.cu), which yields the same pivots as pivoting Y itself because
// Q0 is orthogonal (DESIGN.md "pivot parity").
//
// Each CTA owns a block of `block` consecutive rows held COLUMN-major in shared memory and runs
// the LAPACK ?larfg/?larf recurrences on it (norms accumulated in double, warp-shuffle
// reductions, one warp per trailing column).  The w x w R factors of `g` consecutive blocks are
// stacked and factored again, level by level, until one remains.
#include "rc_internal.cuh"

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;

template <class T>
struct HouseScalars {
    T tau;      // H = I - tau v v^H
    T scale;    // v[1:] = x[1:] * scale
    T beta;     // resulting diagonal entry (real)
};

// LAPACK ?larfg: alpha = x[0], xnorm2 = ||x[1:]||^2.  H^H x = beta e_0.
template <class T>
__device__ __forceinline__ HouseScalars<T> larfg(T alpha, double xnorm2) {
    HouseScalars<T> h;
    double ar = (double)rc_real(alpha), ai = (double)rc_imag(alpha);
    if (xnorm2 == 0.0 && ai == 0.0) {
        h.tau = rc_zero<T>();
        h.scale = rc_zero<T>();
        h.beta = alpha;
        return h;
    }
    double beta = -copysign(sqrt(ar * ar + ai * ai + xnorm2), ar);
    // tau = ((beta - ar)/beta, -ai/beta) ; scale = 1 / (alpha - beta)
    double dr = ar - beta, di = ai;
    double den = dr * dr + di * di;
    h.tau = rc_make<T>((beta - ar) / beta, -ai / beta);
    h.scale = rc_make<T>(dr / den, -di / den);
    h.beta = rc_make<T>(beta, 0.0);
    return h;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = rc_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += red[i];
    return s;
}

// Factor kernel.  y: rows x w (row-major, ld); block rows per CTA (>= w); sp = padded block.
// In place: R on/above the diagonal of the block's first w rows, reflectors below.
// rstack: nblocks x (w x w) row-major upper-triangular copies; tau: nblocks x w.
template <class T>
__global__ void __launch_bounds__(NT)
house_block_qr_kernel(T* __restrict__ y, int64_t ld, int64_t rows, int w, int block, int sp,
                      T* __restrict__ rstack, T* __restrict__ tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);            // w columns of sp
    T* vbuf = S + (size_t)w * sp;                      // sp
    __shared__ double red[NW];
    __shared__ double s_norm2;

    const int64_t r0 = (int64_t)blockIdx.x * block;
    const int nrows = (int)((rows - r0 < block) ? rows - r0 : block);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < block * w; e += NT) {
        int r = e / w, c = e - r * w;
        S[(size_t)c * sp + r] = (r < nrows) ? y[(r0 + r) * ld + c] : rc_zero<T>();
    }
    __syncthreads();
    {   // norm of column 0 below the diagonal
        double a = 0.0;
        for (int r = 1 + tid; r < block; r += NT) a += rc_abs2(S[r]);
        a = block_sum(a, red);
        if (tid == 0) s_norm2 = a;
        __syncthreads();
    }
    const int steps = (w < block) ? w : block;
    for (int j = 0; j < steps; ++j) {
        T* colj = S + (size_t)j * sp;
        HouseScalars<T> h = larfg<T>(colj[j], s_norm2);
        // v = [1 ; x * scale]
        for (int r = j + 1 + tid; r < block; r += NT) {
            T v = colj[r] * h.scale;
            vbuf[r] = v;
            colj[r] = v;
        }
        if (tid == 0) tau_out[(int64_t)blockIdx.x * w + j] = h.tau;
        __syncthreads();
        if (tid == 0) colj[j] = h.beta;   // nobody reads colj[j] again before the write-back
        const T ctau = rc_conj(h.tau);
        for (int c = j + 1 + warp; c < w; c += NW) {
            T* colc = S + (size_t)c * sp;
            T part = rc_zero<T>();
            for (int r = j + 1 + lane; r < block; r += 32) part = rc_cfma(vbuf[r], colc[r], part);
            part = rc_warp_sum(part);
            T f = ctau * (colc[j] + part);
            double nrm = 0.0;
            for (int r = j + 1 + lane; r < block; r += 32) {
                T v = colc[r] - f * vbuf[r];
                colc[r] = v;
                if (c == j + 1 && r > j + 1) nrm += rc_abs2(v);
            }
            __syncwarp();
            if (lane == 0) colc[j] = colc[j] - f;
            if (c == j + 1) {
                nrm = rc_warp_sum(nrm);
                if (lane == 0) s_norm2 = nrm;
            }
        }
        __syncthreads();
    }
    // tau for the (identity) reflectors beyond `steps`
    for (int j = steps + tid; j < w; j += NT) tau_out[(int64_t)blockIdx.x * w + j] = rc_zero<T>();
    // write back V/R in place and the R copy
    for (int e = tid; e < nrows * w; e += NT) {
        int r = e / w, c = e - r * w;
        y[(r0 + r) * ld + c] = S[(size_t)c * sp + r];
    }
    T* rout = rstack + (int64_t)blockIdx.x * w * w;
    for (int e = tid; e < w * w; e += NT) {
        int r = e / w, c = e - r * w;
        rout[e] = (c >= r && r < block) ? S[(size_t)c * sp + r] : rc_zero<T>();
    }
}

// Apply kernel: out block (block rows x ncc cols) = H_0 ... H_{w-1} [ctop_chunk ; 0].
// v: rows x w reflectors (row-major ldv) as left by the factor kernel; cin: chunk b is the w x nc
// matrix at rows [b*w, (b+1)*w) of cin (row-major ldc).  grid.y splits the nc columns.
template <class T>
__global__ void __launch_bounds__(NT)
house_block_apply_kernel(const T* __restrict__ v, int64_t ldv, int64_t rows, int w, int block, int sp,
                         const T* __restrict__ tau, const T* __restrict__ cin, int64_t ldc, int nc, int ncc,
                         T* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Sv = reinterpret_cast<T*>(smem_raw);            // w columns of sp
    T* Sc = Sv + (size_t)w * sp;                       // ncc columns of sp
    T* stau = Sc + (size_t)ncc * sp;                   // w
    const int64_t r0 = (int64_t)blockIdx.x * block;
    const int nrows = (int)((rows - r0 < block) ? rows - r0 : block);
    const int c0 = blockIdx.y * ncc;
    const int ncl = (nc - c0 < ncc) ? nc - c0 : ncc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < block * w; e += NT) {
        int r = e / w, c = e - r * w;
        Sv[(size_t)c * sp + r] = (r < nrows) ? v[(r0 + r) * ldv + c] : rc_zero<T>();
    }
    for (int e = tid; e < block * ncl; e += NT) {
        int r = e / ncl, c = e - r * ncl;
        Sc[(size_t)c * sp + r] = (r < w) ? cin[((int64_t)blockIdx.x * w + r) * ldc + c0 + c] : rc_zero<T>();
    }
    for (int j = tid; j < w; j += NT) stau[j] = tau[(int64_t)blockIdx.x * w + j];
    __syncthreads();
    const int steps = (w < block) ? w : block;
    for (int c = warp; c < ncl; c += NW) {
        T* colc = Sc + (size_t)c * sp;
        for (int j = steps - 1; j >= 0; --j) {
            const T* vj = Sv + (size_t)j * sp;
            T part = rc_zero<T>();
            for (int r = j + 1 + lane; r < block; r += 32) part = rc_cfma(vj[r], colc[r], part);
            part = rc_warp_sum(part);
            T f = stau[j] * (colc[j] + part);
            for (int r = j + 1 + lane; r < block; r += 32) colc[r] = colc[r] - f * vj[r];
            __syncwarp();
            if (lane == 0) colc[j] = colc[j] - f;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < nrows * ncl; e += NT) {
        int r = e / ncl, c = e - r * ncl;
        out[(r0 + r) * ldo + c0 + c] = Sc[(size_t)c * sp + r];
    }
}

inline size_t smem_budget(rc_ctx* c) {
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    return lim - 2048;   // static smem + slack
}

template <class T>
void plan_block(rc_ctx* c, int64_t w, int64_t rows, int64_t& block, int64_t& sp) {
    // factor kernel keeps (w + 1) columns of sp; prefer <= ~100 KB so two CTAs share an SM
    size_t target = std::min<size_t>(smem_budget(c), (size_t)100 * 1024);
    int64_t bmax = (int64_t)(target / sizeof(T) / (w + 1)) - 1;
    if (bmax < 2 * w) bmax = 2 * w;
    int64_t g = bmax / w;
    if (g < 2) g = 2;
    block = g * w;
    if (block > rows) block = std::max<int64_t>(rows, w);
    sp = block | 1;
    RC_REQUIRE((size_t)(w + 1) * sp * sizeof(T) <= smem_budget(c), "tsqr: panel width %lld too large for shared memory", (long long)w);
}

}  // namespace

int64_t tsqr_max_width(rc_ctx* c, int dtype) {
    size_t sz = rc_dtype_size(dtype);
    size_t budget = smem_budget(c);
    int64_t w = 8;
    // need (w + 1) * (2w | 1) for the factor and (w + 8 + 1) * (2w | 1) + w for the apply
    while ((size_t)((w + 1) + 8 + 1) * ((2 * (w + 1)) | 1) * sz + (w + 1) * sz <= budget) ++w;
    return w;
}

template <class T>
TsqrFactor<T>::~TsqrFactor() {
    for (auto& L : levels) {
        if (L.owns_v && L.v) cudaFreeAsync(L.v, ctx->stream);
        if (L.tau) cudaFreeAsync(L.tau, ctx->stream);
    }
    if (r) cudaFreeAsync(r, ctx->stream);
}

template <class T>
void tsqr_factor(rc_ctx* c, T* y, int64_t ld, int64_t m, int64_t w, TsqrFactor<T>& f) {
    RC_REQUIRE(m > 0 && w > 0, "tsqr: empty matrix");
    f.ctx = c;
    f.m = m;
    f.w = w;
    T* cur = y;
    int64_t cur_ld = ld, cur_rows = m;
    bool owns = false;
    for (;;) {
        int64_t block, sp;
        plan_block<T>(c, w, cur_rows, block, sp);
        int64_t nblocks = (cur_rows + block - 1) / block;
        typename TsqrFactor<T>::Level L;
        L.v = cur; L.ldv = cur_ld; L.rows = cur_rows; L.block = block; L.nblocks = nblocks; L.owns_v = owns;
        RC_CUDA(cudaMallocAsync((void**)&L.tau, sizeof(T) * nblocks * w, c->stream));
        T* rstack = nullptr;
        RC_CUDA(cudaMallocAsync((void**)&rstack, sizeof(T) * nblocks * w * w, c->stream));
        size_t smem = (size_t)(w + 1) * sp * sizeof(T);
        RC_CUDA(cudaFuncSetAttribute(house_block_qr_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        house_block_qr_kernel<T><<<(unsigned)nblocks, NT, smem, c->stream>>>(cur, cur_ld, cur_rows, (int)w, (int)block, (int)sp, rstack, L.tau);
        RC_CHECK_LAUNCH(c);
        f.levels.push_back(L);
        if (nblocks == 1) { f.r = rstack; break; }
        cur = rstack; cur_ld = w; cur_rows = nblocks * w; owns = true;
    }
}

template <class T>
void tsqr_apply_q(rc_ctx* c, const TsqrFactor<T>& f, const T* ctop, int64_t ldc, int64_t nc, T* out, int64_t ldo) {
    if (nc == 0) return;
    const int64_t w = f.w;
    const T* cin = ctop;
    int64_t cin_ld = ldc;
    DevBuf<T> tmp[2];
    int flip = 0;
    for (int l = (int)f.levels.size() - 1; l >= 0; --l) {
        const auto& L = f.levels[l];
        int64_t sp = L.block | 1;
        // columns per CTA limited by shared memory
        size_t budget = smem_budget(c);
        size_t fixed = ((size_t)w * sp + w) * sizeof(T);
        RC_REQUIRE(fixed + (size_t)sp * sizeof(T) <= budget, "tsqr_apply: panel too wide");
        int64_t ncc = (int64_t)((budget - fixed) / (sp * sizeof(T)));
        if (ncc > nc) ncc = nc;
        // keep a few column chunks so small levels still spread over SMs
        int64_t nchunks = (nc + ncc - 1) / ncc;
        if (L.nblocks * nchunks < c->sm_count && ncc > 8) {
            int64_t want = std::min<int64_t>((c->sm_count + L.nblocks - 1) / L.nblocks, (nc + 7) / 8);
            ncc = (nc + want - 1) / want;
            nchunks = (nc + ncc - 1) / ncc;
        }
        T* dst; int64_t dst_ld;
        if (l == 0) { dst = out; dst_ld = ldo; }
        else { tmp[flip].alloc(c, (size_t)L.rows * nc); dst = tmp[flip].p; dst_ld = nc; }
        size_t smem = ((size_t)(w + ncc) * sp + w) * sizeof(T);
        RC_CUDA(cudaFuncSetAttribute(house_block_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)L.nblocks, (unsigned)nchunks);
        house_block_apply_kernel<T><<<grid, NT, smem, c->stream>>>(L.v, L.ldv, L.rows, (int)w, (int)L.block, (int)sp, L.tau,
                                                                  cin, cin_ld, (int)nc, (int)ncc, dst, dst_ld);
        RC_CHECK_LAUNCH(c);
        cin = dst; cin_ld = dst_ld;
        flip ^= 1;
    }
}

#define INST(T)                                                                              \
    template struct TsqrFactor<T>;                                                           \
    template void tsqr_factor<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, TsqrFactor<T>&);     \
    template void tsqr_apply_q<T>(rc_ctx*, const TsqrFactor<T>&, const T*, int64_t, int64_t, T*, int64_t);
INST(float)
INST(double)
INST(c32)
INST(c64)
