// Column-pivoted Householder QR: the B200 replacement for LAPACK ?geqp3 + ?orgqr/?ungqr as the
// crate calls them (reference N1/N2: src/pivoted_qr.rs:104-111, 139-173).
//
// One persistent cooperative kernel runs all min(p, n) steps with ONE grid-wide sync per step:
//   argmax of the partial column norms (first maximum wins, like ?geqp3's idamax) ->
//   Householder reflector of the pivot column (LAPACK ?larfg) -> rank-1 update of the trailing
//   columns (?larf), each warp owning a fixed set of columns.
// Columns are never physically swapped; a logical position per physical column stands for
// ?geqp3's swaps, which reproduces its pivot order including tie-breaks.  Partial norms are
// recomputed exactly (in double) in the same pass that updates a column, instead of ?geqp3's
// downdate-with-safeguard: the two agree wherever the pivot gap exceeds sqrt(eps) effects, and
// the exact norm costs nothing extra here because the column is being touched anyway.
#include <cooperative_groups.h>
#include "rc_internal.cuh"

namespace cg = cooperative_groups;

namespace {

// NT threads per CTA: 256 for the multi-CTA (wide) case, 1024 when one CTA holds the whole
// matrix in shared memory (small factors such as the w x w R of a sketch).

struct __align__(16) Cand {
    double val;
    int lpos;   // logical position (tie-break: smaller wins)
    int phys;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
    // a better than b ?
    if (a.phys < 0) return false;
    if (b.phys < 0) return true;
    if (a.val != b.val) return a.val > b.val;
    return a.lpos < b.lpos;
}
__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        Cand o;
        o.val = __shfl_xor_sync(0xffffffffu, c.val, m);
        o.lpos = __shfl_xor_sync(0xffffffffu, c.lpos, m);
        o.phys = __shfl_xor_sync(0xffffffffu, c.phys, m);
        if (better(o, c)) c = o;
    }
    return c;
}

template <class T>
__device__ __forceinline__ void larfg_dev(T alpha, double xnorm2, T& tau, T& scale, T& beta) {
    double ar = (double)rc_real(alpha), ai = (double)rc_imag(alpha);
    if (xnorm2 == 0.0 && ai == 0.0) {
        tau = rc_zero<T>(); scale = rc_zero<T>(); beta = alpha;
        return;
    }
    double b = -copysign(sqrt(ar * ar + ai * ai + xnorm2), ar);
    double dr = ar - b, di = ai, den = dr * dr + di * di;
    tau = rc_make<T>((b - ar) / b, -ai / b);
    scale = rc_make<T>(dr / den, -di / den);
    beta = rc_make<T>(b, 0.0);
}

// L1-bypassing load: the pivot column was last written by another SM (before the grid-wide sync)
// and a 128-byte line may straddle two columns owned by different SMs, so never serve it from L1.
__device__ __forceinline__ float  ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ c32 ld_cg(const c32* p) { float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return c32(v.x, v.y); }
__device__ __forceinline__ c64 ld_cg(const c64* p) { double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return c64(v.x, v.y); }

// Wide factors (the k x n factor b = Q^H A of compute_from_range_estimate, C^H in two_sided_id): one CTA of NT
// threads per SM; CTA b owns the columns c = b (mod gridDim.x) and keeps as many of them as fit RESIDENT IN
// SHARED MEMORY for the whole factorisation (all of them at config 5, 80 % at config 3), the rest stay in
// global memory (L2).  One grid-wide sync per step: before it every CTA publishes its best candidate (norm,
// logical position) AND that candidate's column to a global staging slot, so after the sync everybody reads
// the winning column from the winner's slot -- no second sync to fetch a column that lives in another SM's
// shared memory.  Every warp updates CU of its columns at a time (CU independent load -> dot ->
// shuffle-reduce -> update chains in flight; the step is latency-bound, not bandwidth-bound).
template <class T, int NT, int LPC, int CU>
__global__ void __launch_bounds__(NT)
pivqr_kernel(T* __restrict__ W, int64_t ldw, int p, int n, int kk, int cap, int nlmax,
             int* __restrict__ ind, T* __restrict__ vbuf, T* __restrict__ tau_out, T* __restrict__ diag,
             Cand* __restrict__ slots, int* __restrict__ slots_disp, T* __restrict__ stage,
             long long* __restrict__ prof) {
    constexpr int NW = NT / 32;
    cg::grid_group grid = cg::this_grid();
    // optional phase timer (option "trace"): cycles of block 0 spent in each phase, summed over the steps
    const bool timing = (prof != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_prev = 0;
#define RC_PHASE(k) do { if (timing) { long long t_now = clock64(); prof[k] += t_now - t_prev; t_prev = t_now; } } while (0)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);                  // pivot column / reflector, p entries
    T* scol = xs + p;                                        // cap resident columns, p entries each
    double* vn = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(scol + (size_t)cap * p) + 7) & ~(uintptr_t)7);
    int* lpos = reinterpret_cast<int*>(vn + nlmax);
    __shared__ Cand s_cand[NW];
    __shared__ int s_disp[NW];
    __shared__ double s_red[NW];
    __shared__ Cand s_win, s_bbest;
    __shared__ int s_windisp;
    __shared__ T s_hs[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int nl = (b < n) ? (n - b + G - 1) / G : 0;         // columns owned by this CTA: c = li * G + b
    auto colp = [&](int li) -> T* { return (li < cap) ? scol + (size_t)li * p : W + (int64_t)(li * G + b) * ldw; };

    // resident columns -> shared memory; initial norms and identity logical order
    for (int li = warp; li < nl; li += NW) {
        const T* src = W + (int64_t)(li * G + b) * ldw;
        T* dst = colp(li);
        double a = 0.0;
        for (int r = lane; r < p; r += 32) { T v = src[r]; if (li < cap) dst[r] = v; a += rc_abs2(v); }
        a = rc_warp_sum(a);
        if (lane == 0) { vn[li] = sqrt(a); lpos[li] = li * G + b; }
    }
    __syncthreads();

    if (timing) t_prev = clock64();
    for (int i = 0; i < kk; ++i) {
        // (a) local candidates over owned, not yet pivoted columns
        Cand best; best.val = -1.0; best.lpos = 0x7fffffff; best.phys = -1;
        int disp = -1;
        for (int li = tid; li < nl; li += NT) {
            int lp = lpos[li];
            if (lp >= i) {
                Cand cnd; cnd.val = vn[li]; cnd.lpos = lp; cnd.phys = li * G + b;
                if (better(cnd, best)) best = cnd;
                if (lp == i) disp = li * G + b;
            }
        }
        best = warp_best(best);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) disp = max(disp, __shfl_xor_sync(0xffffffffu, disp, m));
        if (lane == 0) { s_cand[warp] = best; s_disp[warp] = disp; }
        __syncthreads();
        if (warp == 0) {
            Cand bb; bb.val = -1.0; bb.lpos = 0x7fffffff; bb.phys = -1;
            int d = -1;
            if (lane < NW) { bb = s_cand[lane]; d = s_disp[lane]; }
            bb = warp_best(bb);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, m));
            if (lane == 0) {
                s_bbest = bb;
                slots[(i & 1) * G + b] = bb;
                slots_disp[(i & 1) * G + b] = d;
            }
        }
        __syncthreads();
        // publish this CTA's best column (rows i .. p-1) next to its candidate
        if (s_bbest.phys >= 0) {
            const T* src = colp(s_bbest.phys / G);
            T* dst = stage + ((size_t)(i & 1) * G + b) * p;
            for (int r = i + tid; r < p; r += NT) dst[r] = src[r];
        }
        RC_PHASE(0);      // candidates + publish
        // (b) one grid-wide sync per step
        grid.sync();
        RC_PHASE(1);      // grid sync
        // (c) global winner (every CTA reduces the same slots -> same answer): one slot per thread (written by
        // other SMs: L1 bypassed), so the L2 round trip is paid once, then two levels of warp reductions
        {
            const int nsw = (G + 31) / 32;           // warps that hold slots
            if (warp < nsw) {
                Cand bb; bb.val = -1.0; bb.lpos = 0x7fffffff; bb.phys = -1;
                int d = -1;
                if (tid < G) {
                    const int4 raw = __ldcg(reinterpret_cast<const int4*>(slots + (i & 1) * G + tid));
                    bb.val = __hiloint2double(raw.y, raw.x); bb.lpos = raw.z; bb.phys = raw.w;
                    d = __ldcg(slots_disp + (i & 1) * G + tid);
                }
                bb = warp_best(bb);
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, m));
                if (lane == 0) { s_cand[warp] = bb; s_disp[warp] = d; }
            }
            __syncthreads();
            if (warp == 0) {
                Cand bb; bb.val = -1.0; bb.lpos = 0x7fffffff; bb.phys = -1;
                int d = -1;
                if (lane < nsw) { bb = s_cand[lane]; d = s_disp[lane]; }
                bb = warp_best(bb);
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, m));
                if (lane == 0) { s_win = bb; s_windisp = d; }
            }
        }
        __syncthreads();
        const int dc = s_windisp;           // physical column sitting at logical position i
        // (no winner = every candidate norm is NaN, possible only in a speculative run on garbage, rc_ctx::defer_depth:
        // keep the column at position i so that all indices stay in range)
        const int pv = s_win.phys >= 0 ? s_win.phys : dc;          // physical pivot column
        const int pv_lpos = s_win.phys >= 0 ? s_win.lpos : i;      // where it sat logically
        // (d) logical swap, done by the owners
        if (tid == 0) {
            if (dc >= 0 && dc != pv && (dc % G) == b) lpos[dc / G] = pv_lpos;
            if ((pv % G) == b) lpos[pv / G] = i;
        }
        if (b == 0 && tid == 0) ind[i] = pv;
        RC_PHASE(2);      // winner
        // (e) reflector from the staged pivot column (rows i..p-1), redundantly per CTA
        const T* pcol = stage + ((size_t)(i & 1) * G + (pv % G)) * p;
        double a = 0.0;
        for (int r = i + tid; r < p; r += NT) {
            T v = ld_cg(pcol + r);
            xs[r] = v;
            if (r > i) a += rc_abs2(v);
        }
        a = rc_warp_sum(a);
        if (lane == 0) s_red[warp] = a;
        __syncthreads();
        if (warp == 0) {     // every lane of warp 0 derives the same reflector scalars
            double xnorm2 = (lane < NW) ? s_red[lane] : 0.0;
            xnorm2 = rc_warp_sum(xnorm2);
            T t0, t1, t2;
            larfg_dev<T>(xs[i], xnorm2, t0, t1, t2);
            if (lane == 0) { s_hs[0] = t0; s_hs[1] = t1; s_hs[2] = t2; }
        }
        __syncthreads();
        const T tau = s_hs[0], scale = s_hs[1], beta = s_hs[2];
        for (int r = i + 1 + tid; r < p; r += NT) xs[r] = xs[r] * scale;
        __syncthreads();
        if (b == 0) {
            T* vcol = vbuf + (int64_t)i * p;
            for (int r = tid; r < p; r += NT) vcol[r] = (r < i) ? rc_zero<T>() : (r == i ? rc_one<T>() : xs[r]);
            if (tid == 0) { tau_out[i] = tau; diag[i] = beta; }
        }
        RC_PHASE(3);      // reflector
        // (f) trailing update of owned columns + exact partial norms.  LPC lanes share a column (the rows of a
        // k x n factor are few: 10-20 per lane), so a warp works on (32 / LPC) * CU columns at once and the
        // dependent chain per column is one load pass, a log2(LPC)-step reduction, one update pass.  Dots and
        // norms are accumulated in the working precision: every step recomputes the norms from scratch (no
        // downdating), which is already tighter than ?geqp3, and the f32 -> f64 conversions made this loop
        // issue-bound for single precision (22 000 of 36 000 cycles per step at config 3).
        const T ctau = rc_conj(tau);
        {
            constexpr int CPW = 32 / LPC;
            const int sl = lane % LPC, sc = lane / LPC;
            using R = RealOf<T>;
            const int nres = min(nl, cap);           // resident columns: this loop; the others: the loop below
            for (int q0 = 0; warp + q0 * NW < nres; q0 += CU * CPW) {
                T* col[CU];
                bool act[CU];
                int lis[CU];
#pragma unroll
                for (int j = 0; j < CU; ++j) {
                    const int li = warp + (q0 + j * CPW + sc) * NW;
                    lis[j] = li;
                    act[j] = (li < nres) && (lpos[min(li, nl - 1)] > i);
                    col[j] = scol + (size_t)min(li, nres - 1) * p;
                }
                T part[CU];
#pragma unroll
                for (int j = 0; j < CU; ++j) part[j] = rc_zero<T>();
                for (int r = i + 1 + sl; r < p; r += LPC) {
                    const T xv = xs[r];
#pragma unroll
                    for (int j = 0; j < CU; ++j)
                        if (act[j]) part[j] = rc_cfma(xv, col[j][r], part[j]);
                }
#pragma unroll
                for (int j = 0; j < CU; ++j) {
#pragma unroll
                    for (int m = LPC / 2; m > 0; m >>= 1) part[j] = part[j] + rc_shfl_xor(part[j], m);
                }
                T ci[CU], f[CU];
                R nrm[CU];
#pragma unroll
                for (int j = 0; j < CU; ++j) {
                    ci[j] = act[j] ? col[j][i] : rc_zero<T>();
                    f[j] = ctau * (ci[j] + part[j]);
                    nrm[j] = R(0);
                }
                for (int r = i + 1 + sl; r < p; r += LPC) {
                    const T xv = xs[r];
#pragma unroll
                    for (int j = 0; j < CU; ++j)
                        if (act[j]) {
                            T v = col[j][r] - f[j] * xv;
                            col[j][r] = v;
                            nrm[j] += rc_real(v) * rc_real(v) + rc_imag(v) * rc_imag(v);
                        }
                }
#pragma unroll
                for (int j = 0; j < CU; ++j) {
#pragma unroll
                    for (int m = LPC / 2; m > 0; m >>= 1) nrm[j] += __shfl_xor_sync(0xffffffffu, nrm[j], m);
                }
                if (sl == 0) {
#pragma unroll
                    for (int j = 0; j < CU; ++j)
                        if (act[j]) { col[j][i] = ci[j] - f[j]; vn[lis[j]] = sqrt((double)nrm[j]); }
                }
            }
            // Columns that did not fit into shared memory live in global memory (L2).  One column per warp at
            // a time with all its rows PREFETCHED INTO REGISTERS (RG per lane, loads issued back to back): the
            // column costs one L2 round trip instead of one per row iteration -- mixed into the loop above they
            // stretched the update phase from ~5 000 to 21 000 cycles per step at config 3.
            // (10 rows per lane for the 4- and 8-byte scalars: every step of a 320-row factor -- config 3 with the pivot
            // decisions in double -- takes the prefetched path; with 8 the first 64 steps fell back to the row loop below,
            // one dependent L2 round trip per 32 rows, and the update phase averaged 29 000 cycles per step)
            constexpr int RG = sizeof(T) <= 8 ? 10 : 4;
            for (int li = cap + ((warp - cap % NW + NW) % NW); li < nl; li += NW) {
                if (lpos[li] <= i) continue;                       // warp-uniform
                T* col = W + (int64_t)(li * G + b) * ldw;
                const int nrows = p - (i + 1);
                T part = rc_zero<T>();
                R nrm = R(0);
                if (nrows <= 32 * RG) {
                    T v[RG];
#pragma unroll
                    for (int u = 0; u < RG; ++u) { const int r = i + 1 + lane + 32 * u; v[u] = (r < p) ? col[r] : rc_zero<T>(); }
                    const T ci = col[i];
#pragma unroll
                    for (int u = 0; u < RG; ++u) { const int r = i + 1 + lane + 32 * u; if (r < p) part = rc_cfma(xs[r], v[u], part); }
                    part = rc_warp_sum(part);
                    const T f = ctau * (ci + part);
#pragma unroll
                    for (int u = 0; u < RG; ++u) {
                        const int r = i + 1 + lane + 32 * u;
                        if (r < p) {
                            const T nv = v[u] - f * xs[r];
                            col[r] = nv;
                            nrm += rc_real(nv) * rc_real(nv) + rc_imag(nv) * rc_imag(nv);
                        }
                    }
                    nrm = rc_warp_sum(nrm);
                    if (lane == 0) { col[i] = ci - f; vn[li] = sqrt((double)nrm); }
                } else {
                    for (int r = i + 1 + lane; r < p; r += 32) part = rc_cfma(xs[r], col[r], part);
                    part = rc_warp_sum(part);
                    const T ci = col[i];
                    const T f = ctau * (ci + part);
                    for (int r = i + 1 + lane; r < p; r += 32) {
                        const T nv = col[r] - f * xs[r];
                        col[r] = nv;
                        nrm += rc_real(nv) * rc_real(nv) + rc_imag(nv) * rc_imag(nv);
                    }
                    nrm = rc_warp_sum(nrm);
                    if (lane == 0) { col[i] = ci - f; vn[li] = sqrt((double)nrm); }
                }
            }
        }
        __syncthreads();   // xs is rewritten next step; lpos/vn written by lane 0 are read by the block
        RC_PHASE(4);      // trailing update
    }
#undef RC_PHASE
    // final logical order for the never-pivoted columns (n > p)
    for (int li = tid; li < nl; li += NT) {
        int lp = lpos[li];
        if (lp >= kk) ind[lp] = li * G + b;
    }
    // resident columns back to global memory (gather_r_kernel reads the factored W)
    for (int li = warp; li < min(nl, cap); li += NW) {
        T* dst = W + (int64_t)(li * G + b) * ldw;
        const T* src = scol + (size_t)li * p;
        for (int r = lane; r < p; r += 32) dst[r] = src[r];
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Medium factors (the w x w triangle of a wide sketch: 138 x 138 c64 at config 5, 266 x 266 in double at config 4) do
// not fit one CTA's shared memory but fit a CLUSTER's: up to 8 CTAs, every column resident in the shared memory of its
// owner, candidates exchanged through distributed shared memory (st.shared::cluster into every peer's slot table) and
// ONE hardware cluster barrier per step instead of a grid-wide sync through L2; the winning column is read straight out
// of its owner's shared memory.  Same pivot rule, same arithmetic as pivqr_kernel (the cooperative kernel spent ~18 000
// cycles per step on these shapes, 14 000 of them in L2 round trips: publish, grid sync, slot reads, staged column).
template <class T, int NT, int LPC>
__global__ void __launch_bounds__(NT)
pivqr_cluster_kernel(T* __restrict__ W, int64_t ldw, int p, int n, int kk, int nlmax,
                     int* __restrict__ ind, T* __restrict__ vbuf, T* __restrict__ tau_out, T* __restrict__ diag) {
    constexpr int NW = NT / 32;
    constexpr int MAXG = 8;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);                  // pivot column / reflector, p entries
    T* scol = xs + p;                                        // nlmax resident columns, p entries each
    double* vn = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(scol + (size_t)nlmax * p) + 7) & ~(uintptr_t)7);
    int* lpos = reinterpret_cast<int*>(vn + nlmax);
    __shared__ Cand c_slots[2][MAXG];                        // written by every CTA of the cluster (DSMEM)
    __shared__ int c_disp[2][MAXG];
    __shared__ Cand s_cand[NW];
    __shared__ int s_disp[NW];
    __shared__ double s_red[NW];
    __shared__ T s_hs[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int)cluster.num_blocks(), b = (int)cluster.block_rank();
    const int nl = (b < n) ? (n - b + G - 1) / G : 0;        // columns owned by this CTA: c = li * G + b

    for (int li = warp; li < nl; li += NW) {
        const T* src = W + (int64_t)(li * G + b) * ldw;
        T* dst = scol + (size_t)li * p;
        double a = 0.0;
        for (int r = lane; r < p; r += 32) { T v = src[r]; dst[r] = v; a += rc_abs2(v); }
        a = rc_warp_sum(a);
        if (lane == 0) { vn[li] = sqrt(a); lpos[li] = li * G + b; }
    }
    __syncthreads();

    for (int i = 0; i < kk; ++i) {
        // (a) local candidates over owned, not yet pivoted columns
        Cand best; best.val = -1.0; best.lpos = 0x7fffffff; best.phys = -1;
        int disp = -1;
        for (int li = tid; li < nl; li += NT) {
            int lp = lpos[li];
            if (lp >= i) {
                Cand cnd; cnd.val = vn[li]; cnd.lpos = lp; cnd.phys = li * G + b;
                if (better(cnd, best)) best = cnd;
                if (lp == i) disp = li * G + b;
            }
        }
        best = warp_best(best);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) disp = max(disp, __shfl_xor_sync(0xffffffffu, disp, m));
        if (lane == 0) { s_cand[warp] = best; s_disp[warp] = disp; }
        __syncthreads();
        if (warp == 0) {
            Cand bb; bb.val = -1.0; bb.lpos = 0x7fffffff; bb.phys = -1;
            int d = -1;
            if (lane < NW) { bb = s_cand[lane]; d = s_disp[lane]; }
            bb = warp_best(bb);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, m));
            // lane g publishes this CTA's candidate into CTA g's slot table
            if (lane < G) {
                Cand* rs = cluster.map_shared_rank(&c_slots[i & 1][b], lane);
                int* rd = cluster.map_shared_rank(&c_disp[i & 1][b], lane);
                *rs = bb;
                *rd = d;
            }
        }
        // (b) one cluster barrier per step (release / acquire: the slot writes above are visible behind it)
        cluster.sync();
        // (c) winner: every thread reduces the G slots of its own table
        Cand win; win.val = -1.0; win.lpos = 0x7fffffff; win.phys = -1;
        int dc = -1;
        for (int g = 0; g < G; ++g) {
            const Cand cg_ = c_slots[i & 1][g];
            if (better(cg_, win)) win = cg_;
            dc = max(dc, c_disp[i & 1][g]);
        }
        const int pv = win.phys >= 0 ? win.phys : dc;          // (no winner: every norm is NaN, speculative run on garbage)
        const int pv_lpos = win.phys >= 0 ? win.lpos : i;
        if (tid == 0) {
            if (dc >= 0 && dc != pv && (dc % G) == b) lpos[dc / G] = pv_lpos;
            if ((pv % G) == b) lpos[pv / G] = i;
        }
        if (b == 0 && tid == 0) ind[i] = pv;
        // (d) reflector from the pivot column, read out of its owner's shared memory (it is never written again)
        const T* pcol = cluster.map_shared_rank(scol + (size_t)(pv / G) * p, pv % G);
        double a = 0.0;
        for (int r = i + tid; r < p; r += NT) {
            T v = pcol[r];
            xs[r] = v;
            if (r > i) a += rc_abs2(v);
        }
        a = rc_warp_sum(a);
        if (lane == 0) s_red[warp] = a;
        __syncthreads();
        if (warp == 0) {
            double xnorm2 = (lane < NW) ? s_red[lane] : 0.0;
            xnorm2 = rc_warp_sum(xnorm2);
            T t0, t1, t2;
            larfg_dev<T>(xs[i], xnorm2, t0, t1, t2);
            if (lane == 0) { s_hs[0] = t0; s_hs[1] = t1; s_hs[2] = t2; }
        }
        __syncthreads();
        const T tau = s_hs[0], scale = s_hs[1], beta = s_hs[2];
        for (int r = i + 1 + tid; r < p; r += NT) xs[r] = xs[r] * scale;
        __syncthreads();
        if (b == 0) {
            T* vcol = vbuf + (int64_t)i * p;
            for (int r = tid; r < p; r += NT) vcol[r] = (r < i) ? rc_zero<T>() : (r == i ? rc_one<T>() : xs[r]);
            if (tid == 0) { tau_out[i] = tau; diag[i] = beta; }
        }
        // (e) trailing update of the owned columns + exact partial norms: LPC lanes per column
        const T ctau = rc_conj(tau);
        {
            constexpr int CPW = 32 / LPC;
            const int sl = lane % LPC, sc = lane / LPC;
            using R = RealOf<T>;
            for (int q0 = 0; warp + q0 * NW < nl; q0 += CPW) {
                const int li = warp + (q0 + sc) * NW;
                const bool act = (li < nl) && (lpos[min(li, nl - 1)] > i);
                T* col = scol + (size_t)min(li, nl - 1) * p;
                T part = rc_zero<T>();
                if (act)
                    for (int r = i + 1 + sl; r < p; r += LPC) part = rc_cfma(xs[r], col[r], part);
#pragma unroll
                for (int m = LPC / 2; m > 0; m >>= 1) part = part + rc_shfl_xor(part, m);
                T ci = rc_zero<T>(), f = rc_zero<T>();
                R nrm = R(0);
                if (act) {
                    ci = col[i];
                    f = ctau * (ci + part);
                    for (int r = i + 1 + sl; r < p; r += LPC) {
                        T v = col[r] - f * xs[r];
                        col[r] = v;
                        nrm += rc_real(v) * rc_real(v) + rc_imag(v) * rc_imag(v);
                    }
                }
#pragma unroll
                for (int m = LPC / 2; m > 0; m >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, m);
                if (act && sl == 0) { col[i] = ci - f; vn[li] = sqrt((double)nrm); }
            }
        }
        __syncthreads();   // xs is rewritten next step; lpos / vn written by single lanes are read by the block
    }
    for (int li = tid; li < nl; li += NT) {
        int lp = lpos[li];
        if (lp >= kk) ind[lp] = li * G + b;
    }
    for (int li = warp; li < nl; li += NW) {
        T* dst = W + (int64_t)(li * G + b) * ldw;
        const T* src = scol + (size_t)li * p;
        for (int r = lane; r < p; r += 32) dst[r] = src[r];
    }
    cluster.sync();        // no CTA may exit while a peer can still read its shared memory
}

// Small factors (the l x l R of a sketch): one CTA, everything in shared memory, and the serial part
// of a step (argmax -> pivot column -> ?larfg scalars) done by ONE warp, so a step costs two block
// barriers instead of seven.  Same pivot rule and numerics as pivqr_kernel.
template <class T, int NT, int LPC>
__global__ void __launch_bounds__(NT)
pivqr_small_kernel(T* __restrict__ Wg, int64_t ldwg, int p, int n, int kk, int* __restrict__ ind,
                   T* __restrict__ vbuf, T* __restrict__ tau_out, T* __restrict__ diag) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);                 // reflector, p entries
    T* W = xs + p;                                          // p x n column-major
    double* vn = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(W + (size_t)p * n) + 7) & ~(uintptr_t)7);
    int* lpos = reinterpret_cast<int*>(vn + n);
    __shared__ T s_tau;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < p * n; e += NT) { int c = e / p, r = e - c * p; W[e] = Wg[(int64_t)c * ldwg + r]; }
    __syncthreads();
    for (int c = warp; c < n; c += NW) {
        const T* col = W + (size_t)c * p;
        double a = 0.0;
        for (int r = lane; r < p; r += 32) a += rc_abs2(col[r]);
        a = rc_warp_sum(a);
        if (lane == 0) { vn[c] = sqrt(a); lpos[c] = c; }
    }
    __syncthreads();
    for (int i = 0; i < kk; ++i) {
        if (warp == 0) {
            Cand best; best.val = -1.0; best.lpos = 0x7fffffff; best.phys = -1;
            int disp = -1;
            for (int c = lane; c < n; c += 32) {
                int lp = lpos[c];
                if (lp >= i) {
                    Cand cnd; cnd.val = vn[c]; cnd.lpos = lp; cnd.phys = c;
                    if (better(cnd, best)) best = cnd;
                    if (lp == i) disp = c;
                }
            }
            best = warp_best(best);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) disp = max(disp, __shfl_xor_sync(0xffffffffu, disp, m));
            // (all candidate norms NaN -- a speculative run on the output of a Cholesky that broke down, rc_ctx::defer_depth
            // -- leaves no winner: take the column already sitting at position i, so every index stays in range)
            const int pv = best.phys >= 0 ? best.phys : disp;
            if (best.phys < 0) best.lpos = i;
            if (lane == 0) {
                if (disp >= 0 && disp != pv) lpos[disp] = best.lpos;
                lpos[pv] = i;
                ind[i] = pv;
            }
            const T* pcol = W + (size_t)pv * p;
            double a = 0.0;
            for (int r = i + lane; r < p; r += 32) { T v = pcol[r]; xs[r] = v; if (r > i) a += rc_abs2(v); }
            a = rc_warp_sum(a);
            __syncwarp();
            T tau, scale, beta;
            larfg_dev<T>(xs[i], a, tau, scale, beta);      // every lane: identical inputs, identical result
            __syncwarp();
            for (int r = i + 1 + lane; r < p; r += 32) xs[r] = xs[r] * scale;
            __syncwarp();
            T* vcol = vbuf + (int64_t)i * p;
            for (int r = lane; r < p; r += 32) vcol[r] = (r < i) ? rc_zero<T>() : (r == i ? rc_one<T>() : xs[r]);
            if (lane == 0) { tau_out[i] = tau; diag[i] = beta; s_tau = tau; }
        }
        __syncthreads();
        const T ctau = rc_conj(s_tau);
        // trailing update: LPC lanes per column, 32 / LPC columns per warp at a time (for the l x l factors of
        // the path, p <= 128, all columns are updated in ONE pass of the 32 warps with 3-step reductions)
        {
            constexpr int CPW = 32 / LPC;
            const int sl = lane % LPC, sc = lane / LPC;
            using A = typename AccOf<T>::type;
            for (int cb = warp * CPW; cb < n; cb += NW * CPW) {
                const int c = cb + sc;
                const bool act = (c < n) && (lpos[min(c, n - 1)] > i);
                T* col = W + (size_t)min(c, n - 1) * p;
                A part = rc_zero<A>();
                if (act)
                    for (int r = i + 1 + sl; r < p; r += LPC) part = rc_cfma(rc_widen(xs[r]), rc_widen(col[r]), part);
#pragma unroll
                for (int m = LPC / 2; m > 0; m >>= 1) part = part + rc_shfl_xor(part, m);
                T ci = rc_zero<T>(), f = rc_zero<T>();
                double nrm = 0.0;
                if (act) {
                    ci = col[i];
                    f = ctau * rc_narrow<T>(rc_widen(ci) + part);
                    for (int r = i + 1 + sl; r < p; r += LPC) {
                        T v = col[r] - f * xs[r];
                        col[r] = v;
                        nrm += rc_abs2(v);
                    }
                }
#pragma unroll
                for (int m = LPC / 2; m > 0; m >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, m);
                if (act && sl == 0) { col[i] = ci - f; vn[c] = sqrt(nrm); }
            }
        }
        __syncthreads();
    }
    for (int c = tid; c < n; c += NT) { int lp = lpos[c]; if (lp >= kk) ind[lp] = c; }
    __syncthreads();
    for (int e = tid; e < p * n; e += NT) { int c = e / p, r = e - c * p; Wg[(int64_t)c * ldwg + r] = W[e]; }
}

// ---------------------------------------------------------------------------------------------------------------
// Small factors, fused: pivoted QR + R + Q in ONE kernel (one CTA), for the l x l triangles of the tall sketches --
// the chain that sits on the critical path of every tall pivoted QR (it used to be transpose -> pivqr_small_kernel ->
// gather_r_kernel -> form_q_kernel -> transpose: ~285 us for 74 x 74 f64).  What changed against pivqr_small_kernel:
//   * ONE block barrier per step.  Every warp reduces the per-warp candidates itself and derives the reflector
//     scalars redundantly (the candidate carries the trailing norm of its column AND the norm below the diagonal, so
//     ?larfg needs no pass over the pivot column); there is no serial warp-0 phase and no second barrier.
//   * the logical position of a column is kept by the lane group that owns the column (registers), so the swap
//     bookkeeping of ?geqp3 needs no cross-thread traffic; tie-breaks are unchanged (first maximum in swap order).
//   * a lane's share of its column is loaded into registers once per step, updated there and stored once: a
//     read-modify-write loop over shared memory serialises on its own stores (possible aliasing), which is what made
//     the old update phase latency-bound.
//   * reflectors stay in shared memory; Q = H_0 ... H_{kk-1} I[:, :ncq] is formed by the same lane groups without a
//     single barrier (columns are independent), R and Q leave the kernel row-major in the caller's precision.
// Tin: storage type at the boundary; T: arithmetic type (double / c64 for f32 / c32 inputs when pivot decisions are
// taken in double, rc_ctx::pivot_f64).  LPC lanes per column; RPL rows per lane (p <= LPC * RPL).
struct __align__(16) CandF {
    double val;    // norm of rows i .. p-1 of the column (the pivot criterion)
    double xn2;    // squared norm of rows i+1 .. p-1 (?larfg's xnorm^2)
    int lpos, phys;
    int pad0, pad1;
};
__device__ __forceinline__ bool betterF(const CandF& a, const CandF& b) {
    if (a.phys < 0) return false;
    if (b.phys < 0) return true;
    if (a.val != b.val) return a.val > b.val;
    return a.lpos < b.lpos;
}
__device__ __forceinline__ CandF shflF(const CandF& c, int m) {
    CandF o;
    o.val = __shfl_xor_sync(0xffffffffu, c.val, m);
    o.xn2 = __shfl_xor_sync(0xffffffffu, c.xn2, m);
    o.lpos = __shfl_xor_sync(0xffffffffu, c.lpos, m);
    o.phys = __shfl_xor_sync(0xffffffffu, c.phys, m);
    return o;
}
// ?larfg without divisions or square roots on the critical path: one rsqrt and one reciprocal
template <class T>
__device__ __forceinline__ void larfg_fast(T alpha, double xnorm2, T& tau, T& scale, T& beta) {
    const double ar = (double)rc_real(alpha), ai = (double)rc_imag(alpha);
    if (xnorm2 == 0.0 && ai == 0.0) { tau = rc_zero<T>(); scale = rc_zero<T>(); beta = alpha; return; }
    const double s2 = ar * ar + ai * ai + xnorm2;
    const double rs = rsqrt(s2);
    const double b = -copysign(s2 * rs, ar), ib = -copysign(rs, ar);           // beta and 1 / beta
    const double dr = ar - b, di = ai;
    const double inv = __drcp_rn(dr * dr + di * di);
    tau = rc_make<T>((b - ar) * ib, -ai * ib);
    scale = rc_make<T>(dr * inv, -di * inv);
    beta = rc_make<T>(b, 0.0);
}
template <class D, class S> __device__ __forceinline__ D cast_scalar(S v) { return rc_make<D>((double)rc_real(v), (double)rc_imag(v)); }

constexpr int FQ_NT = 512;          // 128 registers per thread for the column tiles
constexpr int FQ_MAXP = 2;          // columns per lane group

template <class Tin, class T, int LPC, int RPL>
__global__ void __launch_bounds__(FQ_NT)
pivqr_fused_kernel(const Tin* __restrict__ in, int64_t ldin, int mode, int p, int n, int kk, int ncq,
                   int* __restrict__ ind, Tin* __restrict__ rout, int64_t ldr, Tin* __restrict__ qout, int64_t ldq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* W = reinterpret_cast<T*>(smem_raw);                 // p x n column-major (later: Q, p x ncq)
    T* V = W + (size_t)p * n;                              // p x kk reflectors, column-major
    T* s_tau = V + (size_t)p * kk;                         // kk
    T* s_diag = s_tau + kk;                                // kk
    int* s_pos2col = reinterpret_cast<int*>(s_diag + kk);  // n
    __shared__ CandF s_cand[2][FQ_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5, NG = NT / LPC;
    const int g = tid / LPC, sl = tid % LPC;
    // ---- load: mode 0 = `in` is p x n row-major, 1 = p x n column-major, 2 = `in` holds S (n x p row-major), M = S^H
    for (int e = tid; e < p * n; e += NT) {
        int rr, cc;
        T v;
        if (mode == 0) { rr = e / n; cc = e - rr * n; v = cast_scalar<T>(in[(int64_t)rr * ldin + cc]); }
        else if (mode == 1) { cc = e / p; rr = e - cc * p; v = cast_scalar<T>(in[(int64_t)cc * ldin + rr]); }
        else { cc = e / p; rr = e - cc * p; v = rc_conj(cast_scalar<T>(in[(int64_t)cc * ldin + rr])); }
        W[rr + (size_t)cc * p] = v;
    }
    __syncthreads();
    // ---- per-column state of this lane group: logical position (swap order of ?geqp3), pivoted flag, norms
    int lp[FQ_MAXP];
    bool alive[FQ_MAXP];
    double vn[FQ_MAXP], xn2[FQ_MAXP];
#pragma unroll
    for (int ps = 0; ps < FQ_MAXP; ++ps) {
        const int c = g + ps * NG;
        lp[ps] = c; alive[ps] = c < n; vn[ps] = 0.0; xn2[ps] = 0.0;
        double rest = 0.0, first = 0.0;
        if (c < n)
            for (int r = sl; r < p; r += LPC) { const double a2 = rc_abs2(W[r + (size_t)c * p]); if (r == 0) first = a2; else rest += a2; }
#pragma unroll
        for (int m = LPC / 2; m > 0; m >>= 1) { rest += __shfl_xor_sync(0xffffffffu, rest, m); first += __shfl_xor_sync(0xffffffffu, first, m); }
        vn[ps] = sqrt(rest + first); xn2[ps] = rest;
    }
    auto publish = [&](int buf) {
        CandF best; best.val = -1.0; best.xn2 = 0.0; best.lpos = 0x7fffffff; best.phys = -1; best.pad0 = best.pad1 = 0;
#pragma unroll
        for (int ps = 0; ps < FQ_MAXP; ++ps)
            if (alive[ps]) {
                CandF cd; cd.val = vn[ps]; cd.xn2 = xn2[ps]; cd.lpos = lp[ps]; cd.phys = g + ps * NG; cd.pad0 = cd.pad1 = 0;
                if (betterF(cd, best)) best = cd;
            }
#pragma unroll
        for (int m = LPC; m < 32; m <<= 1) { CandF o = shflF(best, m); if (betterF(o, best)) best = o; }
        if (lane == 0) s_cand[buf][warp] = best;
    };
    publish(0);
    __syncthreads();
    for (int i = 0; i < kk; ++i) {
        // ---- winner (every warp, redundantly)
        CandF win; win.val = -1.0; win.xn2 = 0.0; win.lpos = 0x7fffffff; win.phys = -1; win.pad0 = win.pad1 = 0;
        if (lane < NW) win = s_cand[i & 1][lane];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) { CandF o = shflF(win, m); if (betterF(o, win)) win = o; }
        const int pv = win.phys;          // < 0: every candidate norm is NaN (speculative run on garbage): no reflector
        T tau = rc_zero<T>(), scale = rc_zero<T>(), beta = rc_zero<T>();
        const T* pcol = W + (size_t)(pv >= 0 ? pv : 0) * p;
        if (pv >= 0) larfg_fast<T>(pcol[i], (i == p - 1) ? 0.0 : win.xn2, tau, scale, beta);
        if (warp == 0) {
            for (int r = lane; r < p; r += 32) V[r + (size_t)i * p] = (r < i) ? rc_zero<T>() : (r == i ? rc_one<T>() : pcol[r] * scale);
            if (lane == 0) { s_tau[i] = tau; s_diag[i] = beta; }
        }
        const T ctau = rc_conj(tau);
        // ---- own columns: bookkeeping of the swap, trailing update in registers, new norms
#pragma unroll
        for (int ps = 0; ps < FQ_MAXP; ++ps) {
            const int c = g + ps * NG;
            if (c < n) {
                if (pv < 0) { if (lp[ps] == i) alive[ps] = false; }
                else if (c == pv) { lp[ps] = i; alive[ps] = false; }
                else if (lp[ps] == i) lp[ps] = win.lpos;
            }
            const bool act = alive[ps] && pv >= 0;
            T* col = W + (size_t)(c < n ? c : 0) * p;
            // rows i+1 .. p-1 in strides of LPC: only the first `nu` tile slots are live at this step (warp-uniform)
            const int nu = (p - i - 1 + LPC - 1) / LPC;
            T tile[RPL], vt[RPL];
            T part = rc_zero<T>();
#pragma unroll
            for (int u = 0; u < RPL; ++u) {
                if (u < nu) {
                    const int r = i + 1 + sl + u * LPC;
                    const bool in = act && r < p;
                    tile[u] = in ? col[r] : rc_zero<T>();
                    vt[u] = in ? pcol[r] * scale : rc_zero<T>();
                    part = rc_cfma(vt[u], tile[u], part);
                }
            }
#pragma unroll
            for (int m = LPC / 2; m > 0; m >>= 1) part = part + rc_shfl_xor(part, m);
            const T ci = act ? col[i] : rc_zero<T>();
            const T f = ctau * (ci + part);
            double rest = 0.0, first = 0.0;
#pragma unroll
            for (int u = 0; u < RPL; ++u) {
                if (u < nu) {
                    const int r = i + 1 + sl + u * LPC;
                    tile[u] = tile[u] - f * vt[u];
                    const double a2 = rc_abs2(tile[u]);
                    if (u == 0 && sl == 0) first = a2; else rest += a2;
                    if (act && r < p) col[r] = tile[u];
                }
            }
#pragma unroll
            for (int m = LPC / 2; m > 0; m >>= 1) { rest += __shfl_xor_sync(0xffffffffu, rest, m); first += __shfl_xor_sync(0xffffffffu, first, m); }
            if (act) {
                if (sl == 0) col[i] = ci - f;
                vn[ps] = sqrt(rest + first); xn2[ps] = rest;
            }
        }
        publish((i + 1) & 1);
        __syncthreads();                                   // the one barrier of the step
    }
    // ---- pivot vector and the position -> column map
#pragma unroll
    for (int ps = 0; ps < FQ_MAXP; ++ps) {
        const int c = g + ps * NG;
        if (c < n && sl == 0) { const int l = min(max(lp[ps], 0), n - 1); ind[l] = c; s_pos2col[l] = c; }
    }
    __syncthreads();
    // ---- R (kk x n, row-major): column j of R is the column that ended at logical position j
    for (int e = tid; e < kk * n; e += NT) {
        const int row = e / n, j = e - row * n;
        T v;
        if (row > j) v = rc_zero<T>();
        else if (row == j) v = s_diag[row];
        else v = W[row + (size_t)s_pos2col[j] * p];
        rout[(int64_t)row * ldr + j] = cast_scalar<Tin>(v);
    }
    if (ncq <= 0) return;
    __syncthreads();
    // ---- Q = H_0 ... H_{kk-1} I[:, :ncq], column by column in registers (no barriers: columns are independent)
#pragma unroll
    for (int ps = 0; ps < FQ_MAXP; ++ps) {
        const int c = g + ps * NG;
        const bool act = c < ncq;
        T tile[RPL];
#pragma unroll
        for (int u = 0; u < RPL; ++u) { const int r = sl + u * LPC; tile[u] = (act && r == c) ? rc_one<T>() : rc_zero<T>(); }
        // e_c is untouched by the reflectors j > c (zeros from row j on).  The trip count is the same for the whole warp
        // (the shuffles need every lane): the largest column of the warp's lane groups decides, the others idle.
        const int cmax = (warp * 32 + 31) / LPC + ps * NG;
        for (int j = min(cmax, kk - 1); j >= 0; --j) {
            const bool on = act && j <= c;
            const T* v = V + (size_t)j * p;
            const int u0 = j / LPC, u1 = (p + LPC - 1) / LPC;          // live tile slots: rows j .. p-1 (warp-uniform)
            T part = rc_zero<T>(), vr[RPL];
#pragma unroll
            for (int u = 0; u < RPL; ++u)
                if (u >= u0 && u < u1) {
                    const int r = sl + u * LPC;
                    vr[u] = (on && r >= j && r < p) ? v[r] : rc_zero<T>();
                    part = rc_cfma(vr[u], tile[u], part);
                }
#pragma unroll
            for (int m = LPC / 2; m > 0; m >>= 1) part = part + rc_shfl_xor(part, m);
            const T f = s_tau[j] * part;
#pragma unroll
            for (int u = 0; u < RPL; ++u)
                if (u >= u0 && u < u1) tile[u] = tile[u] - f * vr[u];
        }
#pragma unroll
        for (int u = 0; u < RPL; ++u) { const int r = sl + u * LPC; if (act && r < p) qout[(int64_t)r * ldq + c] = cast_scalar<Tin>(tile[u]); }
    }
}

// r (kk x n row-major) from the factored column-major W, logical order `ind`.
template <class T>
__global__ void gather_r_kernel(const T* __restrict__ W, int64_t ldw, int kk, int n, const int* __restrict__ ind,
                                const T* __restrict__ diag, T* __restrict__ r, int64_t ldr) {
    int64_t total = (int64_t)kk * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int row = (int)(e / n), j = (int)(e - (int64_t)row * n);
        T v;
        if (row > j) v = rc_zero<T>();
        else if (row == j) v = diag[row];
        else v = W[(int64_t)ind[j] * ldw + row];
        r[(int64_t)row * ldr + j] = v;
    }
}

// qc (p x nc column-major) = H_0 ... H_{kk-1} I[:, :nc] ; one warp per output column.
template <class T>
__global__ void form_q_kernel(const T* __restrict__ vbuf, const T* __restrict__ tau, int p, int kk, int nc,
                              T* __restrict__ qc) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nc) return;
    T* col = qc + (int64_t)c * p;
    for (int r = lane; r < p; r += 32) col[r] = (r == c) ? rc_one<T>() : rc_zero<T>();
    __syncwarp();
    // e_c is untouched by reflectors j > c (their v has zeros above j and e_c is zero from j on)
    int jstart = (c < kk - 1) ? c : kk - 1;
    for (int j = jstart; j >= 0; --j) {
        const T* v = vbuf + (int64_t)j * p;
        using A = typename AccOf<T>::type;
        A part = rc_zero<A>();
        for (int r = j + lane; r < p; r += 32) part = rc_cfma(rc_widen(v[r]), rc_widen(col[r]), part);
        part = rc_warp_sum(part);
        T f = tau[j] * rc_narrow<T>(part);
        for (int r = j + lane; r < p; r += 32) col[r] = col[r] - f * v[r];
        __syncwarp();
    }
}

}  // namespace

// Cluster route (pivqr_cluster_kernel): false when the factor does not fit 8 CTAs' shared memory or the launch is refused.
template <class T>
static bool pivqr_cluster_launch(rc_ctx* c, T* wc, int64_t ldw, int p, int n, int kk, int* ind, T* vbuf, T* tau, T* diag, size_t lim) {
    if (c->cluster_qr == 0) return false;
    constexpr int NT = 1024;
    int G = 8;
    while (G > 2 && n < 32 * (G / 2)) G /= 2;                 // at least ~32 columns per CTA
    const int nlmax = (n + G - 1) / G;
    const size_t smem = ((size_t)p + (size_t)nlmax * p) * sizeof(T) + (size_t)nlmax * (sizeof(double) + sizeof(int)) + 64;
    if (smem + 8192 > lim) return false;
    void (*kern)(T*, int64_t, int, int, int, int, int*, T*, T*, T*) =
        p <= 160 ? pivqr_cluster_kernel<T, NT, 8> : (p <= 320 ? pivqr_cluster_kernel<T, NT, 16> : pivqr_cluster_kernel<T, NT, 32>);
    RC_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (const void*)kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return false; }
    RC_CUDA(cudaLaunchKernelEx(&cfg, kern, wc, ldw, p, n, kk, nlmax, ind, vbuf, tau, diag));
    RC_COUNT_LAUNCH(c);
    return true;
}

template <class T>
static void pivqr_factor_native(rc_ctx* c, T* wc, int64_t ldw, int64_t p, int64_t n, T* r, int64_t ldr, int* ind,
                                T* vbuf, T* tau) {
    int kk = (int)std::min(p, n);
    RC_REQUIRE(p > 0 && n > 0, "pivoted_qr: empty matrix");
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    DevBuf<T> diag(c, (size_t)kk);
    int pi = (int)p, ni = (int)n;
    size_t smem_all = ((size_t)p + (size_t)p * n + 1) * sizeof(T) + (size_t)n * (sizeof(double) + sizeof(int)) + 16;
    if (smem_all + 8192 <= lim && n <= 2048) {
        // small factor: one CTA, matrix resident in shared memory
        constexpr int NTS = 1024;
        if (p <= 256) {
            RC_CUDA(cudaFuncSetAttribute(pivqr_small_kernel<T, NTS, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all));
            pivqr_small_kernel<T, NTS, 8><<<1, NTS, smem_all, c->stream>>>(wc, ldw, pi, ni, kk, ind, vbuf, tau, diag.p);
        } else {
            RC_CUDA(cudaFuncSetAttribute(pivqr_small_kernel<T, NTS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all));
            pivqr_small_kernel<T, NTS, 32><<<1, NTS, smem_all, c->stream>>>(wc, ldw, pi, ni, kk, ind, vbuf, tau, diag.p);
        }
        RC_CHECK_LAUNCH(c);
    } else if (pivqr_cluster_launch<T>(c, wc, ldw, pi, ni, kk, ind, vbuf, tau, diag.p, lim)) {
        // medium factor: a cluster of CTAs holds every column in (distributed) shared memory
    } else {
        constexpr int NT = 1024;
        constexpr int NW = NT / 32;
        constexpr int CUMAX = std::is_same<T, c64>::value ? 2 : 4;     // c64 x 4 columns per lane group spills at 64 registers
        RC_REQUIRE((size_t)p * sizeof(T) + 16384 <= lim, "pivoted_qr: %lld rows exceed the shared-memory column buffer", (long long)p);
        int64_t want = (n + NW * 2 - 1) / (NW * 2);       // at least ~2 columns per warp
        int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, c->sm_count));
        int nlmax = (int)((n + grid - 1) / grid);
        // resident columns per CTA: whatever fits next to the reflector and the norm / position tables
        size_t fixed = (size_t)p * sizeof(T) + (size_t)nlmax * (sizeof(double) + sizeof(int)) + 64;
        int cap = (int)std::min<int64_t>(nlmax, (int64_t)((lim - 6144 - fixed) / ((size_t)p * sizeof(T))));
        if (cap < 0) cap = 0;
        size_t smem = fixed + (size_t)cap * p * sizeof(T);
        // lanes per column: ~10-20 rows per lane; columns in flight per warp: enough for one pass over its columns
        const int lpc = p <= 160 ? 8 : (p <= 320 ? 16 : 32);
        const int per_warp = (nlmax + NW - 1) / NW;
        const bool wide_cu = per_warp > 32 / lpc;
        void* kernel = nullptr;
#define RC_PICK(L) kernel = wide_cu ? (void*)pivqr_kernel<T, NT, L, CUMAX> : (void*)pivqr_kernel<T, NT, L, 1>
        if (lpc == 8) { RC_PICK(8); } else if (lpc == 16) { RC_PICK(16); } else { RC_PICK(32); }
#undef RC_PICK
        RC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DevBuf<Cand> slots(c, (size_t)2 * grid);
        DevBuf<int> slots_disp(c, (size_t)2 * grid);
        DevBuf<T> stage(c, (size_t)2 * grid * p);
        DevBuf<long long> prof;
        long long* prof_p = nullptr;
        if (c->trace) { prof.alloc(c, 8); RC_CUDA(cudaMemsetAsync(prof.p, 0, 64, c->stream)); prof_p = prof.p; }
        void* args[] = {&wc, &ldw, &pi, &ni, &kk, &cap, &nlmax, &ind, &vbuf, &tau, &diag.p, &slots.p, &slots_disp.p, &stage.p, &prof_p};
        RC_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(NT), args, smem, c->stream));
        if (c->trace) {
            long long h[8];
            RC_CUDA(cudaMemcpyAsync(h, prof.p, 64, cudaMemcpyDeviceToHost, c->stream));
            RC_CUDA(cudaStreamSynchronize(c->stream));
            fprintf(stderr, "[rc trace]     pivqr %d x %d, %d CTAs, %d resident cols/CTA: cycles per step: candidates+publish %lld, grid sync %lld, "
                            "winner %lld, reflector %lld, update %lld\n", pi, ni, grid, cap, h[0] / kk, h[1] / kk, h[2] / kk, h[3] / kk, h[4] / kk);
        }
        RC_COUNT_LAUNCH(c);
    }
    int64_t total = (int64_t)kk * n;
    int nb = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    gather_r_kernel<T><<<nb, 256, 0, c->stream>>>(wc, ldw, kk, ni, ind, diag.p, r, ldr);
    RC_CHECK_LAUNCH(c);
}

// Single-precision inputs are factored in double (option "pivot_precision" = 1, the default): the f32 / c32 data is
// widened exactly, every pivot decision -- trailing updates included -- is taken with double-precision arithmetic,
// and R, the reflectors and tau are rounded back once at the end.  The pivot sequence is then the one ?geqp3 picks
// in double precision on the same single-precision input (sgeqp3's own downdated norms are only good to
// ~sqrt(eps_f32), so "bit-exact wherever the gap exceeds 1e-6" is only attainable against that sequence).
template <class T>
void pivqr_factor(rc_ctx* c, T* wc, int64_t ldw, int64_t p, int64_t n, T* r, int64_t ldr, int* ind,
                  T* vbuf, T* tau) {
    using W = typename AccOf<T>::type;
    if constexpr (!std::is_same<T, W>::value) {
        if (c->pivot_f64) {
            const int64_t kk = std::min(p, n);
            DevBuf<W> wc2(c, (size_t)p * n), r2(c, (size_t)kk * n), v2(c, (size_t)p * kk), tau2(c, (size_t)kk);
            k_cast<W, T>(c, wc2.p, p, wc, ldw, n, p);                  // column-major: n columns of p entries
            pivqr_factor_native<W>(c, wc2.p, p, p, n, r2.p, n, ind, v2.p, tau2.p);
            k_cast<T, W>(c, r, ldr, r2.p, n, kk, n);
            k_cast<T, W>(c, vbuf, p, v2.p, p, kk, p);
            k_cast<T, W>(c, tau, kk, tau2.p, kk, 1, kk);
            return;
        }
    }
    pivqr_factor_native<T>(c, wc, ldw, p, n, r, ldr, ind, vbuf, tau);
}

// Fused small-factor route (pivqr_fused_kernel): false when the shape does not fit it.
template <class Tin, class T>
static bool pivqr_fused_launch(rc_ctx* c, const Tin* in, int64_t ldin, int mode, int64_t p, int64_t n, int64_t ncq,
                               Tin* r, int64_t ldr, int* ind, Tin* q, int64_t ldq) {
    const int64_t kk = std::min(p, n);
    constexpr int lpc = 8;
    if (p > 8 * 20 || n > (FQ_NT / lpc) * FQ_MAXP || ncq > kk) return false;
    size_t smem = ((size_t)p * n + (size_t)p * kk + 2 * (size_t)kk) * sizeof(T) + (size_t)n * sizeof(int) + 64;
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    if (smem + 4096 > lim) return false;
    // threads: one lane group per column when that fits (multiple of a warp), else FQ_MAXP columns per group
    int64_t groups = n <= FQ_NT / lpc ? n : (n + FQ_MAXP - 1) / FQ_MAXP;
    int nt = (int)std::min<int64_t>(FQ_NT, (groups * lpc + 31) / 32 * 32);
    int pi = (int)p, ni = (int)n, kki = (int)kk, ncqi = (int)ncq;
    if (p <= 8 * 10) {
        RC_CUDA(cudaFuncSetAttribute(pivqr_fused_kernel<Tin, T, 8, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pivqr_fused_kernel<Tin, T, 8, 10><<<1, nt, smem, c->stream>>>(in, ldin, mode, pi, ni, kki, ncqi, ind, r, ldr, q, ldq);
    } else {
        RC_CUDA(cudaFuncSetAttribute(pivqr_fused_kernel<Tin, T, 8, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pivqr_fused_kernel<Tin, T, 8, 20><<<1, nt, smem, c->stream>>>(in, ldin, mode, pi, ni, kki, ncqi, ind, r, ldr, q, ldq);
    }
    RC_CHECK_LAUNCH(c);
    return true;
}
// in: mode 0 = p x n row-major (ldin), 1 = p x n column-major, 2 = S (n x p row-major) with M = S^H.
// r: kk x n row-major; q: p x ncq row-major; ind: n.  Returns false when the shape needs the general route.
template <class T>
bool pivqr_fused(rc_ctx* c, const T* in, int64_t ldin, int mode, int64_t p, int64_t n, int64_t ncq,
                 T* r, int64_t ldr, int* ind, T* q, int64_t ldq) {
    if (c->fused_small_qr == 0) return false;        // option "fused_small_qr" = 0: the unfused kernels (A/B, tests)
    using W = typename AccOf<T>::type;
    if constexpr (!std::is_same<T, W>::value) {
        if (c->pivot_f64) return pivqr_fused_launch<T, W>(c, in, ldin, mode, p, n, ncq, r, ldr, ind, q, ldq);
    }
    return pivqr_fused_launch<T, T>(c, in, ldin, mode, p, n, ncq, r, ldr, ind, q, ldq);
}

// Compact-WY triangle of a block of nb reflectors (?larft, forward / columnwise) from their Gram matrix
// S = V^H V (nb x nb row-major):  T[i][i] = tau_i,  T[0:i, i] = -tau_i T[0:i, 0:i] S[0:i, i].  One CTA, thread a owns row a.
template <class T>
__global__ void larft_kernel(const T* __restrict__ S, int64_t lds, const T* __restrict__ tau, int nb, T* __restrict__ Tout, int64_t ldt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Ts = reinterpret_cast<T*>(smem_raw);            // nb x nb
    const int a = threadIdx.x;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) Ts[e] = rc_zero<T>();
    __syncthreads();
    for (int i = 0; i < nb; ++i) {
        const T ti = tau[i];
        if (a < i) {
            T acc = rc_zero<T>();
            for (int b = a; b < i; ++b) acc = acc + Ts[a * nb + b] * S[(int64_t)b * lds + i];
            Ts[a * nb + i] = rc_zero<T>() - ti * acc;
        } else if (a == i) {
            Ts[a * nb + i] = ti;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) Tout[(int64_t)(e / nb) * ldt + (e % nb)] = Ts[e];
}

// Large factors: Q = H_0 ... H_{kk-1} I[:, :nc] by blocks of FQ_NB reflectors in compact-WY form,
// Q <- (I - V_b T_b V_b^H) Q from the last block to the first, all the work in GEMMs (the one-warp-per-column kernel
// re-reads every reflector for every column: p^2 nc words of L2 traffic).  Block b = reflectors j0 .. j1-1 only touches
// rows >= j0 and columns >= j0 of Q (e_c is invariant under the reflectors j > c), so the GEMMs shrink with j0.
constexpr int FQ_NB = 64;
template <class T>
static void pivqr_form_q_blocked(rc_ctx* c, const T* vbuf, const T* tau, int64_t p, int64_t kk, int64_t nc, T* q, int64_t ldq) {
    // reflectors as a p x kk row-major matrix (vbuf is column-major p x kk = row-major kk x p)
    DevBuf<T> vr(c, (size_t)p * kk);
    k_transpose<T>(c, vr.p, kk, vbuf, p, kk, p, false);
    k_eye<T>(c, q, p, nc, ldq);
    DevBuf<T> s(c, (size_t)FQ_NB * FQ_NB), t(c, (size_t)FQ_NB * FQ_NB), w(c, (size_t)FQ_NB * nc), tw(c, (size_t)FQ_NB * nc);
    RC_CUDA(cudaFuncSetAttribute(larft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FQ_NB * FQ_NB * sizeof(T))));
    const int64_t nblk = (kk + FQ_NB - 1) / FQ_NB;
    for (int64_t b = nblk - 1; b >= 0; --b) {
        const int64_t j0 = b * FQ_NB, nb = std::min<int64_t>(FQ_NB, kk - j0);
        if (j0 >= nc) continue;                          // no column of I[:, :nc] is touched by this block
        const int64_t rows = p - j0, cols = nc - j0;
        const T* vb = vr.p + j0 * kk + j0;               // (p - j0) x nb, ld = kk
        T* qb = q + j0 * ldq + j0;                        // (p - j0) x (nc - j0)
        gemm<T>(c, RC_OP_H, RC_OP_N, nb, nb, rows, vb, kk, vb, kk, s.p, FQ_NB, rc_one<T>(), rc_zero<T>());
        larft_kernel<T><<<1, FQ_NB, FQ_NB * FQ_NB * sizeof(T), c->stream>>>(s.p, FQ_NB, tau + j0, (int)nb, t.p, FQ_NB);
        RC_CHECK_LAUNCH(c);
        gemm<T>(c, RC_OP_H, RC_OP_N, nb, cols, rows, vb, kk, qb, ldq, w.p, cols, rc_one<T>(), rc_zero<T>());
        gemm<T>(c, RC_OP_N, RC_OP_N, nb, cols, nb, t.p, FQ_NB, w.p, cols, tw.p, cols, rc_one<T>(), rc_zero<T>());
        gemm<T>(c, RC_OP_N, RC_OP_N, rows, cols, nb, vb, kk, tw.p, cols, qb, ldq, rc_zero<T>() - rc_one<T>(), rc_one<T>());
    }
}

template <class T>
void pivqr_form_q(rc_ctx* c, const T* vbuf, const T* tau, int64_t p, int64_t kk, int64_t nc, T* q, int64_t ldq) {
    if (nc == 0) return;
    if (p >= 512 && nc >= 128 && kk >= 128) { pivqr_form_q_blocked<T>(c, vbuf, tau, p, kk, nc, q, ldq); return; }
    DevBuf<T> qc(c, (size_t)p * nc);
    int warps = 4;
    int nb = (int)((nc + warps - 1) / warps);
    form_q_kernel<T><<<nb, warps * 32, 0, c->stream>>>(vbuf, tau, (int)p, (int)kk, (int)nc, qc.p);
    RC_CHECK_LAUNCH(c);
    // column-major p x nc  ==  row-major nc x p  -> transpose to row-major p x nc
    k_transpose<T>(c, q, ldq, qc.p, p, nc, p, false);
}

#define INST(T)                                                                                         \
    template bool pivqr_fused<T>(rc_ctx*, const T*, int64_t, int, int64_t, int64_t, int64_t, T*, int64_t, int*, T*, int64_t); \
    template void pivqr_factor<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, T*, int64_t, int*, T*, T*);    \
    template void pivqr_form_q<T>(rc_ctx*, const T*, const T*, int64_t, int64_t, int64_t, T*, int64_t);
INST(float)
INST(double)
INST(c32)
INST(c64)
