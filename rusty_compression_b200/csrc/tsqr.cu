// Tall-skinny Householder QR (TSQR) and application of its implicit Q.
//
// Replaces LAPACK ?geqp3's factorisation work and ?orgqr/?ungqr (reference N1/N2:
// src/pivoted_qr.rs:104-111, 139-173) for the tall sketches Y = A*Omega: an unpivoted
// Householder TSQR gives Y = Q0 R0 in one pass over Y; the column pivoting is then done on the
// small w x w factor R0 (pivqr.cu), which yields the same pivots as pivoting Y itself because
// Q0 is orthogonal (DESIGN.md "pivot parity").
//
// Each CTA owns a block of `block` consecutive rows held COLUMN-major in shared memory.
// Columns are processed in panels of NB: inside a panel the LAPACK ?larfg/?larf recurrences run
// reflector by reflector (norms accumulated in double, warp-shuffle reductions, one warp per
// panel column); the rest of the block is then updated with the panel's compact-WY block
// reflector I - V T V^H (?larft / ?larfb) as three small GEMM-shaped loops with one thread per
// output entry.  The w x w R factors of `g` consecutive blocks are stacked and factored again,
// level by level, until one remains; on those levels the loops only visit the rows that can be
// non-zero in a stack of upper triangles (TRI mode).  Q is applied top-down with the same block
// reflectors, panel by panel.
#include <type_traits>
#include "rc_internal.cuh"

namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int NB = 8;     // panel width

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

// Active rows of a panel starting at column c0 with nbp columns.
//  dense: rows c0 .. block-1                                   (row = c0 + a)
//  TRI  : the block is a stack of g upper triangles of size w: top rows c0 .. c0+nbp-1, then rows
//         k*w .. k*w + (c0+nbp-1) of every lower triangle k >= 1   (row = arows[a])
template <bool TRI>
__device__ __forceinline__ int build_rows(int* arows, int c0, int nbp, int w, int block) {
    if (!TRI) return block - c0;
    const int g = block / w;
    const int lim = min(c0 + nbp, w);          // rows 0 .. lim-1 of each lower triangle
    const int na = nbp + (g - 1) * lim;
    for (int a = threadIdx.x; a < na; a += NT) {
        int r;
        if (a < nbp) r = c0 + a;
        else { int b = a - nbp; int k = 1 + b / lim; r = k * w + (b - (k - 1) * lim); }
        arows[a] = r;
    }
    return na;
}
template <bool TRI>
__device__ __forceinline__ int row_of(const int* arows, int c0, int a) { return TRI ? arows[a] : c0 + a; }


// F(q, cc) = v_q^H X(:, cc) for the nbp reflectors of a panel (q < nbp <= 8) and ncols columns.
// One warp per column pair; lane = (row segment, q): q = lane & 7 walks the reflectors, seg =
// lane >> 3 takes every 4th active row, and two xor-shuffles combine the segments.
//   V: reflector columns (column c0+q at V + (c0+q)*sp), X: column cc at X + cc*spx.
template <class T, bool TRI>
__device__ __forceinline__ void panel_vhx(const T* __restrict__ V, int sp, const T* __restrict__ X, int spx, int ncols,
                                          const int* arows, int c0, int nbp, int na, T* __restrict__ F, int ldf, int NWARPS) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 7, seg = lane >> 3;
    const bool qok = q < nbp;
    const T* vq = V + (size_t)(c0 + (qok ? q : 0)) * sp;
    const int d = c0 + q;
    for (int cc = 2 * warp; cc < ncols; cc += 2 * NWARPS) {
        const bool two = (cc + 1 < ncols);
        const T* x0 = X + (size_t)cc * spx;
        const T* x1 = X + (size_t)(two ? cc + 1 : cc) * spx;
        using A = typename AccOf<T>::type;
        A a0 = rc_zero<A>(), a1 = rc_zero<A>();
        if (qok) {
            for (int a = seg; a < na; a += 4) {
                int r = TRI ? arows[a] : c0 + a;
                if (r > d) {
                    A v = rc_widen(vq[r]);
                    a0 = rc_cfma(v, rc_widen(x0[r]), a0);
                    a1 = rc_cfma(v, rc_widen(x1[r]), a1);
                }
            }
        }
        a0 = a0 + rc_shfl_xor(a0, 8);  a1 = a1 + rc_shfl_xor(a1, 8);
        a0 = a0 + rc_shfl_xor(a0, 16); a1 = a1 + rc_shfl_xor(a1, 16);
        if (seg == 0 && qok) {
            F[q * ldf + cc] = rc_narrow<T>(a0 + rc_widen(x0[d]));      // implicit unit diagonal of v_q
            if (two) F[q * ldf + cc + 1] = rc_narrow<T>(a1 + rc_widen(x1[d]));
        }
    }
}

// X(:, cc) -= V_p F(:, cc) on the active rows.  One warp per column, lanes over rows.
template <class T, bool TRI>
__device__ __forceinline__ void panel_update(const T* __restrict__ V, int sp, T* __restrict__ X, int spx, int ncols,
                                             const int* arows, int c0, int nbp, int na, const T* __restrict__ F, int ldf, int NWARPS) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int cc = warp; cc < ncols; cc += NWARPS) {
        T* x = X + (size_t)cc * spx;
        T fq[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) fq[q] = (q < nbp) ? F[q * ldf + cc] : rc_zero<T>();
        for (int a = lane; a < na; a += 32) {
            int r = TRI ? arows[a] : c0 + a;
            T acc = rc_zero<T>();
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                if (q < nbp) {
                    int d = c0 + q;
                    T vv = (r > d) ? V[(size_t)d * sp + r] : (r == d ? rc_one<T>() : rc_zero<T>());
                    acc = rc_fma(vv, fq[q], acc);
                }
            }
            x[r] = x[r] - acc;
        }
    }
}


// ---- FP64 tensor-pipe (DMMA.8x8x4) versions of the two panel loops, used for T = double.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// masked reflector entry: V(r, d) with the implicit unit diagonal and zeros above it
__device__ __forceinline__ double vmask(const double* __restrict__ V, int sp, int d, int r, bool ok) {
    if (!ok) return 0.0;
    return (r > d) ? V[(size_t)d * sp + r] : (r == d ? 1.0 : 0.0);
}

template <bool TRI>
__device__ __forceinline__ void panel_vhx_dmma(const double* __restrict__ V, int sp, const double* __restrict__ X, int spx, int ncols,
                                               const int* arows, int c0, int nbp, int na, double* __restrict__ F, int ldf, int NWARPS) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (ncols + 7) >> 3;
    for (int tile = warp; tile < ntile; tile += NWARPS) {
        const int cc = tile * 8 + g;                 // column of the B fragment held by this lane
        const bool cok = cc < ncols;
        const double* xc = X + (size_t)(cok ? cc : 0) * spx;
        double acc0 = 0.0, acc1 = 0.0, bcc0 = 0.0, bcc1 = 0.0;   // two accumulator pairs -> two DMMA chains
        int a0 = 0;
        for (; a0 + 8 <= na; a0 += 8) {
            int ra = TRI ? arows[a0 + t] : c0 + a0 + t;
            int rb = TRI ? arows[a0 + 4 + t] : c0 + a0 + 4 + t;
            double va = vmask(V, sp, c0 + g, ra, g < nbp), vb = vmask(V, sp, c0 + g, rb, g < nbp);
            double xa = cok ? xc[ra] : 0.0, xb = cok ? xc[rb] : 0.0;
            dmma884(acc0, acc1, va, xa);
            dmma884(bcc0, bcc1, vb, xb);
        }
        for (; a0 < na; a0 += 4) {
            int a = a0 + t;
            bool rok = a < na;
            int ra = rok ? (TRI ? arows[a] : c0 + a) : 0;
            double va = rok ? vmask(V, sp, c0 + g, ra, g < nbp) : 0.0;
            double xa = (rok && cok) ? xc[ra] : 0.0;
            dmma884(acc0, acc1, va, xa);
        }
        acc0 += bcc0; acc1 += bcc1;
        // C fragment: row g (= reflector q), columns tile*8 + 2t, +1
        const int oc = tile * 8 + 2 * t;
        if (g < nbp) {
            if (oc < ncols) F[g * ldf + oc] = acc0;
            if (oc + 1 < ncols) F[g * ldf + oc + 1] = acc1;
        }
    }
}

template <bool TRI>
__device__ __forceinline__ void panel_update_dmma(const double* __restrict__ V, int sp, double* __restrict__ X, int spx, int ncols,
                                                  const int* arows, int c0, int nbp, int na, const double* __restrict__ F, int ldf, int NWARPS) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile_c = (ncols + 7) >> 3, ntile_r = (na + 7) >> 3;
    for (int w2 = warp; w2 < ntile_c * ntile_r; w2 += NWARPS) {
        const int tc = w2 % ntile_c, trr = w2 / ntile_c;
        // B fragments: -F(q = t (+4), cc = tc*8 + g)
        const int cc = tc * 8 + g;
        const bool cok = cc < ncols;
        double b0 = (cok && t < nbp) ? -F[t * ldf + cc] : 0.0;
        double b1 = (cok && t + 4 < nbp) ? -F[(t + 4) * ldf + cc] : 0.0;
        // A fragments: V(r_g, q = t (+4))
        const int a = trr * 8 + g;
        const bool rok = a < na;
        const int r = rok ? (TRI ? arows[a] : c0 + a) : 0;
        double a0 = rok ? vmask(V, sp, c0 + t, r, t < nbp) : 0.0;
        double a1 = rok ? vmask(V, sp, c0 + t + 4, r, t + 4 < nbp) : 0.0;
        // C fragment: X(r_g, tc*8 + 2t, +1)
        const int oc = tc * 8 + 2 * t;
        double* x0 = X + (size_t)oc * spx + r;
        double* x1 = X + (size_t)(oc + 1) * spx + r;
        const bool s0 = rok && oc < ncols, s1 = rok && oc + 1 < ncols;
        double c0v = s0 ? *x0 : 0.0, c1v = s1 ? *x1 : 0.0;
        dmma884(c0v, c1v, a0, b0);
        dmma884(c0v, c1v, a1, b1);
        if (s0) *x0 = c0v;
        if (s1) *x1 = c1v;
    }
}

template <class T, bool TRI>
__device__ __forceinline__ void panel_vhx_any(const T* V, int sp, const T* X, int spx, int ncols, const int* arows, int c0, int nbp,
                                              int na, T* F, int ldf, int NWARPS) {
    if constexpr (std::is_same<T, double>::value) panel_vhx_dmma<TRI>(V, sp, X, spx, ncols, arows, c0, nbp, na, F, ldf, NWARPS);
    else panel_vhx<T, TRI>(V, sp, X, spx, ncols, arows, c0, nbp, na, F, ldf, NWARPS);
}
template <class T, bool TRI>
__device__ __forceinline__ void panel_update_any(const T* V, int sp, T* X, int spx, int ncols, const int* arows, int c0, int nbp,
                                                 int na, const T* F, int ldf, int NWARPS) {
    if constexpr (std::is_same<T, double>::value) panel_update_dmma<TRI>(V, sp, X, spx, ncols, arows, c0, nbp, na, F, ldf, NWARPS);
    else panel_update<T, TRI>(V, sp, X, spx, ncols, arows, c0, nbp, na, F, ldf, NWARPS);
}

// ---------------------------------------------------------------------------------- factor
// y: rows x w (row-major, ld); `block` rows per CTA (>= w); sp = padded column length in smem.
// In place: R on/above the diagonal of the block's first w rows, reflectors below.
// rstack: nblocks x (w x w) row-major upper-triangular copies; tau: nblocks x w;
// tpan: nblocks x npan x (NB x NB) compact-WY T factors (row-major).
template <class T, bool TRI>
__global__ void __launch_bounds__(NT)
house_block_qr_kernel(T* __restrict__ y, int64_t ld, int64_t rows, int w, int block, int sp,
                      T* __restrict__ rstack, T* __restrict__ tau_out, T* __restrict__ tpan_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);             // w columns of sp
    T* Wt = S + (size_t)w * sp;                         // NB x w
    T* Wt2 = Wt + (size_t)NB * w;                       // NB x w
    T* Tp = Wt2 + (size_t)NB * w;                       // NB x NB
    T* Gp = Tp + NB * NB;                               // NB x NB
    T* staus = Gp + NB * NB;                            // NB
    int* arows = reinterpret_cast<int*>(staus + NB);    // block ints (TRI only)
    __shared__ double red[NW];
    __shared__ HouseScalars<T> s_h[2];

    const int64_t r0 = (int64_t)blockIdx.x * block;
    const int nrows = (int)((rows - r0 < block) ? rows - r0 : block);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npan = (w + NB - 1) / NB;

    for (int e = tid; e < block * w; e += NT) {
        int r = e / w, c = e - r * w;
        S[(size_t)c * sp + r] = (r < nrows) ? y[(r0 + r) * ld + c] : rc_zero<T>();
    }
    __syncthreads();
    const int steps = (w < block) ? w : block;
    for (int c0 = 0; c0 < steps; c0 += NB) {
        const int nbp = min(NB, steps - c0);
        const int pan = c0 / NB;
        const int na = build_rows<TRI>(arows, c0, nbp, w, block);
        if (TRI) __syncthreads();
        {   // squared norm of column c0 below its diagonal; one thread derives the reflector scalars
            const T* col = S + (size_t)c0 * sp;
            double a2 = 0.0;
            for (int a = tid; a < na; a += NT) { int r = row_of<TRI>(arows, c0, a); if (r > c0) a2 += rc_abs2(col[r]); }
            a2 = block_sum(a2, red);
            if (tid == 0) s_h[0] = larfg<T>(col[c0], a2);
            __syncthreads();
        }
        // ---- (1) reflector-by-reflector inside the panel
        for (int jj = 0; jj < nbp; ++jj) {
            const int j = c0 + jj;
            T* colj = S + (size_t)j * sp;
            const HouseScalars<T> h = s_h[jj & 1];
            for (int a = tid; a < na; a += NT) { int r = row_of<TRI>(arows, c0, a); if (r > j) colj[r] = colj[r] * h.scale; }
            if (tid == 0) { staus[jj] = h.tau; tau_out[(int64_t)blockIdx.x * w + j] = h.tau; }
            __syncthreads();
            if (tid == 0) colj[j] = h.beta;      // nobody reads colj[j] again before the write-back
            const T ctau = rc_conj(h.tau);
            for (int c = j + 1 + warp; c < c0 + nbp; c += NW) {
                T* colc = S + (size_t)c * sp;
                using A = typename AccOf<T>::type;
                A part = rc_zero<A>();
                for (int a = lane; a < na; a += 32) { int r = row_of<TRI>(arows, c0, a); if (r > j) part = rc_cfma(rc_widen(colj[r]), rc_widen(colc[r]), part); }
                part = rc_warp_sum(part);
                T f = ctau * rc_narrow<T>(rc_widen(colc[j]) + part);
                double nrm = 0.0;
                for (int a = lane; a < na; a += 32) {
                    int r = row_of<TRI>(arows, c0, a);
                    if (r > j) {
                        T v = colc[r] - f * colj[r];
                        colc[r] = v;
                        if (c == j + 1 && r > j + 1) nrm += rc_abs2(v);
                    }
                }
                __syncwarp();
                if (lane == 0) colc[j] = colc[j] - f;
                if (c == j + 1) {        // this warp owns the next pivot column: derive its reflector scalars
                    nrm = rc_warp_sum(nrm);
                    if (lane == 0) s_h[(jj + 1) & 1] = larfg<T>(colc[j + 1], nrm);
                }
            }
            __syncthreads();
        }
        // ---- (2) compact-WY T of the panel (?larft): T(q,q) = tau_q, T(0:q,q) = -tau_q T(0:q,0:q) V(:,0:q)^H v_q
        for (int e = warp; e < nbp * nbp; e += NW) {
            int q1 = e / nbp, q2 = e - q1 * nbp;
            if (q1 >= q2) continue;
            const T* v1 = S + (size_t)(c0 + q1) * sp;
            const T* v2 = S + (size_t)(c0 + q2) * sp;
            using A = typename AccOf<T>::type;
            A part = rc_zero<A>();
            for (int a = lane; a < na; a += 32) { int r = row_of<TRI>(arows, c0, a); if (r > c0 + q2) part = rc_cfma(rc_widen(v1[r]), rc_widen(v2[r]), part); }
            part = rc_warp_sum(part);
            if (lane == 0) Gp[q1 * NB + q2] = rc_narrow<T>(part + rc_widen(rc_conj(v1[c0 + q2])));    // + conj(v1[c0+q2]) * 1
        }
        __syncthreads();
        if (warp == 0) {
            // lane i owns row i of T; columns are filled left to right (each needs the previous ones)
            const int i = lane;
            if (i < NB) for (int q = 0; q < NB; ++q) Tp[i * NB + q] = rc_zero<T>();
            __syncwarp();
            for (int q = 0; q < nbp; ++q) {
                if (i < q) {
                    T acc = rc_zero<T>();
                    for (int l2 = i; l2 < q; ++l2) acc = rc_fma(Tp[i * NB + l2], Gp[l2 * NB + q], acc);
                    Tp[i * NB + q] = -(staus[q] * acc);
                } else if (i == q) {
                    Tp[q * NB + q] = staus[q];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < NB * NB) tpan_out[((int64_t)blockIdx.x * npan + pan) * NB * NB + tid] = Tp[tid];
        // ---- (3) block update of the trailing columns: A -= V (T^H (V^H A))
        const int ct0 = c0 + nbp, ntr = w - ct0;
        if (ntr > 0) {
            T* Xtr = S + (size_t)ct0 * sp;
            panel_vhx_any<T, TRI>(S, sp, Xtr, sp, ntr, arows, c0, nbp, na, Wt, w, NW);
            __syncthreads();
            for (int e = tid; e < nbp * ntr; e += NT) {   // Wt2 = T^H Wt, one thread per entry
                int q = e / ntr, cc = e - q * ntr;
                T acc = rc_zero<T>();
                for (int q2 = 0; q2 <= q; ++q2) acc = rc_cfma(Tp[q2 * NB + q], Wt[q2 * w + cc], acc);
                Wt2[q * w + cc] = acc;
            }
            __syncthreads();
            panel_update_any<T, TRI>(S, sp, Xtr, sp, ntr, arows, c0, nbp, na, Wt2, w, NW);
        }
        __syncthreads();
    }
    // tau / T for the (identity) reflectors beyond `steps`
    for (int j = steps + tid; j < w; j += NT) tau_out[(int64_t)blockIdx.x * w + j] = rc_zero<T>();
    for (int pan = (steps + NB - 1) / NB; pan < npan; ++pan)
        if (tid < NB * NB) tpan_out[((int64_t)blockIdx.x * npan + pan) * NB * NB + tid] = rc_zero<T>();
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

// ---------------------------------------------------------------------------------- apply
// out block (block rows x ncl cols) = H_0 ... H_{w-1} [ctop_chunk ; 0], applied panel by panel with
// the compact-WY block reflectors, last panel first.  v: rows x w reflectors (row-major ldv) as
// left by the factor kernel; cin: chunk b is the w x nc matrix at rows [b*w, (b+1)*w) of cin
// (row-major ldc).  grid.y splits the nc columns into chunks of ncc.
template <class T, bool TRI>
__global__ void __launch_bounds__(NT)
house_block_apply_kernel(const T* __restrict__ v, int64_t ldv, int64_t rows, int w, int block, int sp,
                         const T* __restrict__ tpan, const T* __restrict__ cin, int64_t ldc, int nc, int ncc,
                         T* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Sv = reinterpret_cast<T*>(smem_raw);            // w columns of sp
    T* Sc = Sv + (size_t)w * sp;                       // ncc columns of sp
    T* F = Sc + (size_t)ncc * sp;                      // NB x ncc
    T* F2 = F + (size_t)NB * ncc;                      // NB x ncc
    T* Tp = F2 + (size_t)NB * ncc;                     // NB x NB
    int* arows = reinterpret_cast<int*>(Tp + NB * NB); // block ints (TRI only)
    const int64_t r0 = (int64_t)blockIdx.x * block;
    const int nrows = (int)((rows - r0 < block) ? rows - r0 : block);
    const int c0g = blockIdx.y * ncc;
    const int ncl = (nc - c0g < ncc) ? nc - c0g : ncc;
    const int tid = threadIdx.x;
    const int npan = (w + NB - 1) / NB;
    const int steps = (w < block) ? w : block;

    for (int e = tid; e < block * w; e += NT) {
        int r = e / w, c = e - r * w;
        Sv[(size_t)c * sp + r] = (r < nrows) ? v[(r0 + r) * ldv + c] : rc_zero<T>();
    }
    for (int e = tid; e < block * ncl; e += NT) {
        int r = e / ncl, c = e - r * ncl;
        Sc[(size_t)c * sp + r] = (r < w) ? cin[((int64_t)blockIdx.x * w + r) * ldc + c0g + c] : rc_zero<T>();
    }
    __syncthreads();
    for (int pan = (steps + NB - 1) / NB - 1; pan >= 0; --pan) {
        const int c0 = pan * NB;
        const int nbp = min(NB, steps - c0);
        const int na = build_rows<TRI>(arows, c0, nbp, w, block);
        if (tid < NB * NB) Tp[tid] = tpan[((int64_t)blockIdx.x * npan + pan) * NB * NB + tid];
        __syncthreads();
        // F = V_p^H X ; F <- T_p F ; X -= V_p F
        panel_vhx_any<T, TRI>(Sv, sp, Sc, sp, ncl, arows, c0, nbp, na, F, ncc, NW);
        __syncthreads();
        for (int e = tid; e < nbp * ncl; e += NT) {       // F2 = T_p F (upper triangular), one thread per entry
            int q = e / ncl, cc = e - q * ncl;
            T acc = rc_zero<T>();
            for (int q2 = q; q2 < nbp; ++q2) acc = rc_fma(Tp[q * NB + q2], F[q2 * ncc + cc], acc);
            F2[q * ncc + cc] = acc;
        }
        __syncthreads();
        panel_update_any<T, TRI>(Sv, sp, Sc, sp, ncl, arows, c0, nbp, na, F2, ncc, NW);
        __syncthreads();
    }
    for (int e = tid; e < nrows * ncl; e += NT) {
        int r = e / ncl, c = e - r * ncl;
        out[(r0 + r) * ldo + c0g + c] = Sc[(size_t)c * sp + r];
    }
}

inline size_t smem_budget(rc_ctx* c) {
    size_t lim = c->smem_optin ? c->smem_optin : (size_t)227 * 1024;
    return lim - 2048;   // static smem + slack
}

template <class T>
size_t factor_smem(int64_t w, int64_t sp, int64_t block) {
    return ((size_t)w * sp + (size_t)2 * NB * w + 2 * NB * NB + NB) * sizeof(T) + (size_t)block * sizeof(int) + 16;
}
template <class T>
size_t apply_smem(int64_t w, int64_t sp, int64_t block, int64_t ncc) {
    return ((size_t)(w + ncc) * sp + (size_t)2 * NB * ncc + NB * NB) * sizeof(T) + (size_t)block * sizeof(int) + 16;
}

// rows per CTA: leaf blocks sized for ~100 KB (two CTAs per SM); stacked-triangle levels may use
// a larger fan-in because they only touch the non-zero rows.
template <class T>
void plan_block(rc_ctx* c, int64_t w, int64_t rows, bool tri, int64_t& block, int64_t& sp) {
    size_t target = tri ? smem_budget(c) - 4096 : std::min<size_t>(smem_budget(c), (size_t)100 * 1024);
    int64_t bmax = (int64_t)(target / sizeof(T) / (w + 1)) - 2 * NB;
    if (bmax < 2 * w) bmax = 2 * w;
    int64_t g = bmax / w;
    if (g < 2) g = 2;
    if (tri && g > 4) g = 4;
    block = g * w;
    if (block > rows) block = tri ? (rows + w - 1) / w * w : std::max<int64_t>(rows, w);
    if (block < w) block = w;
    sp = block | 1;
    RC_REQUIRE(factor_smem<T>(w, sp, block) <= smem_budget(c), "tsqr: panel width %lld too large for shared memory", (long long)w);
}

}  // namespace

int64_t tsqr_max_width(rc_ctx* c, int dtype) {
    size_t sz = rc_dtype_size(dtype);
    size_t budget = smem_budget(c);
    int64_t w = 8;
    // need the factor of a 2w-row block and the apply with at least 8 right-hand-side columns
    for (;;) {
        int64_t w2 = w + 1, blk = 2 * w2, sp = blk | 1;
        size_t f = ((size_t)w2 * sp + (size_t)2 * NB * w2 + 2 * NB * NB + NB) * sz + blk * 4 + 16;
        size_t a = ((size_t)(w2 + 8) * sp + (size_t)2 * NB * 8 + NB * NB) * sz + blk * 4 + 16;
        if (f > budget || a > budget) break;
        w = w2;
    }
    return w;
}

template <class T>
TsqrFactor<T>::~TsqrFactor() {
    for (auto& L : levels) {
        if (L.owns_v && L.v) rc_dev_free(ctx, L.v);
        if (L.tau) rc_dev_free(ctx, L.tau);
        if (L.tpan) rc_dev_free(ctx, L.tpan);
    }
    if (r) rc_dev_free(ctx, r);
}

template <class T>
void tsqr_factor(rc_ctx* c, T* y, int64_t ld, int64_t m, int64_t w, TsqrFactor<T>& f) {
    RC_REQUIRE(m > 0 && w > 0, "tsqr: empty matrix");
    f.ctx = c;
    f.m = m;
    f.w = w;
    T* cur = y;
    int64_t cur_ld = ld, cur_rows = m;
    bool owns = false, tri = false;
    const int64_t npan = (w + NB - 1) / NB;
    for (;;) {
        int64_t block, sp;
        plan_block<T>(c, w, cur_rows, tri, block, sp);
        int64_t nblocks = (cur_rows + block - 1) / block;
        typename TsqrFactor<T>::Level L;
        L.v = cur; L.ldv = cur_ld; L.rows = cur_rows; L.block = block; L.nblocks = nblocks; L.owns_v = owns; L.tri = tri;
        L.tau = static_cast<T*>(rc_dev_alloc(c, sizeof(T) * nblocks * w));
        L.tpan = static_cast<T*>(rc_dev_alloc(c, sizeof(T) * nblocks * npan * NB * NB));
        T* rstack = nullptr;
        rstack = static_cast<T*>(rc_dev_alloc(c, sizeof(T) * nblocks * w * w));
        size_t smem = factor_smem<T>(w, sp, block);
        if (tri) {
            RC_CUDA(cudaFuncSetAttribute(house_block_qr_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            house_block_qr_kernel<T, true><<<(unsigned)nblocks, NT, smem, c->stream>>>(cur, cur_ld, cur_rows, (int)w, (int)block, (int)sp, rstack, L.tau, L.tpan);
        } else {
            RC_CUDA(cudaFuncSetAttribute(house_block_qr_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            house_block_qr_kernel<T, false><<<(unsigned)nblocks, NT, smem, c->stream>>>(cur, cur_ld, cur_rows, (int)w, (int)block, (int)sp, rstack, L.tau, L.tpan);
        }
        RC_CHECK_LAUNCH(c);
        f.levels.push_back(L);
        if (nblocks == 1) { f.r = rstack; break; }
        cur = rstack; cur_ld = w; cur_rows = nblocks * w; owns = true; tri = true;
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
        size_t budget = smem_budget(c);
        RC_REQUIRE(apply_smem<T>(w, sp, L.block, 1) <= budget, "tsqr_apply: panel too wide");
        // as many right-hand-side columns per CTA as shared memory allows ...
        int64_t ncc = nc;
        while (apply_smem<T>(w, sp, L.block, ncc) > budget) --ncc;
        // ... but keep enough CTAs to cover the SMs on small levels
        int64_t nchunks = (nc + ncc - 1) / ncc;
        if (L.nblocks * nchunks < c->sm_count && ncc > 16) {
            int64_t want = std::min<int64_t>((c->sm_count + L.nblocks - 1) / L.nblocks, (nc + 15) / 16);
            ncc = std::min(ncc, (nc + want - 1) / want);
            nchunks = (nc + ncc - 1) / ncc;
        }
        T* dst; int64_t dst_ld;
        if (l == 0) { dst = out; dst_ld = ldo; }
        else { tmp[flip].alloc(c, (size_t)L.rows * nc); dst = tmp[flip].p; dst_ld = nc; }
        size_t smem = apply_smem<T>(w, sp, L.block, ncc);
        dim3 grid((unsigned)L.nblocks, (unsigned)nchunks);
        if (L.tri) {
            RC_CUDA(cudaFuncSetAttribute(house_block_apply_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            house_block_apply_kernel<T, true><<<grid, NT, smem, c->stream>>>(L.v, L.ldv, L.rows, (int)w, (int)L.block, (int)sp, L.tpan,
                                                                           cin, cin_ld, (int)nc, (int)ncc, dst, dst_ld);
        } else {
            RC_CUDA(cudaFuncSetAttribute(house_block_apply_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            house_block_apply_kernel<T, false><<<grid, NT, smem, c->stream>>>(L.v, L.ldv, L.rows, (int)w, (int)L.block, (int)sp, L.tpan,
                                                                            cin, cin_ld, (int)nc, (int)ncc, dst, dst_ld);
        }
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
