// Bandwidth-bound utility kernels: fills, copies, (conj-)transposes, gathers (the crate's
// permutation.rs), triangular masks, norms (RelDiff / MaxColNorm), Philox Gaussian generator.
#include "rc_internal.cuh"
#include "philox.cuh"

namespace {

constexpr int TB = 256;

template <class T>
__global__ void fill_kernel(T* p, int64_t rows, int64_t cols, int64_t ld, T v) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        p[i * ld + j] = v;
    }
}

template <class T>
__global__ void eye_kernel(T* p, int64_t rows, int64_t cols, int64_t ld) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        p[i * ld + j] = (i == j) ? rc_one<T>() : rc_zero<T>();
    }
}

template <class T>
__global__ void copy_kernel(T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = src[i * lds + j];
    }
}

template <class T>
__global__ void strided_kernel(T* dst, int64_t ldd, const T* src, int64_t rs, int64_t cs, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = src[i * rs + j * cs];
    }
}

// precision change of a rows x cols block (f32 <-> f64, c32 <-> c64): pivot decisions on single-precision inputs
// are taken in double (DESIGN.md "pivot parity")
template <class D, class S>
__global__ void cast_kernel(D* dst, int64_t ldd, const S* src, int64_t lds, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        const S v = src[i * lds + j];
        dst[i * ldd + j] = rc_make<D>((double)rc_real(v), (double)rc_imag(v));
    }
}

// 32x32 smem-tiled transpose (coalesced both ways), optional conjugation.
template <class T>
__global__ void transpose_kernel(T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, bool conj) {
    __shared__ T tile[32][33];
    for (int64_t by = blockIdx.y; by * 32 < rows; by += gridDim.y) {
        int64_t r0 = by * 32, c0 = (int64_t)blockIdx.x * 32;
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int64_t i = r0 + r, j = c0 + threadIdx.x;
            if (i < rows && j < cols) tile[r][threadIdx.x] = src[i * lds + j];
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int64_t j = c0 + r, i = r0 + threadIdx.x;   // dst[j][i]
            if (i < rows && j < cols) {
                T v = tile[threadIdx.x][r];
                dst[j * ldd + i] = conj ? rc_conj(v) : v;
            }
        }
        __syncthreads();
    }
}

template <class T>
__global__ void conj_kernel(T* p, int64_t rows, int64_t cols, int64_t ld) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        p[i * ld + j] = rc_conj(p[i * ld + j]);
    }
}

template <class T>
__global__ void gather_cols_kernel(T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = src[i * lds + idx[j]];
    }
}

template <class T>
__global__ void gather_rows_kernel(T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = src[(int64_t)idx[i] * lds + j];
    }
}

template <class T>
__global__ void triu_kernel(T* p, int64_t rows, int64_t cols, int64_t ld) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        if (j < i) p[i * ld + j] = rc_zero<T>();
    }
}

template <class T>
__global__ void scale_rows_kernel(T* p, int64_t rows, int64_t cols, int64_t ld, const RealOf<T>* s) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        p[i * ld + j] = p[i * ld + j] * s[i];
    }
}

template <class T>
__global__ void sub_kernel(T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = a[i * lda + j] - b[i * ldb + j];
    }
}

template <class T>
__global__ void add_kernel(T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        dst[i * ldd + j] = a[i * lda + j] + b[i * ldb + j];
    }
}

template <class T> struct GaussStore;
template <> struct GaussStore<float> { static __device__ float make(double re, double) { return (float)re; } };
template <> struct GaussStore<double> { static __device__ double make(double re, double) { return re; } };
template <> struct GaussStore<c32> { static __device__ c32 make(double re, double im) { return c32((float)re, (float)im); } };
template <> struct GaussStore<c64> { static __device__ c64 make(double re, double im) { return c64(re, im); } };

template <class T>
__global__ void gaussian_kernel(T* p, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, uint32_t stream, int64_t row_offset) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        double re, im;
        rc_philox_gaussian_pair((uint64_t)((row_offset + i) * cols + j), seed, stream, re, im);
        p[i * ld + j] = GaussStore<T>::make(re, im);
    }
}

// BASELINE config 5 input: A_ij = exp(i kappa |x_i - y_j|) / |x_i - y_j| for points in two unit
// boxes a distance `shift` apart.  Point coordinate d of point i = the [0,1) uniform of Philox counter 3 i + d
// (stream 301 for x, 302 for y), so any row shard regenerates its own points (oracle/inputs.py mirrors it).
__global__ void helmholtz_points_kernel(double* pts, int64_t npts, int64_t first, uint64_t seed, uint32_t stream, double shift) {
    int64_t n = npts * 3;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / 3; int d = (int)(e - i * 3);
        uint32_t w[4];
        uint64_t ctr = (uint64_t)((first + i) * 3 + d);
        rc_philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        double u = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) * 1.1102230246251565404e-16;   // [0, 1)
        pts[e] = u + (d == 0 ? shift : 0.0);
    }
}
template <class T>
__global__ void helmholtz_fill_kernel(T* a, int64_t rows, int64_t cols, int64_t ld, const double* __restrict__ x,
                                      const double* __restrict__ y, double kappa) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        double dx = x[3 * i] - y[3 * j], dy = x[3 * i + 1] - y[3 * j + 1], dz = x[3 * i + 2] - y[3 * j + 2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        double sn, cs;
        sincos(kappa * d, &sn, &cs);
        a[i * ld + j] = GaussStore<T>::make(cs / d, sn / d);
    }
}

// Column norms^2: grid.x over 32-column strips, grid.y over row chunks; each warp-row of the
// block walks rows, lanes map to consecutive columns (coalesced); partials via atomicAdd(double).
template <class T>
__global__ void col_norms2_kernel(const T* a, int64_t lda, int64_t rows, int64_t cols, double* out) {
    __shared__ double part[8][33];
    int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (j < cols) {
        for (int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; i < rows; i += (int64_t)gridDim.y * blockDim.y)
            acc += rc_abs2(a[i * lda + j]);
    }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < cols) {
        double s = 0.0;
        for (int r = 0; r < (int)blockDim.y; ++r) s += part[r][threadIdx.x];
        atomicAdd(out + j, s);
    }
}

template <class T, bool DIFF>
__global__ void fro2_kernel(const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols, double* out) {
    __shared__ double part[TB / 32];
    int64_t n = rows * cols;
    double acc = 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / cols, j = e - i * cols;
        T v = a[i * lda + j];
        if (DIFF) v = v - b[i * ldb + j];
        acc += rc_abs2(v);
    }
    acc = rc_warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < TB / 32; ++w) s += part[w];
        atomicAdd(out, s);
    }
}

// X (rows x cols complex, interleaved) -> real (2*rows) x (2*cols):
//   row 2k   = X_k viewed as interleaved reals  (re, im, re, im, ...)
//   row 2k+1 = i * X_k                           (-im, re, -im, re, ...)
// so that [A viewed as real m x 2n] * dst = (A X) viewed as real m x 2 cols.
__global__ void expand_rhs_c64_kernel(double* dst, int64_t ldd, const c64* x, int64_t ldx, int64_t rows, int64_t cols) {
    int64_t n = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = e / cols, j = e - k * cols;
        c64 v = x[k * ldx + j];
        double* r0 = dst + (2 * k) * ldd + 2 * j;
        double* r1 = dst + (2 * k + 1) * ldd + 2 * j;
        r0[0] = v.re; r0[1] = v.im;
        r1[0] = -v.im; r1[1] = v.re;
    }
}

template <class R>
__global__ void convert_real_kernel(R* dst, const double* src, int64_t n) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        dst[e] = (R)src[e];
}

inline int nblocks_for(int64_t n) {
    int64_t b = (n + TB - 1) / TB;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}

}  // namespace

template <class T> void k_fill(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld, T v) {
    if (rows * cols == 0) return;
    fill_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld, v);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_eye(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld) {
    if (rows * cols == 0) return;
    eye_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_copy(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    copy_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, lds, rows, cols);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_strided_to_dense(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t rs, int64_t cs, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    strided_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, rs, cs, rows, cols);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_transpose(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, bool conj) {
    if (rows * cols == 0) return;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)std::min<int64_t>((rows + 31) / 32, 65535));
    transpose_kernel<T><<<grid, dim3(32, 8), 0, c->stream>>>(dst, ldd, src, lds, rows, cols, conj);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_conj_inplace(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld) {
    if (!ScalarTraits<T>::is_complex || rows * cols == 0) return;
    conj_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_gather_cols(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx) {
    if (rows * cols == 0) return;
    gather_cols_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, lds, rows, cols, idx);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_gather_rows(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols, const int* idx) {
    if (rows * cols == 0) return;
    gather_rows_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, lds, rows, cols, idx);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_triu(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld) {
    if (rows * cols == 0) return;
    triu_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_scale_rows(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld, const RealOf<T>* s) {
    if (rows * cols == 0) return;
    scale_rows_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld, s);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_sub(rc_ctx* c, T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    sub_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, a, lda, b, ldb, rows, cols);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_add(rc_ctx* c, T* dst, int64_t ldd, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    add_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, a, lda, b, ldb, rows, cols);
    RC_CHECK_LAUNCH(c);
}
// dst[:, j] = factor * src[:, j] for the flagged columns j
template <class T>
__global__ void replace_flagged_columns_kernel(T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds,
                                               int64_t rows, int64_t cols, const int* __restrict__ flags, RealOf<T> factor) {
    const int64_t total = rows * cols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / cols, j = e - i * cols;
        if (flags[j]) dst[i * ldd + j] = src[i * lds + j] * factor;
    }
}
template <class T> void k_replace_flagged_columns(rc_ctx* c, T* dst, int64_t ldd, const T* src, int64_t lds, int64_t rows, int64_t cols,
                                                  const int* flags, double factor) {
    if (rows * cols == 0) return;
    replace_flagged_columns_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, lds, rows, cols, flags, (RealOf<T>)factor);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_gaussian(rc_ctx* c, T* p, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, uint32_t stream, int64_t row_offset) {
    if (rows * cols == 0) return;
    gaussian_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(p, rows, cols, ld, seed, stream, row_offset);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_helmholtz(rc_ctx* c, T* a, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, double kappa, double shift, int64_t row_offset) {
    if (rows * cols == 0) return;
    DevBuf<double> x(c, (size_t)rows * 3), y(c, (size_t)cols * 3);
    helmholtz_points_kernel<<<nblocks_for(rows * 3), TB, 0, c->stream>>>(x.p, rows, row_offset, seed, 301u, 0.0);
    RC_CHECK_LAUNCH(c);
    helmholtz_points_kernel<<<nblocks_for(cols * 3), TB, 0, c->stream>>>(y.p, cols, 0, seed, 302u, shift);
    RC_CHECK_LAUNCH(c);
    helmholtz_fill_kernel<T><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(a, rows, cols, ld, x.p, y.p, kappa);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_col_norms2(rc_ctx* c, const T* a, int64_t lda, int64_t rows, int64_t cols, double* out) {
    RC_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * cols, c->stream));
    if (rows * cols == 0) return;
    int64_t gy = std::min<int64_t>((rows + 8 * 64 - 1) / (8 * 64), 1024);
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)std::max<int64_t>(gy, 1));
    col_norms2_kernel<T><<<grid, dim3(32, 8), 0, c->stream>>>(a, lda, rows, cols, out);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_fro2(rc_ctx* c, const T* a, int64_t lda, int64_t rows, int64_t cols, double* out) {
    RC_CUDA(cudaMemsetAsync(out, 0, sizeof(double), c->stream));
    if (rows * cols == 0) return;
    fro2_kernel<T, false><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(a, lda, nullptr, 0, rows, cols, out);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_diff_fro2(rc_ctx* c, const T* a, int64_t lda, const T* b, int64_t ldb, int64_t rows, int64_t cols, double* out) {
    RC_CUDA(cudaMemsetAsync(out, 0, sizeof(double), c->stream));
    if (rows * cols == 0) return;
    fro2_kernel<T, true><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(a, lda, b, ldb, rows, cols, out);
    RC_CHECK_LAUNCH(c);
}
void k_expand_rhs_c64(rc_ctx* c, double* dst, int64_t ldd, const c64* x, int64_t ldx, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    expand_rhs_c64_kernel<<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, x, ldx, rows, cols);
    RC_CHECK_LAUNCH(c);
}
template <class T> void k_convert_real(rc_ctx* c, RealOf<T>* dst, const double* src, int64_t n) {
    if (n == 0) return;
    convert_real_kernel<RealOf<T>><<<nblocks_for(n), TB, 0, c->stream>>>(dst, src, n);
    RC_CHECK_LAUNCH(c);
}

template <class D, class S> void k_cast(rc_ctx* c, D* dst, int64_t ldd, const S* src, int64_t lds, int64_t rows, int64_t cols) {
    if (rows * cols == 0) return;
    cast_kernel<D, S><<<nblocks_for(rows * cols), TB, 0, c->stream>>>(dst, ldd, src, lds, rows, cols);
    RC_CHECK_LAUNCH(c);
}
template void k_cast<double, float>(rc_ctx*, double*, int64_t, const float*, int64_t, int64_t, int64_t);
template void k_cast<float, double>(rc_ctx*, float*, int64_t, const double*, int64_t, int64_t, int64_t);
template void k_cast<c64, c32>(rc_ctx*, c64*, int64_t, const c32*, int64_t, int64_t, int64_t);
template void k_cast<c32, c64>(rc_ctx*, c32*, int64_t, const c64*, int64_t, int64_t, int64_t);

#define INST(T)                                                                                          \
    template void k_fill<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, T);                                   \
    template void k_eye<T>(rc_ctx*, T*, int64_t, int64_t, int64_t);                                       \
    template void k_copy<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t);                   \
    template void k_strided_to_dense<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t, int64_t); \
    template void k_transpose<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t, bool);        \
    template void k_conj_inplace<T>(rc_ctx*, T*, int64_t, int64_t, int64_t);                              \
    template void k_gather_cols<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t, const int*); \
    template void k_gather_rows<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t, const int*); \
    template void k_triu<T>(rc_ctx*, T*, int64_t, int64_t, int64_t);                                      \
    template void k_scale_rows<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, const RealOf<T>*);              \
    template void k_sub<T>(rc_ctx*, T*, int64_t, const T*, int64_t, const T*, int64_t, int64_t, int64_t); \
    template void k_add<T>(rc_ctx*, T*, int64_t, const T*, int64_t, const T*, int64_t, int64_t, int64_t); \
    template void k_gaussian<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, uint64_t, uint32_t, int64_t);     \
    template void k_replace_flagged_columns<T>(rc_ctx*, T*, int64_t, const T*, int64_t, int64_t, int64_t, const int*, double); \
    template void k_helmholtz<T>(rc_ctx*, T*, int64_t, int64_t, int64_t, uint64_t, double, double, int64_t); \
    template void k_col_norms2<T>(rc_ctx*, const T*, int64_t, int64_t, int64_t, double*);                 \
    template void k_fro2<T>(rc_ctx*, const T*, int64_t, int64_t, int64_t, double*);                       \
    template void k_diff_fro2<T>(rc_ctx*, const T*, int64_t, const T*, int64_t, int64_t, int64_t, double*); \
    template void k_convert_real<T>(rc_ctx*, RealOf<T>*, const double*, int64_t);
INST(float)
INST(double)
INST(c32)
INST(c64)
// index vectors
template void k_copy<int>(rc_ctx*, int*, int64_t, const int*, int64_t, int64_t, int64_t);
