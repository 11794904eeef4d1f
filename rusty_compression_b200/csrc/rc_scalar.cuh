// Scalar layer for the four types the reference monomorphises over (f32, f64, c32, c64;
// reference: src/types.rs:9 `pub use ndarray_linalg::{c32, c64, Scalar}`).
// Complex values are interleaved (re, im) pairs exactly like num::Complex (#[repr(C)]).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#define RC_HD __host__ __device__ __forceinline__

template <class R>
struct __align__(sizeof(R) * 2) Cx {
    R re, im;
    RC_HD Cx() {}
    RC_HD Cx(R r) : re(r), im(R(0)) {}
    RC_HD Cx(R r, R i) : re(r), im(i) {}
};
using c32 = Cx<float>;
using c64 = Cx<double>;

template <class R> RC_HD Cx<R> operator+(Cx<R> a, Cx<R> b) { return Cx<R>(a.re + b.re, a.im + b.im); }
template <class R> RC_HD Cx<R> operator-(Cx<R> a, Cx<R> b) { return Cx<R>(a.re - b.re, a.im - b.im); }
template <class R> RC_HD Cx<R> operator-(Cx<R> a) { return Cx<R>(-a.re, -a.im); }
template <class R> RC_HD Cx<R> operator*(Cx<R> a, Cx<R> b) {
    return Cx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> RC_HD Cx<R> operator*(Cx<R> a, R b) { return Cx<R>(a.re * b, a.im * b); }
template <class R> RC_HD Cx<R> operator*(R b, Cx<R> a) { return Cx<R>(a.re * b, a.im * b); }
template <class R> RC_HD Cx<R> operator/(Cx<R> a, R b) { return Cx<R>(a.re / b, a.im / b); }
template <class R> RC_HD Cx<R> operator/(Cx<R> a, Cx<R> b) {
    // Smith's algorithm (what LAPACK's ?ladiv family guards against: overflow in |b|^2).
    if (fabs((double)b.re) >= fabs((double)b.im)) {
        R r = b.im / b.re, d = b.re + b.im * r;
        return Cx<R>((a.re + a.im * r) / d, (a.im - a.re * r) / d);
    } else {
        R r = b.re / b.im, d = b.re * r + b.im;
        return Cx<R>((a.re * r + a.im) / d, (a.im * r - a.re) / d);
    }
}
template <class R> RC_HD Cx<R>& operator+=(Cx<R>& a, Cx<R> b) { a.re += b.re; a.im += b.im; return a; }
template <class R> RC_HD Cx<R>& operator-=(Cx<R>& a, Cx<R> b) { a.re -= b.re; a.im -= b.im; return a; }
template <class R> RC_HD Cx<R>& operator*=(Cx<R>& a, Cx<R> b) { a = a * b; return a; }

template <class T> struct ScalarTraits;
template <> struct ScalarTraits<float> {
    using Real = float; static constexpr bool is_complex = false; static constexpr int code = 0;
};
template <> struct ScalarTraits<double> {
    using Real = double; static constexpr bool is_complex = false; static constexpr int code = 1;
};
template <> struct ScalarTraits<c32> {
    using Real = float; static constexpr bool is_complex = true; static constexpr int code = 2;
};
template <> struct ScalarTraits<c64> {
    using Real = double; static constexpr bool is_complex = true; static constexpr int code = 3;
};
template <class T> using RealOf = typename ScalarTraits<T>::Real;

RC_HD float  rc_conj(float a) { return a; }
RC_HD double rc_conj(double a) { return a; }
template <class R> RC_HD Cx<R> rc_conj(Cx<R> a) { return Cx<R>(a.re, -a.im); }

RC_HD float  rc_real(float a) { return a; }
RC_HD double rc_real(double a) { return a; }
template <class R> RC_HD R rc_real(Cx<R> a) { return a.re; }
RC_HD float  rc_imag(float) { return 0.f; }
RC_HD double rc_imag(double) { return 0.0; }
template <class R> RC_HD R rc_imag(Cx<R> a) { return a.im; }

// |a|^2 accumulated in double regardless of the storage type (pivot decisions and
// Householder norms are taken on these; see DESIGN.md "pivot parity").
RC_HD double rc_abs2(float a) { return (double)a * (double)a; }
RC_HD double rc_abs2(double a) { return a * a; }
template <class R> RC_HD double rc_abs2(Cx<R> a) {
    return (double)a.re * (double)a.re + (double)a.im * (double)a.im;
}
RC_HD double rc_abs(float a) { return fabs((double)a); }
RC_HD double rc_abs(double a) { return fabs(a); }
template <class R> RC_HD double rc_abs(Cx<R> a) { return hypot((double)a.re, (double)a.im); }

template <class T> RC_HD T rc_zero() { return T(RealOf<T>(0)); }
template <class T> RC_HD T rc_one() { return T(RealOf<T>(1)); }
template <class T> RC_HD T rc_from_real(RealOf<T> r) { return T(r); }

// build a scalar from double parts (imaginary part dropped for real types)
template <class T> struct MakeScalar;
template <> struct MakeScalar<float> { static RC_HD float make(double re, double) { return (float)re; } };
template <> struct MakeScalar<double> { static RC_HD double make(double re, double) { return re; } };
template <> struct MakeScalar<c32> { static RC_HD c32 make(double re, double im) { return c32((float)re, (float)im); } };
template <> struct MakeScalar<c64> { static RC_HD c64 make(double re, double im) { return c64(re, im); } };
template <class T> RC_HD T rc_make(double re, double im) { return MakeScalar<T>::make(re, im); }

// fused a*b + c
RC_HD float  rc_fma(float a, float b, float c) { return fmaf(a, b, c); }
RC_HD double rc_fma(double a, double b, double c) { return fma(a, b, c); }
template <class R> RC_HD Cx<R> rc_fma(Cx<R> a, Cx<R> b, Cx<R> c) {
    return Cx<R>(c.re + a.re * b.re - a.im * b.im, c.im + a.re * b.im + a.im * b.re);
}
// conj(a)*b + c
RC_HD float  rc_cfma(float a, float b, float c) { return fmaf(a, b, c); }
RC_HD double rc_cfma(double a, double b, double c) { return fma(a, b, c); }
template <class R> RC_HD Cx<R> rc_cfma(Cx<R> a, Cx<R> b, Cx<R> c) {
    return Cx<R>(c.re + a.re * b.re + a.im * b.im, c.im + a.re * b.im - a.im * b.re);
}

// Accumulation type for inner products: single-precision data is accumulated in double so that
// the Householder dots do not lose digits to the summation length.
template <class T> struct AccOf { using type = T; };
template <> struct AccOf<float> { using type = double; };
template <> struct AccOf<c32> { using type = c64; };
RC_HD double rc_widen(float a) { return (double)a; }
RC_HD double rc_widen(double a) { return a; }
RC_HD c64 rc_widen(c32 a) { return c64((double)a.re, (double)a.im); }
RC_HD c64 rc_widen(c64 a) { return a; }
template <class T> RC_HD T rc_narrow(typename AccOf<T>::type a) { return rc_make<T>((double)rc_real(a), (double)rc_imag(a)); }

#ifdef __CUDACC__
__device__ __forceinline__ float  rc_shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double rc_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class R> __device__ __forceinline__ Cx<R> rc_shfl_xor(Cx<R> v, int m) {
    return Cx<R>(__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.im, m));
}
__device__ __forceinline__ float  rc_shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double rc_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class R> __device__ __forceinline__ Cx<R> rc_shfl(Cx<R> v, int src) {
    return Cx<R>(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}
template <class T> __device__ __forceinline__ T rc_warp_sum(T v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = v + rc_shfl_xor(v, m);
    return v;
}
#endif
