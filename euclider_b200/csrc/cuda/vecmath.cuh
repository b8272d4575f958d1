// Small fixed-size f64 vectors kept in registers.  Operation order follows nalgebra 0.8 (the
// reference's vector crate): component-wise loops, dot = left-to-right sum in index order,
// normalize = v / norm(v).  The translation unit is compiled with -fmad=false so no a*b+c is
// contracted (Rust/LLVM never does).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "eucl_detmath.h"

// The scalar type of the trace loop: the reference's `type F` (src/main.rs:46-49) -- f64, or f32 with its
// `low_precision` feature.  The device code is compiled twice (EUCL_REAL = double in namespace eucl, float in
// namespace eucl_f32); scene tables, camera constants and LinearSpace expressions stay f64 in both builds (the
// reference parses JSON numbers as f64 and evaluates meval expressions in f64) and are narrowed where they are used.
#ifndef EUCL_REAL
#define EUCL_REAL double
#define EUCL_NS eucl
#define EUCL_REAL_IS_DOUBLE 1
#else
#define EUCL_REAL_IS_DOUBLE 0
#endif
#define R(x) ((EUCL_REAL)(x))

namespace EUCL_NS {

using real = EUCL_REAL;
constexpr bool kRealIsDouble = EUCL_REAL_IS_DOUBLE != 0;
constexpr real kMinNormal = kRealIsDouble ? (real)2.2250738585072014e-308 : (real)1.17549435e-38f; // smallest normal `real`

// sign bit / bit pattern of a real (the reference's signum and the memo keys compare these)
__device__ __forceinline__ bool sign_negative(double v) { return __double2hiint(v) < 0; }
__device__ __forceinline__ bool sign_negative(float v) { return __float_as_int(v) < 0; }
__device__ __forceinline__ long long bits_of(double v) { return __double_as_longlong(v); }
__device__ __forceinline__ long long bits_of(float v) { return (long long)__float_as_int(v); }

// how small vectors / colours are passed to the out-of-line helpers: by value (registers) or by reference (local memory)
#ifndef EUCL_BYVAL
#define EUCL_BYVAL 1
#endif
#if EUCL_BYVAL
#define EUCL_VARG(T) const T
#else
#define EUCL_VARG(T) const T&
#endif

template <int D>
struct Vec {
    real c[D];
    __device__ __forceinline__ real& operator[](int k) { return c[k]; }
    __device__ __forceinline__ real operator[](int k) const { return c[k]; }
};

template <int D>
__device__ __forceinline__ Vec<D> operator+(const Vec<D>& a, const Vec<D>& b) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] + b[k];
    return r;
}
template <int D>
__device__ __forceinline__ Vec<D> operator-(const Vec<D>& a, const Vec<D>& b) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] - b[k];
    return r;
}
template <int D>
__device__ __forceinline__ Vec<D> operator-(const Vec<D>& a) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = -a[k];
    return r;
}
template <int D>
__device__ __forceinline__ Vec<D> operator*(const Vec<D>& a, real s) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] * s;
    return r;
}
template <int D>
__device__ __forceinline__ Vec<D> operator/(const Vec<D>& a, real s) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] / s;
    return r;
}
template <int D>
__device__ __forceinline__ real dot(const Vec<D>& a, const Vec<D>& b) {
    real s = a[0] * b[0];
#pragma unroll
    for (int k = 1; k < D; ++k) s = s + a[k] * b[k];
    return s;
}
template <int D>
__device__ __forceinline__ real norm_squared(const Vec<D>& a) { return dot(a, a); }
template <int D>
__device__ __forceinline__ real norm(const Vec<D>& a) { return sqrt(norm_squared(a)); }
template <int D>
__device__ __forceinline__ Vec<D> normalize(const Vec<D>& a) { return a / norm(a); }

// Out-of-line copies of the deterministic libm (include/eucl_detmath.h): these bodies are 40-100
// instructions each and are called from many places; inlining them everywhere made the shade kernel
// 0.5 MB of SASS and instruction-fetch bound (profiles/r1_ncu_k_shade_baseline.txt: stall_no_instruction).
// BEGIN_KEEP64
// Both precisions evaluate them in f64 and narrow the result (f32 build: one extra rounding of an almost correctly
// rounded value -- what a good acosf returns in all but rare halfway cases; the f32 oracle does the same, so the two agree).
static __device__ __noinline__ real dm_acos(real x) { return (real)eucl_det::det_acos((double)x); }
static __device__ __noinline__ real dm_asin(real x) { return (real)eucl_det::det_asin((double)x); }
static __device__ __noinline__ real dm_sin(real x) { return (real)eucl_det::det_sin((double)x); }
static __device__ __noinline__ real dm_cos(real x) { return (real)eucl_det::det_cos((double)x); }
static __device__ __noinline__ real dm_atan(real x) { return (real)eucl_det::det_atan((double)x); }
static __device__ __noinline__ real dm_atan2(real y, real x) { return (real)eucl_det::det_atan2((double)y, (double)x); }
// END_KEEP64

// util.rs:712-722: acos of the normalised dot product, NaN -> 0
template <int D>
__device__ __forceinline__ real angle_cos(const Vec<D>& a, const Vec<D>& b) {
    return dot(a, b) / (norm(a) * norm(b));
}
__device__ __forceinline__ real angle_from_cos(real c) {
    real r = dm_acos(c); // deterministic libm shared with the oracle (include/eucl_detmath.h)
    return isnan(r) ? R(0.0) : r;
}
template <int D>
__device__ __forceinline__ real angle_between(const Vec<D>& a, const Vec<D>& b) {
    return angle_from_cos(angle_cos(a, b));
}

// Rust f64::signum: +-1 by sign bit, NaN stays NaN
__device__ __forceinline__ real rust_signum(real x) {
    if (isnan(x)) return x;
    return signbit(x) ? -R(1.0) : R(1.0);
}

// a vector from the (f64) scene tables  // KEEP64
template <int D>
__device__ __forceinline__ Vec<D> load_vec(const double* p, int stride) { // KEEP64
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = (real)p[k * stride];
    return r;
}

} // namespace EUCL_NS
