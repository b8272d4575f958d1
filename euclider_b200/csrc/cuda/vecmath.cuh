// Small fixed-size f64 vectors kept in registers.  Operation order follows nalgebra 0.8 (the
// reference's vector crate): component-wise loops, dot = left-to-right sum in index order,
// normalize = v / norm(v).  The translation unit is compiled with -fmad=false so no a*b+c is
// contracted (Rust/LLVM never does).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "eucl_detmath.h"

namespace eucl {

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
    double c[D];
    __device__ __forceinline__ double& operator[](int k) { return c[k]; }
    __device__ __forceinline__ double operator[](int k) const { return c[k]; }
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
__device__ __forceinline__ Vec<D> operator*(const Vec<D>& a, double s) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] * s;
    return r;
}
template <int D>
__device__ __forceinline__ Vec<D> operator/(const Vec<D>& a, double s) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = a[k] / s;
    return r;
}
template <int D>
__device__ __forceinline__ double dot(const Vec<D>& a, const Vec<D>& b) {
    double s = a[0] * b[0];
#pragma unroll
    for (int k = 1; k < D; ++k) s = s + a[k] * b[k];
    return s;
}
template <int D>
__device__ __forceinline__ double norm_squared(const Vec<D>& a) { return dot(a, a); }
template <int D>
__device__ __forceinline__ double norm(const Vec<D>& a) { return sqrt(norm_squared(a)); }
template <int D>
__device__ __forceinline__ Vec<D> normalize(const Vec<D>& a) { return a / norm(a); }

// Out-of-line copies of the deterministic libm (include/eucl_detmath.h): these bodies are 40-100
// instructions each and are called from many places; inlining them everywhere made the shade kernel
// 0.5 MB of SASS and instruction-fetch bound (profiles/r1_ncu_k_shade_baseline.txt: stall_no_instruction).
static __device__ __noinline__ double dm_acos(double x) { return eucl_det::det_acos(x); }
static __device__ __noinline__ double dm_asin(double x) { return eucl_det::det_asin(x); }
static __device__ __noinline__ double dm_sin(double x) { return eucl_det::det_sin(x); }
static __device__ __noinline__ double dm_cos(double x) { return eucl_det::det_cos(x); }
static __device__ __noinline__ double dm_atan(double x) { return eucl_det::det_atan(x); }
static __device__ __noinline__ double dm_atan2(double y, double x) { return eucl_det::det_atan2(y, x); }

// util.rs:712-722: acos of the normalised dot product, NaN -> 0
template <int D>
__device__ __forceinline__ double angle_cos(const Vec<D>& a, const Vec<D>& b) {
    return dot(a, b) / (norm(a) * norm(b));
}
__device__ __forceinline__ double angle_from_cos(double c) {
    double r = dm_acos(c); // deterministic libm shared with the oracle (include/eucl_detmath.h)
    return isnan(r) ? 0.0 : r;
}
template <int D>
__device__ __forceinline__ double angle_between(const Vec<D>& a, const Vec<D>& b) {
    return angle_from_cos(angle_cos(a, b));
}

// Rust f64::signum: +-1 by sign bit, NaN stays NaN
__device__ __forceinline__ double rust_signum(double x) {
    if (isnan(x)) return x;
    return signbit(x) ? -1.0 : 1.0;
}

template <int D>
__device__ __forceinline__ Vec<D> load_vec(const double* p, int stride) {
    Vec<D> r;
#pragma unroll
    for (int k = 0; k < D; ++k) r[k] = p[k * stride];
    return r;
}

} // namespace eucl
