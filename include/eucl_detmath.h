/*
 * eucl_detmath.h -- a small, deterministic double-precision libm (acos, asin, sin, cos, atan,
 * atan2) built ONLY from IEEE-754 +, -, *, / and sqrt, so that the same inputs give the same bits
 * on the host (g++ -ffp-contract=off) and on the device (nvcc -fmad=false).
 *
 * Why: the reference (Rust) calls the platform libm for these functions (std f64::acos etc.);
 * results differ between libms in the last ulp, and euclider's picture depends on last-ulp
 * effects (u8 truncation, alpha == 255 tests, CSG knife edges).  To compare the CUDA path with
 * the CPU oracle BIT FOR BIT both use this file.  The algorithms are the classic fdlibm ones
 * (Sun Microsystems' freely distributable libm, the basis of musl / FreeBSD msun and hence of
 * Rust's *-musl targets): rational minimax approximations with hi/lo constant splitting, error
 * below 1 ulp.  tests/test_detmath.py checks them against glibc.
 *
 * Usable from C++ and CUDA (all functions are EUCL_HD inline).
 *
 * The algorithms, argument reductions and coefficient tables of det_acos, det_asin, det_sin, det_cos, det_atan and
 * det_atan2 are those of fdlibm (e_acos.c, e_asin.c, k_sin.c, k_cos.c, e_rem_pio2.c, s_atan.c, e_atan2.c), which is
 * distributed under the following notice:
 *
 * ====================================================
 * Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
 *
 * Developed at SunSoft, a Sun Microsystems, Inc. business.
 * Permission to use, copy, modify, and distribute this
 * software is freely granted, provided that this notice
 * is preserved.
 * ====================================================
 */
#ifndef EUCL_DETMATH_H
#define EUCL_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EUCL_HD __host__ __device__ __forceinline__
#else
#define EUCL_HD inline
#endif

namespace eucl_det {

EUCL_HD uint64_t bits_of(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, sizeof u);
    return u;
#endif
}
EUCL_HD double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, sizeof x);
    return x;
#endif
}
EUCL_HD uint32_t hi_word_abs(double x) { return (uint32_t)(bits_of(x) >> 32) & 0x7fffffffu; }
EUCL_HD double clear_low_word(double x) { return from_bits(bits_of(x) & 0xffffffff00000000ull); }
EUCL_HD double det_fabs(double x) { return from_bits(bits_of(x) & 0x7fffffffffffffffull); }
EUCL_HD bool det_signbit(double x) { return (bits_of(x) >> 63) != 0; }
EUCL_HD double det_sqrt(double x) { return sqrt(x); } /* IEEE correctly rounded on host and device */

/* shared rational kernel of asin / acos on z in [0, 0.25] */
EUCL_HD double asin_ratio(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}

EUCL_HD double det_acos(double x) {
    const double pi = 3.14159265358979311600e+00, pio2_hi = 1.57079632679489655800e+00,
                 pio2_lo = 6.12323399573676603587e-17;
    const uint32_t ix = hi_word_abs(x);
    const bool neg = det_signbit(x);
    if (ix >= 0x3ff00000u) { /* |x| >= 1 or NaN */
        if (x == 1.0) return 0.0;
        if (x == -1.0) return pi + 2.0 * pio2_lo;
        return (x - x) / (x - x); /* NaN */
    }
    if (ix < 0x3fe00000u) { /* |x| < 0.5 */
        if (ix <= 0x3c600000u) return pio2_hi + pio2_lo;
        const double r = asin_ratio(x * x);
        return pio2_hi - (x - (pio2_lo - r * x));
    }
    if (neg) { /* x < -0.5 */
        const double z = (1.0 + x) * 0.5;
        const double s = det_sqrt(z);
        const double r = asin_ratio(z);
        const double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    { /* x > 0.5 */
        const double z = (1.0 - x) * 0.5;
        const double s = det_sqrt(z);
        const double df = clear_low_word(s);
        const double c = (z - df * df) / (s + df);
        const double r = asin_ratio(z);
        const double w = r * s + c;
        return 2.0 * (df + w);
    }
}

EUCL_HD double det_asin(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pio4_hi = 7.85398163397448278999e-01;
    const uint32_t ix = hi_word_abs(x);
    const bool neg = det_signbit(x);
    if (ix >= 0x3ff00000u) { /* |x| >= 1 or NaN */
        if (x == 1.0 || x == -1.0) return x * pio2_hi + x * pio2_lo;
        return (x - x) / (x - x);
    }
    if (ix < 0x3fe00000u) { /* |x| < 0.5 */
        if (ix < 0x3e400000u) return x; /* |x| < 2^-27 */
        const double w = asin_ratio(x * x);
        return x + x * w;
    }
    /* 1 > |x| >= 0.5 */
    const double ax = det_fabs(x);
    double w = 1.0 - ax;
    double t = w * 0.5;
    const double r = asin_ratio(t);
    const double s = det_sqrt(t);
    if (ix >= 0x3fef3333u) { /* |x| > 0.975 */
        t = pio2_hi - (2.0 * (s + s * r) - pio2_lo);
    } else {
        w = clear_low_word(s);
        const double c = (t - w * w) / (s + w);
        const double p = 2.0 * s * r - (pio2_lo - 2.0 * c);
        const double q = pio4_hi - 2.0 * w;
        t = pio4_hi - (p - q);
    }
    return neg ? -t : t;
}

/* sin on [-pi/4, pi/4] with a tail y (x + y is the reduced argument) */
EUCL_HD double kernel_sin(double x, double y, bool have_tail) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    if (hi_word_abs(x) < 0x3e400000u) return x; /* |x| < 2^-27 */
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    if (!have_tail) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

/* cos on [-pi/4, pi/4] with a tail y */
EUCL_HD double kernel_cos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const uint32_t ix = hi_word_abs(x);
    if (ix < 0x3e400000u) return 1.0; /* |x| < 2^-27 */
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    if (ix < 0x3fd33333u) return 1.0 - (0.5 * z - (z * r - x * y)); /* |x| < 0.3 */
    double qx;
    if (ix > 0x3fe90000u) qx = 0.28125; /* |x| > 0.78125 */
    else qx = from_bits((uint64_t)(ix - 0x00200000u) << 32); /* about |x| / 4 */
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

/* Argument reduction x = n * pi/2 + (y0 + y1), |y0 + y1| <= pi/4, for |x| up to about 2^19 * pi/2
 * (three-term Cody-Waite with a cancellation check).  Returns n (only n mod 4 matters).
 * Larger |x| are outside the domain this library guarantees (the renderer never produces them):
 * `ok` is cleared and the caller falls back to the platform function. */
EUCL_HD int rem_pio2(double x, double& y0, double& y1, bool& ok) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                 pio2_1t = 6.07710050650619224932e-11, pio2_2 = 6.07710050630396597660e-11,
                 pio2_2t = 2.02226624879595063154e-21, pio2_3 = 2.02226624871116645580e-21,
                 pio2_3t = 8.47842766036889956997e-32;
    const uint32_t ix = hi_word_abs(x);
    ok = ix <= 0x413921fbu;
    if (!ok) {
        y0 = x;
        y1 = 0.0;
        return 0;
    }
    const double t0 = det_fabs(x);
    const int n = (int)(t0 * invpio2 + 0.5);
    const double fn = (double)n;
    double r = t0 - fn * pio2_1;
    double w = fn * pio2_1t;
    const int j = (int)(ix >> 20);
    y0 = r - w;
    int i = j - (int)((hi_word_abs(y0) >> 20) & 0x7ffu);
    if (i > 16) { /* 2nd iteration needed, good to 118 bits */
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y0 = r - w;
        i = j - (int)((hi_word_abs(y0) >> 20) & 0x7ffu);
        if (i > 49) { /* 3rd iteration, 151 bits */
            t = r;
            w = fn * pio2_3;
            r = t - w;
            w = fn * pio2_3t - ((t - r) - w);
            y0 = r - w;
        }
    }
    y1 = (r - y0) - w;
    if (det_signbit(x)) {
        y0 = -y0;
        y1 = -y1;
        return -n;
    }
    return n;
}

EUCL_HD double det_sin(double x) {
    const uint32_t ix = hi_word_abs(x);
    if (ix <= 0x3fe921fbu) return kernel_sin(x, 0.0, false); /* |x| <= ~pi/4 */
    if (ix >= 0x7ff00000u) return x - x;                      /* inf / NaN */
    double y0, y1;
    bool ok;
    const int n = rem_pio2(x, y0, y1, ok);
    if (!ok) return sin(x);
    switch (n & 3) {
    case 0: return kernel_sin(y0, y1, true);
    case 1: return kernel_cos(y0, y1);
    case 2: return -kernel_sin(y0, y1, true);
    default: return -kernel_cos(y0, y1);
    }
}

EUCL_HD double det_cos(double x) {
    const uint32_t ix = hi_word_abs(x);
    if (ix <= 0x3fe921fbu) return kernel_cos(x, 0.0);
    if (ix >= 0x7ff00000u) return x - x;
    double y0, y1;
    bool ok;
    const int n = rem_pio2(x, y0, y1, ok);
    if (!ok) return cos(x);
    switch (n & 3) {
    case 0: return kernel_cos(y0, y1);
    case 1: return -kernel_sin(y0, y1, true);
    case 2: return -kernel_cos(y0, y1);
    default: return kernel_sin(y0, y1, true);
    }
}

EUCL_HD double det_atan(double x) {
    const double atanhi0 = 4.63647609000806093515e-01, atanhi1 = 7.85398163397448278999e-01,
                 atanhi2 = 9.82793723247329054082e-01, atanhi3 = 1.57079632679489655800e+00;
    const double atanlo0 = 2.26987774529616870924e-17, atanlo1 = 3.06161699786838301793e-17,
                 atanlo2 = 1.39033110312309984516e-17, atanlo3 = 6.12323399573676603587e-17;
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    const uint32_t ix = hi_word_abs(x);
    const bool neg = det_signbit(x);
    if (ix >= 0x44100000u) { /* |x| >= 2^66 */
        if (x != x) return x + x;
        return neg ? -(atanhi3 + atanlo3) : atanhi3 + atanlo3;
    }
    int id;
    if (ix < 0x3fdc0000u) { /* |x| < 0.4375 */
        if (ix < 0x3e200000u) return x; /* |x| < 2^-29 */
        id = -1;
    } else {
        x = det_fabs(x);
        if (ix < 0x3ff30000u) {     /* |x| < 1.1875 */
            if (ix < 0x3fe60000u) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else { /* 11/16 <= |x| < 19/16 */
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000u) { /* |x| < 2.4375 */
                id = 2;
                x = (x - 1.5) / (1.0 + 1.5 * x);
            } else { /* 2.4375 <= |x| < 2^66 */
                id = 3;
                x = -1.0 / x;
            }
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const double hi = id == 0 ? atanhi0 : id == 1 ? atanhi1 : id == 2 ? atanhi2 : atanhi3;
    const double lo = id == 0 ? atanlo0 : id == 1 ? atanlo1 : id == 2 ? atanlo2 : atanlo3;
    const double r = hi - ((x * (s1 + s2) - lo) - x);
    return neg ? -r : r;
}

EUCL_HD double det_atan2(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_o_2 = 1.5707963267948965580E+00, pi_o_4 = 7.8539816339744827900E-01,
                 pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    if (x == 1.0) return det_atan(y);
    const int m = (det_signbit(y) ? 1 : 0) | (det_signbit(x) ? 2 : 0);
    const uint32_t ix = hi_word_abs(x), iy = hi_word_abs(y);
    const bool x_inf = ix >= 0x7ff00000u, y_inf = iy >= 0x7ff00000u;
    if (y == 0.0) {
        switch (m) {
        case 0:
        case 1: return y; /* atan(+-0, +anything) = +-0 */
        case 2: return pi; /* atan(+0, -anything) = pi */
        default: return -pi;
        }
    }
    if (x == 0.0) return (m & 1) ? -pi_o_2 : pi_o_2;
    if (x_inf) {
        if (y_inf) {
            switch (m) {
            case 0: return pi_o_4;
            case 1: return -pi_o_4;
            case 2: return 3.0 * pi_o_4;
            default: return -3.0 * pi_o_4;
            }
        }
        switch (m) {
        case 0: return 0.0;
        case 1: return -0.0;
        case 2: return pi;
        default: return -pi;
        }
    }
    if (y_inf) return (m & 1) ? -pi_o_2 : pi_o_2;
    const int k = ((int)iy - (int)ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;   /* |y/x| > 2^60 */
    else if ((m & 2) && k < -60) z = 0.0;    /* |y|/x < -2^60 */
    else z = det_atan(det_fabs(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

} /* namespace eucl_det */

#endif /* EUCL_DETMATH_H */
