// oracle.cc -- CPU restatement of euclider's per-pixel trace loop.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (euclider_b200/, include/) may import,
// link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs do, and only as the checker / the CPU number.
//
// It follows the reference (Limeth/euclider, Rust, CPU-only) function by function, keeping the
// recursive shape of the original -- recursion for the ray tree, lazily pulled + cached streams
// for CSG -- which is deliberately NOT how the CUDA pipeline is organised (wavefront, eager CSG
// lists), so the two are independent statements of the same semantics.  All arithmetic is f64 in
// the reference's operation order; build with -ffp-contract=off (see Makefile).
//
// PARITY PINNING: the reference cannot be compiled here (no rustc/cargo, 20+ crates absent), so
// this oracle is pinned by (1) the reference's own known-answer tests for this path
// (shape.rs:1048-1148, util.rs:947-969,1007-1037; see tests/test_oracle_kats.py), (2) restatements
// of the PUBLISHED definitions of the third-party arithmetic written independently of this file
// (W3C / SVG compositing equations for palette's 17 blends, colorsys for HSV, the Fresnel equations in
// sine / tangent form, exact integer arithmetic for json's number conversion:
// tests/test_independent_pins.py) and (3) code reading.  What remains "parity unpinned" -- noise 0.4.1's
// gradient table, nalgebra 0.8.2's operation order, image 0.18's decoders, meval's grammar -- is listed
// with its alternatives in oracle/ASSUMPTIONS.md.
//
// Builds (oracle/Makefile): liboracle.so (glibc libm), liboracle_det.so (the deterministic libm the CUDA
// path uses: bit-comparable), liboracle_f32.so (-DORACLE_F32: the reference's `low_precision` feature,
// `type F = f32`, src/main.rs:46-49), liboracle_count.so (operation counters).
//
// Citations are relative to /root/reference/src/.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

#include "euclider_b200.h"
#include "eucl_detmath.h"

// BEGIN_KEEP64
// The reference's scalar type `F` (src/main.rs:46-49): f64, or f32 with its `low_precision` feature
// (-DORACLE_F32: liboracle_f32.so).  Scene tables, the camera's JSON pose and LinearSpace expressions stay f64
// (JSON numbers parse as f64, meval evaluates in f64) and are narrowed with `as F` where the reference does.
#ifdef ORACLE_F32
typedef float real;
#else
typedef double real;
#endif
#define R(x) ((real)(x))
// END_KEEP64

namespace {

// ---------------------------------------------------------------------------------------------
// ORACLE_COUNT_FLOPS build (liboracle_count.so): counts the floating-point operations of a render so that
// the work per ray segment can be set against the algorithmic bound of SURVEY.md 8(d) ("F_oracle").
// Vector arithmetic is counted exactly (every +, -, *, / of the Vec operators and dot); the scalar
// tails of the primitive intersectors, the membership tests and angle_between add the constants
// written next to them; one libm call counts as one operation.  Colour arithmetic (palette blends,
// quantisation) is not counted.  Never used for timing.
#ifdef ORACLE_COUNT_FLOPS
static thread_local uint64_t g_flops = 0;
#define FLOPS(n) (g_flops += (uint64_t)(n))
#else
#define FLOPS(n) ((void)0)
#endif

// Transcendental functions.  The reference calls the platform libm through Rust's std (f64::acos
// ...).  Two builds of this oracle exist:
//   liboracle.so      (default)           : the host's glibc -- what the Rust binary would link here
//   liboracle_det.so  (-DORACLE_DETMATH)  : include/eucl_detmath.h, the fdlibm-style libm the CUDA
//                                           path uses, for BIT-EXACT comparison with the GPU
// BEGIN_KEEP64
namespace om64 {
#ifdef ORACLE_DETMATH
inline double acos(double x) { FLOPS(1); return eucl_det::det_acos(x); }
inline double asin(double x) { FLOPS(1); return eucl_det::det_asin(x); }
inline double sin(double x) { FLOPS(1); return eucl_det::det_sin(x); }
inline double cos(double x) { FLOPS(1); return eucl_det::det_cos(x); }
inline double atan(double x) { FLOPS(1); return eucl_det::det_atan(x); }
inline double atan2(double y, double x) { FLOPS(1); return eucl_det::det_atan2(y, x); }
#else
inline double acos(double x) { FLOPS(1); return std::acos(x); }
inline double asin(double x) { FLOPS(1); return std::asin(x); }
inline double sin(double x) { FLOPS(1); return std::sin(x); }
inline double cos(double x) { FLOPS(1); return std::cos(x); }
inline double atan(double x) { FLOPS(1); return std::atan(x); }
inline double atan2(double y, double x) { FLOPS(1); return std::atan2(y, x); }
#endif
} // namespace om64
// `real` flavour: evaluated in f64 and narrowed (f32 build: what a good acosf returns in all but rare halfway cases; the
// CUDA f32 kernels do exactly the same, so the two agree bit for bit)
namespace om {
inline real acos(real x) { return (real)om64::acos((double)x); }
inline real asin(real x) { return (real)om64::asin((double)x); }
inline real sin(real x) { return (real)om64::sin((double)x); }
inline real cos(real x) { return (real)om64::cos((double)x); }
inline real atan(real x) { return (real)om64::atan((double)x); }
inline real atan2(real y, real x) { return (real)om64::atan2((double)y, (double)x); }
} // namespace om
// END_KEEP64

constexpr real PI = R(3.14159265358979323846264338327950288);     // BaseFloat::pi()
constexpr real FRAC_PI_2 = R(1.57079632679489661923132169163975144); // BaseFloat::frac_pi_2()
constexpr real APPROX_EPSILON = R(1.0e-6); // nalgebra 0.8.2 ApproxEq::approx_epsilon (RECOLLECTION)

// ---------------------------------------------------------------------------------------------
// nalgebra 0.8 vector semantics: component-wise loops, dot = left-to-right sum in index order,
// normalize = v / norm(v)
template <int D>
struct Vec {
    real c[D];
    real& operator[](int k) { return c[k]; }
    real operator[](int k) const { return c[k]; }
};
template <int D>
Vec<D> load(const double* p) { // scene tables and poses are f64: `as F`
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = (real)p[k];
    return r;
}
template <int D>
Vec<D> operator+(const Vec<D>& a, const Vec<D>& b) {
    Vec<D> r;
    FLOPS(D);
    for (int k = 0; k < D; ++k) r[k] = a[k] + b[k];
    return r;
}
template <int D>
Vec<D> operator-(const Vec<D>& a, const Vec<D>& b) {
    Vec<D> r;
    FLOPS(D);
    for (int k = 0; k < D; ++k) r[k] = a[k] - b[k];
    return r;
}
template <int D>
Vec<D> operator-(const Vec<D>& a) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = -a[k];
    return r;
}
template <int D>
Vec<D> operator*(const Vec<D>& a, real s) {
    Vec<D> r;
    FLOPS(D);
    for (int k = 0; k < D; ++k) r[k] = a[k] * s;
    return r;
}
template <int D>
Vec<D> operator/(const Vec<D>& a, real s) {
    Vec<D> r;
    FLOPS(D);
    for (int k = 0; k < D; ++k) r[k] = a[k] / s;
    return r;
}
template <int D>
real dot(const Vec<D>& a, const Vec<D>& b) {
    FLOPS(2 * D - 1);
    real s = a[0] * b[0];
    for (int k = 1; k < D; ++k) s = s + a[k] * b[k];
    return s;
}
template <int D>
real norm_squared(const Vec<D>& a) { return dot(a, a); }
template <int D>
real norm(const Vec<D>& a) {
    FLOPS(1);
    return std::sqrt(norm_squared(a));
}
template <int D>
Vec<D> normalize(const Vec<D>& a) { return a / norm(a); }

// util.rs:712-722
template <int D>
real angle_between(const Vec<D>& a, const Vec<D>& b) {
    FLOPS(2); // *, / (the acos counts itself)
    real result = om::acos(dot(a, b) / (norm(a) * norm(b)));
    return std::isnan(result) ? R(0.0) : result;
}

// Rust f64::signum
real rust_signum(real x) {
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -R(1.0) : R(1.0);
}
// Rust f64::min / f64::max (ignore a NaN operand)
real rust_min(real a, real b) { return std::fmin(a, b); }
real rust_max(real a, real b) { return std::fmax(a, b); }

// util.rs:287-299
real remainder_f(real a, real b) {
    real rem = std::fmod(a, b);
    if (rem == R(0.0)) return R(0.0);
    if (a < R(0.0)) return b + rem;
    return rem;
}
int64_t remainder_i(int64_t a, int64_t b) {
    int64_t rem = a % b;
    if (rem == 0) return 0;
    if (a < 0) return b + rem;
    return rem;
}

// ---------------------------------------------------------------------------------------------
// palette 0.2.1 (RECOLLECTION): linear RGBA, no gamma on this path
struct Rgba {
    real r, g, b, a;
};
real clamp01(real v) { return v < R(0.0) ? R(0.0) : (v > R(1.0) ? R(1.0) : v); }

struct Counters {
    uint64_t nan_channel = 0; // App. A.6: NaN colour channel reaching to_pixel (reference panics)
    uint64_t bad_texcoord = 0; // NaN / out-of-range texture index (reference panics)
    uint64_t no_material = 0;  // material_at None on an exiting transmit (reference may panic)
    uint64_t csg_runaway = 0;  // a CSG stream that never terminates (reference hangs)
};

uint8_t channel_to_u8(real c, Counters* cn) {
    if (std::isnan(c)) { // reference: to_u8().unwrap() panics; defined here as 0 and counted
        if (cn) cn->nan_channel++;
        return 0;
    }
    return (uint8_t)(clamp01(c) * R(255.0)); // truncation
}
void to_pixel4(const Rgba& c, uint8_t out[4], Counters* cn) {
    out[0] = channel_to_u8(c.r, cn);
    out[1] = channel_to_u8(c.g, cn);
    out[2] = channel_to_u8(c.b, cn);
    out[3] = channel_to_u8(c.a, cn);
}
Rgba new_u8(const uint8_t p[4]) {
    return Rgba{(real)p[0] / R(255.0), (real)p[1] / R(255.0), (real)p[2] / R(255.0), (real)p[3] / R(255.0)};
}
struct Pre { // PreAlpha<Rgb>
    real r, g, b, a;
};
Pre into_premultiplied(const Rgba& c) {
    real a = clamp01(c.a);
    return Pre{c.r * a, c.g * a, c.b * a, a};
}
Rgba from_premultiplied(const Pre& p) {
    real a = clamp01(p.a);
    if (std::isnormal(a)) return Rgba{p.r / a, p.g / a, p.b / a, a};
    return Rgba{R(0.0), R(0.0), R(0.0), a};
}

// Blend on premultiplied colours: `s` is self (source), `d` the argument (destination).
// darken / difference / over are what the benchmark scenes use; the rest follow the W3C
// compositing formulas palette implements.
Pre blend_pre(int fn, const Pre& s, const Pre& d) {
    const real sa = s.a, da = d.a;
    real alpha = clamp01(sa + da - sa * da);
    auto each = [&](real a, real b) -> real {
        switch (fn) {
        case EUCL_BLEND_OVER: return a + b * (R(1.0) - sa);
        case EUCL_BLEND_INSIDE: return a * da;
        case EUCL_BLEND_OUTSIDE: return a * (R(1.0) - da);
        case EUCL_BLEND_ATOP: return a * da + b * (R(1.0) - sa);
        case EUCL_BLEND_XOR: return a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_PLUS: return a + b;
        case EUCL_BLEND_MULTIPLY: return a * b + a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_SCREEN: return a + b - a * b;
        case EUCL_BLEND_OVERLAY:
            if (b * R(2.0) <= da) return R(2.0) * a * b + a * (R(1.0) - da) + b * (R(1.0) - sa);
            return a * (R(1.0) + da) + b * (R(1.0) + sa) - R(2.0) * a * b - da * sa;
        case EUCL_BLEND_DARKEN: return rust_min(a * da, b * sa) + a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_LIGHTEN: return rust_max(a * da, b * sa) + a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_DODGE:
            if (a == sa && !std::isnormal(b)) return a * (R(1.0) - da);
            if (a == sa) return sa * da + a * (R(1.0) - da) + b * (R(1.0) - sa);
            return sa * da * rust_min(R(1.0), (b / da) * sa / (sa - a)) + a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_BURN:
            if (!std::isnormal(a) && b == da) return sa * da + b * (R(1.0) - sa);
            if (!std::isnormal(a)) return b * (R(1.0) - sa);
            return sa * da * (R(1.0) - rust_min(R(1.0), (R(1.0) - b / da) * sa / a)) + a * (R(1.0) - da) + b * (R(1.0) - sa);
        case EUCL_BLEND_HARD_LIGHT:
            if (a * R(2.0) <= sa) return R(2.0) * a * b + a * (R(1.0) - da) + b * (R(1.0) - sa);
            return a * (R(1.0) + da) + b * (R(1.0) + sa) - R(2.0) * a * b - da * sa;
        case EUCL_BLEND_SOFT_LIGHT: {
            real m = std::isnormal(da) ? b / da : R(0.0);
            if (a * R(2.0) <= sa) return b * (sa + (R(2.0) * a - sa) * (R(1.0) - m)) + a * (R(1.0) - da) + b * (R(1.0) - sa);
            if (b * R(4.0) <= da) {
                real m2 = m * m, m3 = m2 * m;
                return da * (R(2.0) * a - sa) * (m3 * R(16.0) - m2 * R(12.0) - m * R(3.0)) + a - a * da + b;
            }
            return da * (R(2.0) * a - sa) * (std::sqrt(m) - m) + a - a * da + b;
        }
        case EUCL_BLEND_DIFFERENCE: return a + b - R(2.0) * rust_min(a * da, b * sa);
        case EUCL_BLEND_EXCLUSION: return a + b - R(2.0) * a * b;
        }
        return a;
    };
    switch (fn) {
    case EUCL_BLEND_INSIDE: alpha = clamp01(sa * da); break;
    case EUCL_BLEND_OUTSIDE: alpha = clamp01(sa * (R(1.0) - da)); break;
    case EUCL_BLEND_ATOP: alpha = clamp01(da); break;
    case EUCL_BLEND_XOR: alpha = clamp01(sa + da - R(2.0) * sa * da); break;
    case EUCL_BLEND_PLUS: alpha = clamp01(sa + da); break;
    default: break;
    }
    return Pre{each(s.r, d.r), each(s.g, d.g), each(s.b, d.b), alpha};
}

// util.rs:265-285
Rgba combine_palette_color(const Rgba& a, const Rgba& b, real a_ratio) {
    if (a_ratio <= R(0.0)) return b;
    if (a_ratio >= R(1.0)) return a;
    return Rgba{a.r * a_ratio + b.r * (R(1.0) - a_ratio), a.g * a_ratio + b.g * (R(1.0) - a_ratio),
                a.b * a_ratio + b.b * (R(1.0) - a_ratio), a.a * a_ratio + b.a * (R(1.0) - a_ratio)};
}

// surface.rs:295-390: blend_function_ratio is combine_palette_color, the named ones go through
// premultiplied alpha
Rgba blend_rgba(int fn, real ratio, const Rgba& source, const Rgba& destination) {
    if (fn == EUCL_BLEND_RATIO) return combine_palette_color(source, destination, ratio);
    return from_premultiplied(blend_pre(fn, into_premultiplied(source), into_premultiplied(destination)));
}

// palette Hsv -> Rgb, hue in degrees
void hsv_to_rgb(real hue_degrees, real saturation, real value, real rgb[3]) {
    real deg = hue_degrees;
    if (std::isfinite(deg)) {
        while (deg >= R(360.0)) deg = deg - R(360.0);
        while (deg < R(0.0)) deg = deg + R(360.0);
    }
    real c = value * saturation;
    real h = deg / R(60.0);
    real x = c * (R(1.0) - std::fabs(std::fmod(h, R(2.0)) - R(1.0)));
    real m = value - c;
    real r, g, b;
    if (h >= R(0.0) && h < R(1.0)) { r = c; g = x; b = R(0.0); }
    else if (h >= R(1.0) && h < R(2.0)) { r = x; g = c; b = R(0.0); }
    else if (h >= R(2.0) && h < R(3.0)) { r = R(0.0); g = c; b = x; }
    else if (h >= R(3.0) && h < R(4.0)) { r = R(0.0); g = x; b = c; }
    else if (h >= R(4.0) && h < R(5.0)) { r = x; g = R(0.0); b = c; }
    else { r = c; g = R(0.0); b = x; }
    rgb[0] = r + m;
    rgb[1] = g + m;
    rgb[2] = b + m;
}

// ---------------------------------------------------------------------------------------------
// noise 0.4.1 Perlin, 4-D (RECOLLECTION; see ASSUMPTIONS.md)
void perlin_grad4(unsigned index, real g[4]) {
    const real diag = R(0.577350269189625764077083524672081875);
    unsigned i = index % 32;
    unsigned zero_at = i / 8, signs = i % 8;
    int bit = 0;
    for (unsigned k = 0; k < 4; ++k) {
        if (k == zero_at) {
            g[k] = R(0.0);
        } else {
            g[k] = (signs >> bit) & 1u ? -diag : diag;
            ++bit;
        }
    }
}
real perlin4(const uint8_t perm[256], const real point[4]) {
    real floored[4], near_d[4], far_d[4];
    long near_c[4], far_c[4];
    for (int k = 0; k < 4; ++k) {
        floored[k] = std::floor(point[k]);
        near_c[k] = (long)floored[k];
        far_c[k] = near_c[k] + 1;
        near_d[k] = point[k] - floored[k];
        far_d[k] = near_d[k] - R(1.0);
    }
    real total = R(0.0);
    bool first = true;
    // corner order f0000, f1000, f0100, f1100, ... (x fastest)
    for (int corner = 0; corner < 16; ++corner) {
        long c[4];
        real d[4];
        for (int k = 0; k < 4; ++k) {
            bool far = (corner >> k) & 1;
            c[k] = far ? far_c[k] : near_c[k];
            d[k] = far ? far_d[k] : near_d[k];
        }
        real attn = R(1.0) - (((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) + d[3] * d[3]);
        real v = R(0.0);
        if (attn > R(0.0)) {
            unsigned h = perm[(unsigned)(c[0] & 0xff)];
            h = perm[h ^ (unsigned)(c[1] & 0xff)];
            h = perm[h ^ (unsigned)(c[2] & 0xff)];
            h = perm[h ^ (unsigned)(c[3] & 0xff)];
            real g[4];
            perlin_grad4(h, g);
            real a2 = attn * attn;
            v = (a2 * a2) * (((d[0] * g[0] + d[1] * g[1]) + d[2] * g[2]) + d[3] * g[3]);
        }
        if (first) {
            total = v;
            first = false;
        } else {
            total = total + v;
        }
    }
    return total * R(4.424369240215691);
}

// ---------------------------------------------------------------------------------------------
// scene access

template <int D>
struct Hit { // shape.rs:88-109 (the `direction` copy is the ray direction; kept by the caller)
    Vec<D> location;
    Vec<D> normal;
    real distance;
};

// --- primitives ------------------------------------------------------------------------
// returns number of hits (0..2), sorted by distance
template <int D>
int intersect_prim(const EuclPrim& pr, const Vec<D>& location, const Vec<D>& direction, Hit<D> out[2]) {
    switch (pr.kind) {
    case EUCL_PRIM_VOID: return 0; // shape.rs:622-631
    case EUCL_PRIM_SPHERE: { // shape.rs:652-731
        Vec<D> center = load<D>(pr.v0);
        real radius = R(pr.s0);
        Vec<D> rel = location - center;
        real a = norm_squared(direction);
        real b = R(2.0) * dot(direction, rel);
        real c = norm_squared(rel) - radius * radius;
        real d = b * b - R(4.0) * a * c;
        FLOPS(7); // reject stage, scalar part
        if (d < R(0.0)) return 0;
        FLOPS(9); // sqrt, two roots
        real d_sqrt = std::sqrt(d);
        real t1 = (-b - d_sqrt) / (R(2.0) * a);
        real t2 = (-b + d_sqrt) / (R(2.0) * a);
        real t_first, t_second = R(0.0);
        bool has_first = false, has_second = false;
        if (t1 >= R(0.0)) {
            t_first = t1;
            has_first = true;
            if (t2 >= R(0.0)) {
                t_second = t2;
                has_second = true;
            }
        } else if (t2 >= R(0.0)) {
            t_first = t2;
            has_first = true;
        }
        if (!has_first) return 0;
        out[0].location = location + direction * t_first;
        out[0].normal = normalize(out[0].location - center);
        out[0].distance = t_first;
        if (!has_second) return 1;
        out[1].location = location + direction * t_second;
        out[1].normal = normalize(out[1].location - center);
        out[1].distance = t_second;
        return 2;
    }
    case EUCL_PRIM_HYPERPLANE:
    case EUCL_PRIM_HALFSPACE: { // shape.rs:779-809, 843-870
        Vec<D> n = load<D>(pr.v0);
        real t = -(dot(n, location) + R(pr.s0)) / dot(n, direction);
        FLOPS(3);
        if (t < R(0.0)) return 0; // NaN and +inf pass, as in the reference
        out[0].location = direction * t + location;
        out[0].normal = n;
        out[0].distance = t;
        if (pr.kind == EUCL_PRIM_HALFSPACE) out[0].normal = out[0].normal * -R(pr.s1);
        return 1;
    }
    case EUCL_PRIM_CYLINDER: { // shape.rs:935-1027
        Vec<D> center = load<D>(pr.v0), axis = load<D>(pr.v1);
        real radius = R(pr.s0);
        Vec<D> a_vec = direction - axis * dot(direction, axis);
        Vec<D> delta_location = location - center;
        Vec<D> c_vec = delta_location - axis * dot(delta_location, axis);
        real a = norm_squared(a_vec);
        real b = (R(1.0) + R(1.0)) * dot(a_vec, c_vec);
        real c = norm_squared(c_vec) - radius * radius;
        real d = b * b - R(4.0) * a * c;
        FLOPS(7);
        if (d < R(0.0)) return 0;
        FLOPS(9);
        real d_sqrt = std::sqrt(d);
        real t1 = (-b - d_sqrt) / (R(2.0) * a);
        real t2 = (-b + d_sqrt) / (R(2.0) * a);
        real t_first, t_second = R(0.0);
        bool has_first = false, has_second = false;
        if (t1 >= R(0.0)) {
            t_first = t1;
            has_first = true;
            if (t2 >= R(0.0)) {
                t_second = t2;
                has_second = true;
            }
        } else if (t2 >= R(0.0)) {
            t_first = t2;
            has_first = true;
        }
        if (!has_first) return 0;
        Vec<D> p1 = location + direction * t_first;
        // get_closest_point_on_axis (shape.rs:929-932) of the FIRST hit, reused for the second
        Vec<D> on_axis = axis * dot(axis, p1 - center) + center;
        out[0].location = p1;
        out[0].normal = normalize(p1 - on_axis);
        out[0].distance = t_first;
        if (!has_second) return 1;
        Vec<D> p2 = location + direction * t_second;
        out[1].location = p2;
        out[1].normal = normalize(p2 - on_axis);
        out[1].distance = t_second;
        return 2;
    }
    }
    return 0;
}

template <int D>
bool prim_inside(const EuclPrim& pr, const Vec<D>& point) {
    switch (pr.kind) {
    case EUCL_PRIM_VOID: return true;       // shape.rs:614-619
    case EUCL_PRIM_SPHERE: {                // shape.rs:734-738
        Vec<D> center = load<D>(pr.v0);
        FLOPS(1);
        return norm_squared(center - point) <= R(pr.s0) * R(pr.s0);
    }
    case EUCL_PRIM_HYPERPLANE: return false; // shape.rs:812-817
    case EUCL_PRIM_HALFSPACE: {              // shape.rs:873-881
        FLOPS(1);
        real result = dot(load<D>(pr.v0), point) + R(pr.s0);
        return R(pr.s1) == rust_signum(result);
    }
    case EUCL_PRIM_CYLINDER: { // shape.rs:1030-1038
        Vec<D> center = load<D>(pr.v0), axis = load<D>(pr.v1);
        Vec<D> on_axis = axis * dot(axis, point - center) + center;
        FLOPS(1);
        return norm_squared(point - on_axis) <= R(pr.s0) * R(pr.s0);
    }
    }
    return false;
}


template <int D>
struct Tracer {
    const EuclFlatScene& s;
    EuclCamera cam;
    double time_seconds; // f64 whatever `real` is: the Duration arithmetic of the reference is not generic over F
    uint32_t max_depth;
    Counters counters;
    uint64_t level_counts[EUCL_MAX_LEVELS] = {0};

    Tracer(const EuclFlatScene& scene, const EuclCamera& camera, double time)
        : s(scene), cam(camera), time_seconds(time), max_depth(camera.max_depth) {}

    // node helpers: post-order layout (include/euclider_b200.h)
    int child_b(int n) const { return n - 1; }
    int child_a(int n) const { return s.nodes[n - 1].first - 1; }

    // shape.rs:587-601
    bool node_inside(int n, const Vec<D>& point) const {
        const EuclNode& nd = s.nodes[n];
        switch (nd.op) {
        case EUCL_CSG_LEAF: return prim_inside(s.prims[nd.prim], point);
        case EUCL_CSG_UNION: return node_inside(child_a(n), point) || node_inside(child_b(n), point);
        case EUCL_CSG_INTERSECTION: return node_inside(child_a(n), point) && node_inside(child_b(n), point);
        case EUCL_CSG_COMPLEMENT: return node_inside(child_a(n), point) && !node_inside(child_b(n), point);
        case EUCL_CSG_SYMDIFF: return node_inside(child_a(n), point) ^ node_inside(child_b(n), point);
        }
        return false;
    }

    // --- lazily pulled, cached CSG streams (util.rs:372-450 Provider; shape.rs:204-497) ------
    struct Opt {
        bool some;
        Hit<D> hit;
    };
    struct NodeState {
        std::vector<Opt> items; // Provider cache, including cached None entries
        int index_a = 0, index_b = 0;
        int leaf_count = -1, leaf_next = 0;
        Hit<D> leaf_hits[2];
    };
    struct Streams {
        const Tracer& tr;
        int node_first;
        Vec<D> location, direction;
        std::vector<NodeState> st;
        bool runaway = false;

        explicit Streams(const Tracer& t) : tr(t), node_first(0) {}
        // one Streams object is reused for every intersect call of a thread (only its storage: the
        // lazily filled caches are reset), so the oracle does not spend its time in malloc
        void reset(int first, int root, const Vec<D>& loc, const Vec<D>& dir) {
            node_first = first;
            location = loc;
            direction = dir;
            runaway = false;
            const size_t n = (size_t)(root - first + 1);
            if (st.size() < n) st.resize(n);
            for (size_t k = 0; k < n; ++k) {
                st[k].items.clear();
                st[k].index_a = st[k].index_b = 0;
                st[k].leaf_count = -1;
                st[k].leaf_next = 0;
            }
        }

        // Provider::get (util.rs:395-419)
        Opt get(int n, int index) {
            NodeState& ns = st[(size_t)(n - node_first)];
            while (index >= (int)ns.items.size()) {
                if (ns.items.size() > 4096) { // the reference would never return
                    runaway = true;
                    return Opt{false, {}};
                }
                Opt item = next(n);
                st[(size_t)(n - node_first)].items.push_back(item);
            }
            return st[(size_t)(n - node_first)].items[(size_t)index];
        }

        Opt next(int n) {
            const EuclNode& nd = tr.s.nodes[n];
            NodeState& ns = st[(size_t)(n - node_first)];
            const Opt none{false, {}};
            if (nd.op == EUCL_CSG_LEAF) {
                if (ns.leaf_count < 0) ns.leaf_count = intersect_prim<D>(tr.s.prims[nd.prim], location, direction, ns.leaf_hits);
                if (ns.leaf_next < ns.leaf_count) return Opt{true, ns.leaf_hits[ns.leaf_next++]};
                ns.leaf_next++;
                return none;
            }
            const int na = tr.child_a(n), nb = tr.child_b(n);
            switch (nd.op) {
            case EUCL_CSG_UNION: // shape.rs:212-264
                for (;;) {
                    if (runaway) return none;
                    Opt a = get(na, ns.index_a), b = get(nb, ns.index_b);
                    if (a.some) {
                        if (b.some) {
                            bool a_closer = a.hit.distance < b.hit.distance;
                            const Hit<D>& closer = a_closer ? a.hit : b.hit;
                            int further = a_closer ? nb : na;
                            int& closer_index = a_closer ? ns.index_a : ns.index_b;
                            if (!tr.node_inside(further, closer.location)) {
                                closer_index += 1;
                                return Opt{true, closer};
                            }
                            closer_index += 1;
                        } else {
                            ns.index_a += 1;
                            if (tr.node_inside(nb, a.hit.location)) return none;
                            return a;
                        }
                    } else {
                        if (b.some) {
                            ns.index_b += 1;
                            if (tr.node_inside(na, b.hit.location)) return none;
                        }
                        return b;
                    }
                }
            case EUCL_CSG_INTERSECTION: // shape.rs:291-340
                for (;;) {
                    if (runaway) return none;
                    Opt a = get(na, ns.index_a), b = get(nb, ns.index_b);
                    if (a.some) {
                        if (b.some) {
                            bool a_closer = a.hit.distance < b.hit.distance;
                            const Hit<D>& closer = a_closer ? a.hit : b.hit;
                            int further = a_closer ? nb : na;
                            (a_closer ? ns.index_a : ns.index_b) += 1;
                            if (tr.node_inside(further, closer.location)) return Opt{true, closer};
                        } else {
                            ns.index_a += 1;
                            if (tr.node_inside(nb, a.hit.location)) return a;
                            return none;
                        }
                    } else {
                        if (b.some) {
                            ns.index_b += 1;
                            if (tr.node_inside(na, b.hit.location)) return b;
                        }
                        return none;
                    }
                }
            case EUCL_CSG_COMPLEMENT: // shape.rs:365-409
                for (;;) {
                    if (runaway) return none;
                    Opt a = get(na, ns.index_a), b = get(nb, ns.index_b);
                    if (a.some) {
                        if (b.some) {
                            if (a.hit.distance < b.hit.distance) {
                                ns.index_a += 1;
                                if (!tr.node_inside(nb, a.hit.location)) return a;
                            } else {
                                ns.index_b += 1;
                                if (tr.node_inside(na, b.hit.location)) {
                                    Opt inv = b;
                                    inv.hit.normal = -inv.hit.normal;
                                    return inv;
                                }
                            }
                        } else {
                            return a; // NOT advanced (reference quirk, shape.rs:392)
                        }
                    } else {
                        if (b.some) {
                            ns.index_b += 1;
                            if (tr.node_inside(na, b.hit.location)) {
                                Opt inv = b;
                                inv.hit.normal = -inv.hit.normal;
                                return inv;
                            }
                        }
                        return none;
                    }
                }
            case EUCL_CSG_SYMDIFF: { // shape.rs:436-496 (not a loop)
                Opt a = get(na, ns.index_a), b = get(nb, ns.index_b);
                if (a.some) {
                    if (b.some) {
                        bool a_closer = a.hit.distance < b.hit.distance;
                        Opt closer = a_closer ? a : b;
                        int further = a_closer ? nb : na;
                        (a_closer ? ns.index_a : ns.index_b) += 1;
                        if (tr.node_inside(further, closer.hit.location)) closer.hit.normal = -closer.hit.normal;
                        return closer;
                    }
                    ns.index_a += 1;
                    if (tr.node_inside(nb, a.hit.location)) a.hit.normal = -a.hit.normal;
                    return a;
                }
                if (b.some) {
                    ns.index_b += 1;
                    if (tr.node_inside(na, b.hit.location)) b.hit.normal = -b.hit.normal;
                }
                return b;
            }
            }
            return none;
        }
    };

    std::unique_ptr<Streams> scratch;

    // Universe::intersect + `provider.iter().next()` (mod.rs:61-83,110-112): first item only
    bool first_intersection(const EuclEntity& e, const Vec<D>& location, const Vec<D>& direction, Hit<D>* out) {
        if (e.node_first == e.node_root) { // plain primitive: no stream machinery needed
            Hit<D> hits[2];
            int n = intersect_prim(s.prims[s.nodes[e.node_root].prim], location, direction, hits);
            if (n == 0) return false;
            *out = hits[0];
            return true;
        }
        if (!scratch) scratch.reset(new Streams(*this));
        Streams& streams = *scratch;
        streams.reset(e.node_first, e.node_root, location, direction);
        auto item = streams.get(e.node_root, 0);
        if (streams.runaway) counters.csg_runaway++;
        if (!item.some) return false;
        *out = item.hit;
        return true;
    }

    // all items of an entity's stream up to the first None (test hook)
    int all_intersections(const EuclEntity& e, const Vec<D>& location, const Vec<D>& direction, int max_items,
                          Hit<D>* out) {
        Streams streams(*this);
        streams.reset(e.node_first, e.node_root, location, direction);
        int n = 0;
        while (n < max_items) {
            auto item = streams.get(e.node_root, n);
            if (!item.some || streams.runaway) break;
            out[n++] = item.hit;
        }
        return n;
    }

    // --- materials (material.rs) ------------------------------------------------------------
    // BEGIN_KEEP64 (meval expressions are f64 in both precisions, material.rs:99-110)
    double eval_expr(int first, int len, const double* vars) const {
        double st[64];
        int sp = 0;
        for (int i = first; i < first + len; ++i) {
            const EuclExprOp& o = s.expr_ops[i];
            switch (o.op) {
            case EUCL_EX_CONST: st[sp++] = o.value; break;
            case EUCL_EX_VAR: st[sp++] = vars[o.arg]; break;
            case EUCL_EX_NEG: st[sp - 1] = -st[sp - 1]; break;
            case EUCL_EX_FUNC1: {
                double x = st[sp - 1], r = x;
                switch (o.arg) {
                case EUCL_FN_SQRT: r = std::sqrt(x); break;
                case EUCL_FN_ABS: r = std::fabs(x); break;
                case EUCL_FN_EXP: r = std::exp(x); break;
                case EUCL_FN_LN: r = std::log(x); break;
                case EUCL_FN_SIN: r = om64::sin(x); break;
                case EUCL_FN_COS: r = om64::cos(x); break;
                case EUCL_FN_TAN: r = std::tan(x); break;
                case EUCL_FN_ASIN: r = om64::asin(x); break;
                case EUCL_FN_ACOS: r = om64::acos(x); break;
                case EUCL_FN_ATAN: r = om64::atan(x); break;
                case EUCL_FN_SINH: r = std::sinh(x); break;
                case EUCL_FN_COSH: r = std::cosh(x); break;
                case EUCL_FN_TANH: r = std::tanh(x); break;
                case EUCL_FN_FLOOR: r = std::floor(x); break;
                case EUCL_FN_CEIL: r = std::ceil(x); break;
                case EUCL_FN_ROUND: r = std::round(x); break;
                case EUCL_FN_SIGNUM: r = std::isnan(x) ? x : (std::signbit(x) ? -1.0 : 1.0); break;
                }
                st[sp - 1] = r;
                break;
            }
            default: {
                double b = st[--sp], a = st[sp - 1], r = 0.0;
                switch (o.op) {
                case EUCL_EX_ADD: r = a + b; break;
                case EUCL_EX_SUB: r = a - b; break;
                case EUCL_EX_MUL: r = a * b; break;
                case EUCL_EX_DIV: r = a / b; break;
                case EUCL_EX_REM: r = std::fmod(a, b); break;
                case EUCL_EX_POW: r = std::pow(a, b); break;
                case EUCL_EX_FUNC2:
                    if (o.arg == EUCL_FN_ATAN2) r = om64::atan2(a, b);
                    else if (o.arg == EUCL_FN_MAX) r = std::fmax(a, b);
                    else r = std::fmin(a, b);
                    break;
                }
                st[sp - 1] = r;
            }
            }
        }
        return sp > 0 ? st[sp - 1] : 0.0;
    }
    // ComponentTransformation::transform_with (material.rs:91-112): every component expression
    // sees the SAME input vector
    void apply_transform(const EuclTransform& t, bool inverse, Vec<D>& v) const {
        double in[D];
        for (int k = 0; k < D; ++k) in[k] = v[k];
        for (int k = 0; k < D; ++k)
            v[k] = (real)(inverse ? eval_expr(t.inv_first[k], t.inv_len[k], in) : eval_expr(t.fwd_first[k], t.fwd_len[k], in));
    }
    // END_KEEP64
    void material_enter(int entity, Vec<D>& direction) const { // material.rs:133-137,150-154
        const EuclMaterial& m = s.materials[s.entities[entity].material];
        if (m.kind != EUCL_MAT_LINEAR_SPACE) return;
        for (int k = 0; k < m.n_transforms; ++k) apply_transform(s.transforms[m.transform_first + k], false, direction);
    }
    void material_exit(int entity, Vec<D>& direction) const { // material.rs:139-142,156-162
        const EuclMaterial& m = s.materials[s.entities[entity].material];
        if (m.kind != EUCL_MAT_LINEAR_SPACE) return;
        for (int k = m.n_transforms - 1; k >= 0; --k) apply_transform(s.transforms[m.transform_first + k], true, direction);
    }

    // mod.rs:229-251
    int material_at(const Vec<D>& location) const {
        for (int e = 0; e < s.n_entities; ++e)
            if (node_inside(s.entities[e].node_root, location)) return e;
        return -1;
    }

    // --- textures (surface.rs:434-542, d3/entity/surface.rs:60-68, d4/entity/surface.rs:11-15)
    Rgba texel(const EuclTexture& t, real xf, real yf) {
        // `<u32 as NumCast>::from(f)`: Some(trunc) iff -1 < f < 2^32; the reference unwraps
        int64_t x = 0, y = 0;
        if (!(xf > -R(1.0) && xf < R(4294967296.0)) || !(yf > -R(1.0) && yf < R(4294967296.0))) {
            counters.bad_texcoord++;
        } else {
            x = (int64_t)xf;
            y = (int64_t)yf;
        }
        if (x >= (int64_t)t.width || y >= (int64_t)t.height) { // image::get_pixel would panic
            counters.bad_texcoord++;
            x = 0;
            y = 0;
        }
        const uint8_t* p = s.texels + t.texel_offset + 4 * ((size_t)y * t.width + (size_t)x);
        return Rgba{(real)p[0], (real)p[1], (real)p[2], (real)p[3]};
    }
    Rgba sample_texture(const EuclMappedTexture& mt, real u, real v) {
        const EuclTexture& t = s.textures[mt.texture];
        real width = (real)t.width, height = (real)t.height;
        if (mt.filter == EUCL_TEX_NEAREST) { // surface.rs:434-451
            real x = std::floor(u * width), y = std::floor(v * height);
            if (!(x > -R(1.0) && x < R(4294967296.0)) || !(y > -R(1.0) && y < R(4294967296.0))) {
                counters.bad_texcoord++;
                x = R(0.0);
                y = R(0.0);
            }
            int64_t xi = remainder_i((int64_t)x, (int64_t)t.width), yi = remainder_i((int64_t)y, (int64_t)t.height);
            Rgba p = texel(t, (real)xi, (real)yi);
            return Rgba{p.r / R(255.0), p.g / R(255.0), p.b / R(255.0), p.a / R(255.0)};
        }
        // surface.rs:453-489
        real x = u * width - R(0.5), y = v * height - R(0.5);
        real offset_x = x - std::floor(x), offset_y = y - std::floor(y);
        Rgba px[4];
        const real ox[4] = {R(0.0), R(1.0), R(0.0), R(1.0)}, oy[4] = {R(0.0), R(0.0), R(1.0), R(1.0)};
        for (int k = 0; k < 4; ++k) px[k] = texel(t, remainder_f(x + ox[k], width), remainder_f(y + oy[k], height));
        auto mix = [&](real p0, real p1, real p2, real p3) {
            return ((p0 * (R(1.0) - offset_x) + p1 * offset_x) * (R(1.0) - offset_y) +
                    (p2 * (R(1.0) - offset_x) + p3 * offset_x) * offset_y) /
                   R(255.0);
        };
        return Rgba{mix(px[0].r, px[1].r, px[2].r, px[3].r), mix(px[0].g, px[1].g, px[2].g, px[3].g),
                    mix(px[0].b, px[1].b, px[2].b, px[3].b), mix(px[0].a, px[1].a, px[2].a, px[3].a)};
    }
    Rgba mapped_color(int mapped, const Vec<D>& point) {
        if (mapped < 0) return Rgba{R(0.0), R(0.0), R(0.0), R(0.0)}; // MappedTextureTransparent
        const EuclMappedTexture& mt = s.mapped_textures[mapped];
        // uv_sphere on the first three components (uv_derank drops w first)
        Vec<3> p;
        for (int k = 0; k < 3; ++k) p[k] = point[k] - R(mt.center[k]);
        p = normalize(p);
        real u = R(0.5) + om::atan2(p[1], p[0]) / (R(2.0) * PI);
        real v = R(0.5) - om::asin(p[2]) / PI;
        return sample_texture(mt, u, v);
    }

    // --- shading ------------------------------------------------------------------------------
    struct Context { // TracingContext (shape.rs:111-123)
        int origin_entity;
        int hit_entity;
        Vec<D> direction; // intersection.direction == the ray direction
        Hit<D> hit;
        Vec<D> normal_closer;
        bool exiting;
    };

    // util.rs:631-666
    Vec<D> general_rotation(const Vec<D>& self, const Vec<D>& other, real angle, const Vec<D>& v) const {
        real original[D][D], result[D][D]; // [row][col]
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) original[r][c] = r == c ? R(1.0) : R(0.0);
        for (int r = 0; r < D; ++r) {
            original[r][0] = self[r];
            original[r][1] = other[r];
        }
        std::memcpy(result, original, sizeof result);
        for (int i = 1; i < D; ++i) {
            for (int j = 0; j < i; ++j) {
                Vec<D> oc, rc;
                for (int r = 0; r < D; ++r) {
                    oc[r] = original[r][i];
                    rc[r] = result[r][j];
                }
                Vec<D> upd = oc - rc * dot(rc, oc);
                for (int r = 0; r < D; ++r) original[r][i] = upd[r];
            }
            Vec<D> col;
            for (int r = 0; r < D; ++r) col[r] = original[r][i];
            col = normalize(col);
            for (int r = 0; r < D; ++r) result[r][i] = col[r];
        }
        real rot[D][D];
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) rot[r][c] = r == c ? R(1.0) : R(0.0);
        rot[0][0] = om::cos(angle);
        rot[0][1] = -om::sin(angle);
        rot[1][0] = om::sin(angle);
        rot[1][1] = om::cos(angle);
        // result * (rotation_matrix * result.transpose()); nalgebra accumulates from zero
        real tmp[D][D], q[D][D];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                real acc = R(0.0);
                for (int k = 0; k < D; ++k) acc = acc + rot[i][k] * result[j][k];
                tmp[i][j] = acc;
            }
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                real acc = R(0.0);
                for (int k = 0; k < D; ++k) acc = acc + result[i][k] * tmp[k][j];
                q[i][j] = acc;
            }
        Vec<D> out;
        for (int i = 0; i < D; ++i) {
            real acc = R(0.0);
            for (int j = 0; j < D; ++j) acc = acc + v[j] * q[i][j];
            out[i] = acc;
        }
        return out;
    }

    real reflection_ratio(const EuclSurface& sf, const Context& c) const {
        if (sf.ratio_op == EUCL_RATIO_UNIFORM) return c.exiting ? R(0.0) : R(sf.ratio_a); // surface.rs:201-211
        // surface.rs:214-244
        Vec<D> normal = -c.normal_closer;
        real from_theta = angle_between(c.direction, normal);
        real from_index = c.exiting ? R(sf.ratio_a) : R(sf.ratio_b);
        real to_index = c.exiting ? R(sf.ratio_b) : R(sf.ratio_a);
        real to_theta = om::asin((from_index / to_index) * om::sin(from_theta));
        if (std::isnan(to_theta)) return R(1.0);
        real product_1_s = from_index * om::cos(from_theta);
        real product_2_s = to_index * om::cos(to_theta);
        real product_1_p = from_index * om::cos(to_theta);
        real product_2_p = to_index * om::cos(from_theta);
        real rs = (product_1_s - product_2_s) / (product_1_s + product_2_s);
        real rp = (product_1_p - product_2_p) / (product_1_p + product_2_p);
        real reflectance_s = rs * rs, reflectance_p = rp * rp;
        return (reflectance_s + reflectance_p) / (R(1.0) + R(1.0));
    }
    Vec<D> reflection_direction(const Context& c) const { // surface.rs:246-256
        return c.normal_closer * -R(2.0) * dot(c.direction, c.normal_closer) + c.direction;
    }
    Vec<D> threshold_direction(const EuclSurface& sf, const Context& c) const {
        if (sf.thr_op == EUCL_THR_IDENTITY) return c.direction; // surface.rs:259-266
        // surface.rs:268-288
        Vec<D> normal = -c.normal_closer;
        real from_theta = angle_between(c.direction, normal);
        real modifier = c.exiting ? R(sf.thr_a) : R(1.0) / R(sf.thr_a);
        real to_theta = om::asin(modifier * om::sin(from_theta));
        real angle_delta = to_theta - from_theta;
        return general_rotation(normal, c.direction, angle_delta, c.direction);
    }

    Rgba surface_color(const EuclSurface& sf, const Context& c) {
        Rgba stack[16];
        int sp = 0;
        for (int i = sf.color_first; i < sf.color_first + sf.color_len; ++i) {
            const EuclColorOp& op = s.color_ops[i];
            switch (op.op) {
            case EUCL_COL_UNIFORM: stack[sp++] = Rgba{R(op.f[0]), R(op.f[1]), R(op.f[2]), R(op.f[3])}; break;
            case EUCL_COL_ILLUM_GLOBAL: { // surface.rs:410-422
                Rgba light{R(op.f[0]), R(op.f[1]), R(op.f[2]), R(op.f[3])}, dark{R(op.f[4]), R(op.f[5]), R(op.f[6]), R(op.f[7])};
                real original_angle = angle_between(c.normal_closer, c.direction);
                real angle = PI - original_angle;
                real ratio = angle / FRAC_PI_2;
                stack[sp++] = combine_palette_color(dark, light, ratio);
                break;
            }
            case EUCL_COL_ILLUM_DIR: { // surface.rs:392-408
                Rgba light{R(op.f[0]), R(op.f[1]), R(op.f[2]), R(op.f[3])}, dark{R(op.f[4]), R(op.f[5]), R(op.f[6]), R(op.f[7])};
                Vec<D> light_direction = load<D>(&op.f[8]);
                Vec<D> normal = c.hit.normal;
                if (angle_between(c.direction, normal) > FRAC_PI_2) normal = -normal;
                real angle = angle_between(normal, -light_direction);
                real ratio = R(1.0) - angle / PI;
                stack[sp++] = combine_palette_color(dark, light, ratio);
                break;
            }
            case EUCL_COL_PERLIN_HUE: { // d3/entity/surface.rs:22-40
                // (time * 1000).as_secs() as f64 / 1000.0
                real time_millis = (real)(std::floor(time_seconds * 1000.0) / 1000.0); // Cast::from(f64) (d3/entity/surface.rs:32)
                real size = R(op.f[0]), speed = R(op.f[1]);
                real location[4] = {c.hit.location[0] / size, c.hit.location[1] / size, c.hit.location[2] / size,
                                      time_millis * speed};
                real value = perlin4(s.perlin_perm, location);
                real rgb[3];
                hsv_to_rgb(value * R(360.0), R(1.0), R(1.0), rgb);
                stack[sp++] = Rgba{rgb[0], rgb[1], rgb[2], R(1.0)};
                break;
            }
            case EUCL_COL_TEXTURE: stack[sp++] = mapped_color(op.i0, c.hit.location); break; // surface.rs:536-542
            case EUCL_COL_BLEND: { // surface.rs:295-307
                Rgba destination = stack[--sp];
                Rgba source = stack[--sp];
                stack[sp++] = blend_rgba(op.i0, R(op.f[0]), source, destination);
                break;
            }
            }
        }
        return sp > 0 ? stack[sp - 1] : Rgba{R(0.0), R(0.0), R(0.0), R(0.0)};
    }

    // mod.rs:85-147
    bool trace_closest(const Vec<D>& location, const Vec<D>& direction, Context* out) {
        bool have = false;
        real closest_distance = R(0.0);
        for (int e = 0; e < s.n_entities; ++e) {
            const EuclEntity& ent = s.entities[e];
            if (ent.surface < 0) continue; // the filter of mod.rs:158-160
            Hit<D> hit;
            if (!first_intersection(ent, location, direction, &hit)) continue;
            bool exiting;
            Vec<D> closer_normal;
            if (angle_between(direction, hit.normal) < FRAC_PI_2) {
                closer_normal = -hit.normal;
                exiting = true;
            } else {
                closer_normal = hit.normal;
                exiting = false;
            }
            if (!have || closest_distance > hit.distance) {
                out->hit_entity = e;
                out->direction = direction;
                out->hit = hit;
                out->normal_closer = closer_normal;
                out->exiting = exiting;
                closest_distance = hit.distance;
                have = true;
            }
        }
        return have;
    }

    // mod.rs:149-184 with ComposableSurface::get_color inlined (surface.rs:62-162)
    Rgba trace(uint32_t depth, int belongs_to, const Vec<D>& location, const Vec<D>& direction, int* primary_hit) {
        level_counts[max_depth - depth]++;
        Context c;
        c.origin_entity = belongs_to;
        if (depth > 0 && trace_closest(location, direction, &c)) {
            if (primary_hit) *primary_hit = c.hit_entity;
            const EuclSurface& sf = s.surfaces[s.entities[c.hit_entity].surface];
            real ratio = rust_max(rust_min(reflection_ratio(sf, c), R(1.0)), R(0.0));
            const Vec<D> offset = c.normal_closer; // used as (+-n * eps) * 128
            // get_intersection_color
            bool have_intersection = false;
            Rgba intersection_color{0, 0, 0, 0};
            if (!(ratio >= R(1.0))) {
                Rgba sc = surface_color(sf, c);
                uint8_t data[4];
                to_pixel4(sc, data, &counters);
                if (data[3] == 255) {
                    intersection_color = sc;
                    have_intersection = true;
                } else {
                    Vec<D> transitioned = threshold_direction(sf, c);
                    Vec<D> new_origin = c.hit.location + (-offset) * APPROX_EPSILON * R(128.0);
                    int destination = c.exiting ? material_at(new_origin) : c.hit_entity;
                    if (destination >= 0) {
                        material_exit(belongs_to, transitioned);
                        material_enter(destination, transitioned);
                        Rgba transition = trace(depth - 1, destination, new_origin, transitioned, nullptr);
                        uint8_t tdata[4];
                        to_pixel4(transition, tdata, &counters);
                        intersection_color =
                            from_premultiplied(blend_pre(EUCL_BLEND_OVER, into_premultiplied(new_u8(data)),
                                                         into_premultiplied(new_u8(tdata))));
                        have_intersection = true;
                    }
                }
            }
            // get_reflection_color
            bool have_reflection = false;
            Rgba reflection_color{0, 0, 0, 0};
            if (!(ratio <= R(0.0))) {
                Vec<D> rd = reflection_direction(c);
                Vec<D> new_origin = c.hit.location + offset * APPROX_EPSILON * R(128.0);
                reflection_color = trace(depth - 1, belongs_to, new_origin, rd, nullptr);
                have_reflection = true;
            }
            if (!have_intersection) {
                if (!have_reflection) { // reference: expect() panics; defined as transparent black
                    counters.no_material++;
                    return Rgba{0, 0, 0, 0};
                }
                return reflection_color;
            }
            if (!have_reflection) return intersection_color;
            return combine_palette_color(reflection_color, intersection_color, ratio);
        }
        if (primary_hit) *primary_hit = -1;
        return mapped_color(s.background, direction); // `direction.to_point()`
    }

    // Universe::trace_path (mod.rs:186-227) with Surface::get_path inlined (surface.rs:164-197)
    void trace_path(real distance, int belongs_to, const Vec<D>& location, const Vec<D>& direction, int budget,
                    Vec<D>* out_location, Vec<D>* out_direction, bool* runaway) {
        Context c;
        c.origin_entity = belongs_to;
        if (budget > 0 && trace_closest(location, direction, &c)) {
            if (!(distance - c.hit.distance <= R(0.0))) {
                real new_distance = distance - c.hit.distance;
                Vec<D> new_origin = c.hit.location + (-c.normal_closer) * APPROX_EPSILON * R(128.0);
                int destination = c.exiting ? material_at(new_origin) : c.hit_entity;
                if (destination >= 0) {
                    Vec<D> transitioned = c.direction;
                    material_exit(belongs_to, transitioned);
                    material_enter(destination, transitioned);
                    trace_path(new_distance, destination, new_origin, transitioned, budget - 1, out_location, out_direction,
                               runaway);
                    return;
                }
            }
        }
        if (budget <= 0) *runaway = true;
        // Material::trace_path (material.rs:54-56,144-146) then exit
        Vec<D> new_location = location + direction * distance;
        Vec<D> new_direction = direction;
        material_exit(belongs_to, new_direction);
        *out_location = new_location;
        *out_direction = new_direction;
    }
    // Universe::trace_path_unknown (mod.rs:273-286); false = None
    bool trace_path_unknown(real distance, const Vec<D>& location, const Vec<D>& direction, Vec<D>* out_location,
                            Vec<D>* out_direction, bool* runaway) {
        int belongs_to = material_at(location);
        if (belongs_to < 0) return false;
        Vec<D> transitioned = direction;
        material_enter(belongs_to, transitioned);
        trace_path(distance, belongs_to, location, transitioned, 4000, out_location, out_direction, runaway);
        return true;
    }

    // camera (d3/entity/camera.rs:164-185,369-390; d4/entity/camera.rs:155-176)
    Vec<D> ray_vector(int x, int y, int width, int height) const {
        real rel_x = (real)(x - width / 2) + (real)(1 - width % 2) / R(2.0);
        real rel_y = (real)(y - height / 2) + (real)(1 - height % 2) / R(2.0);
        real w = (real)width, h = (real)height;
        Vec<D> location = load<D>(cam.location), forward = load<D>(cam.forward), up = load<D>(cam.up), right;
        if (D == 3) {
            Vec<3> cr;
            cr[0] = R(cam.forward[1]) * R(cam.up[2]) - R(cam.forward[2]) * R(cam.up[1]);
            cr[1] = R(cam.forward[2]) * R(cam.up[0]) - R(cam.forward[0]) * R(cam.up[2]);
            cr[2] = R(cam.forward[0]) * R(cam.up[1]) - R(cam.forward[1]) * R(cam.up[0]);
            cr = normalize(cr);
            for (int k = 0; k < 3; ++k) right[k] = cr[k];
        } else {
            right = -load<D>(cam.left);
        }
        real fov_rad = PI * (real)cam.fov_deg / R(180.0);
        real distance = std::sqrt(w * w + h * h) / (R(2.0) * std::tan(fov_rad / R(2.0)));
        Vec<D> center = location + forward * distance;
        Vec<D> screen_point = center + (up * rel_y) + (right * rel_x);
        return normalize(screen_point - location);
    }

    // trace_screen_point + trace_unknown (mod.rs:253-271,371-397) + to_pixel (mod.rs:342)
    void pixel(int x, int y, int width, int height, uint8_t rgb[3], int32_t* hit_id) {
        Vec<D> point = load<D>(cam.location);
        Vec<D> vector = ray_vector(x, y, width, height);
        int belongs_to = material_at(point);
        real r, g, b;
        if (belongs_to >= 0) {
            Vec<D> transitioned = vector;
            material_enter(belongs_to, transitioned);
            int primary = -1;
            Rgba fg = trace(max_depth, belongs_to, point, transitioned, &primary);
            if (hit_id) *hit_id = primary;
            Pre over = blend_pre(EUCL_BLEND_OVER, into_premultiplied(fg), into_premultiplied(Rgba{R(1.0), R(1.0), R(1.0), R(1.0)}));
            Rgba out = from_premultiplied(over);
            r = out.r;
            g = out.g;
            b = out.b;
        } else {
            if (hit_id) *hit_id = -2;
            if ((x / 8 + y / 8) % 2 == 0) {
                r = g = b = R(0.0);
            } else {
                r = R(1.0);
                g = R(0.0);
                b = R(1.0);
            }
        }
        rgb[0] = channel_to_u8(r, &counters);
        rgb[1] = channel_to_u8(g, &counters);
        rgb[2] = channel_to_u8(b, &counters);
    }
};

template <int D>
int render_impl(const EuclFlatScene* scene, const EuclCamera* camera, uint32_t width, uint32_t height, double time,
                uint32_t row_begin, uint32_t row_end, int threads, uint8_t* out_rgb, int32_t* out_hit, uint64_t* stats) {
    if (threads < 1) threads = 1;
    // dynamic scheduling over tiles of 128 pixels (the reference hands out one job per pixel,
    // mod.rs:316-318; tiles keep every host thread busy even for a few sampled rows)
    const uint64_t total = (uint64_t)(row_end - row_begin) * width;
    const uint64_t tile = 128;
    std::atomic<uint64_t> next_tile{0};
    std::vector<Tracer<D>> tracers;
    tracers.reserve((size_t)threads);
    for (int t = 0; t < threads; ++t) tracers.emplace_back(*scene, *camera, time);
#ifdef ORACLE_COUNT_FLOPS
    std::atomic<uint64_t> flops_total{0};
#endif
    auto work = [&](int t) {
#ifdef ORACLE_COUNT_FLOPS
        g_flops = 0;
        struct Flush {
            std::atomic<uint64_t>& total;
            ~Flush() { total += g_flops; }
        } flush{flops_total};
#endif
        Tracer<D>& tr = tracers[(size_t)t];
        for (;;) {
            const uint64_t begin = next_tile.fetch_add(tile);
            if (begin >= total) break;
            const uint64_t end = begin + tile < total ? begin + tile : total;
            for (uint64_t idx = begin; idx < end; ++idx) {
                const uint32_t y = row_begin + (uint32_t)(idx / width), x = (uint32_t)(idx % width);
                tr.pixel((int)x, (int)y, (int)width, (int)height, out_rgb + 3 * idx, out_hit ? out_hit + idx : nullptr);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (stats) {
        // stats[0] = segments, [1] = nodes, [2..5] = undefined-corner counters, [6] = counted flops (counting build), [8 + l] = level l
        std::memset(stats, 0, sizeof(uint64_t) * (8 + EUCL_MAX_LEVELS));
        for (auto& tr : tracers) {
            for (uint32_t l = 0; l <= camera->max_depth && l < EUCL_MAX_LEVELS; ++l) {
                stats[8 + l] += tr.level_counts[l];
                stats[1] += tr.level_counts[l];
                if (l < camera->max_depth) stats[0] += tr.level_counts[l];
            }
            stats[2] += tr.counters.nan_channel;
            stats[3] += tr.counters.bad_texcoord;
            stats[4] += tr.counters.no_material;
            stats[5] += tr.counters.csg_runaway;
        }
#ifdef ORACLE_COUNT_FLOPS
        stats[6] = flops_total.load(); // counted floating-point operations of this call (see FLOPS)
#endif
    }
    return 0;
}

} // namespace

// BEGIN_KEEP64
extern "C" {

// Renders rows [row_begin, row_end) of a width x height frame (row 0 = bottom) into out_rgb
// (3 bytes per pixel, rows packed from row_begin) -- Environment::render, mod.rs:300-357.
int oracle_render(const EuclFlatScene* scene, const EuclCamera* camera, uint32_t width, uint32_t height, double time,
                  uint32_t row_begin, uint32_t row_end, int threads, uint8_t* out_rgb, int32_t* out_hit,
                  uint64_t* stats) {
    if (!scene || !camera || !out_rgb || camera->max_depth + 1 > EUCL_MAX_LEVELS) return -1;
    if (scene->dim == 3)
        return render_impl<3>(scene, camera, width, height, time, row_begin, row_end, threads, out_rgb, out_hit, stats);
    if (scene->dim == 4)
        return render_impl<4>(scene, camera, width, height, time, row_begin, row_end, threads, out_rgb, out_hit, stats);
    return -1;
}

#ifndef ORACLE_F32 // the unit-test hooks probe the f64 build; liboracle_f32.so exports the renderer and the camera path only
// --- unit-test hooks -------------------------------------------------------------------------

// Full intersection stream of entity `entity` (up to the first None): out = n x (distance,
// location[dim], normal[dim]).
int oracle_entity_intersections(const EuclFlatScene* scene, int entity, const double* location, const double* direction,
                                int max_items, double* out) {
    EuclCamera cam{};
    cam.max_depth = 1;
    if (scene->dim == 3) {
        Tracer<3> tr(*scene, cam, 0.0);
        std::vector<Hit<3>> hits((size_t)max_items);
        int n = tr.all_intersections(scene->entities[entity], load<3>(location), load<3>(direction), max_items, hits.data());
        for (int i = 0; i < n; ++i) {
            out[i * 7] = hits[(size_t)i].distance;
            for (int k = 0; k < 3; ++k) {
                out[i * 7 + 1 + k] = hits[(size_t)i].location[k];
                out[i * 7 + 4 + k] = hits[(size_t)i].normal[k];
            }
        }
        return n;
    }
    Tracer<4> tr(*scene, cam, 0.0);
    std::vector<Hit<4>> hits((size_t)max_items);
    int n = tr.all_intersections(scene->entities[entity], load<4>(location), load<4>(direction), max_items, hits.data());
    for (int i = 0; i < n; ++i) {
        out[i * 9] = hits[(size_t)i].distance;
        for (int k = 0; k < 4; ++k) {
            out[i * 9 + 1 + k] = hits[(size_t)i].location[k];
            out[i * 9 + 5 + k] = hits[(size_t)i].normal[k];
        }
    }
    return n;
}

// One primitive against one ray in 2, 3 or 4 dimensions (the reference's KATs are 2-D,
// shape.rs:1048-1148): out = n x (distance, location[dim], normal[dim]); returns n.
int oracle_prim_intersect(int dim, const EuclPrim* prim, const double* location, const double* direction, double* out) {
    auto run = [&](auto tag) {
        constexpr int D = decltype(tag)::value;
        Hit<D> hits[2];
        int n = intersect_prim<D>(*prim, load<D>(location), load<D>(direction), hits);
        for (int i = 0; i < n; ++i) {
            out[i * (1 + 2 * D)] = hits[i].distance;
            for (int k = 0; k < D; ++k) {
                out[i * (1 + 2 * D) + 1 + k] = hits[i].location[k];
                out[i * (1 + 2 * D) + 1 + D + k] = hits[i].normal[k];
            }
        }
        return n;
    };
    if (dim == 2) return run(std::integral_constant<int, 2>{});
    if (dim == 3) return run(std::integral_constant<int, 3>{});
    return run(std::integral_constant<int, 4>{});
}
int oracle_prim_inside(int dim, const EuclPrim* prim, const double* point) {
    if (dim == 2) return prim_inside<2>(*prim, load<2>(point));
    if (dim == 3) return prim_inside<3>(*prim, load<3>(point));
    return prim_inside<4>(*prim, load<4>(point));
}

#endif

// Universe::trace_path_unknown: returns 0 ok, 1 None (start point in no entity), -1 runaway
int oracle_trace_path(const EuclFlatScene* scene, const double* location, const double* direction, double distance,
                      double* out_location, double* out_direction) {
    EuclCamera cam{};
    bool runaway = false, ok;
    if (scene->dim == 3) {
        Tracer<3> tr(*scene, cam, 0.0);
        Vec<3> l, d;
        ok = tr.trace_path_unknown((real)distance, load<3>(location), load<3>(direction), &l, &d, &runaway);
        for (int k = 0; ok && k < 3; ++k) {
            out_location[k] = (double)l[k];
            out_direction[k] = (double)d[k];
        }
    } else {
        Tracer<4> tr(*scene, cam, 0.0);
        Vec<4> l, d;
        ok = tr.trace_path_unknown((real)distance, load<4>(location), load<4>(direction), &l, &d, &runaway);
        for (int k = 0; ok && k < 4; ++k) {
            out_location[k] = (double)l[k];
            out_direction[k] = (double)d[k];
        }
    }
    if (runaway) return -1;
    return ok ? 0 : 1;
}

#ifndef ORACLE_F32
int oracle_entity_inside(const EuclFlatScene* scene, int entity, const double* point) {
    EuclCamera cam{};
    if (scene->dim == 3) return Tracer<3>(*scene, cam, 0.0).node_inside(scene->entities[entity].node_root, load<3>(point));
    return Tracer<4>(*scene, cam, 0.0).node_inside(scene->entities[entity].node_root, load<4>(point));
}

int oracle_material_at(const EuclFlatScene* scene, const double* point) {
    EuclCamera cam{};
    if (scene->dim == 3) return Tracer<3>(*scene, cam, 0.0).material_at(load<3>(point));
    return Tracer<4>(*scene, cam, 0.0).material_at(load<4>(point));
}

double oracle_angle_between(int dim, const double* a, const double* b) {
    if (dim == 2) return angle_between(load<2>(a), load<2>(b));
    if (dim == 3) return angle_between(load<3>(a), load<3>(b));
    return angle_between(load<4>(a), load<4>(b));
}
// f32 variant of util.rs:712-722 for the reference's own f32 test (util.rs:1007-1037)
float oracle_angle_between_f32(const float* a, const float* b) {
    float d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    float na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    float nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    float r = std::acos(d / (na * nb));
    return std::isnan(r) ? 0.0f : r;
}

void oracle_combine_palette_color(const double* a, const double* b, double ratio, double* out) {
    Rgba r = combine_palette_color(Rgba{a[0], a[1], a[2], a[3]}, Rgba{b[0], b[1], b[2], b[3]}, ratio);
    out[0] = r.r;
    out[1] = r.g;
    out[2] = r.b;
    out[3] = r.a;
}
// f32 variant for the reference's own test (util.rs:947-958 runs with f32 literals)
void oracle_combine_palette_color_f32(const float* a, const float* b, float ratio, float* out) {
    for (int k = 0; k < 4; ++k) out[k] = ratio <= 0.0f ? b[k] : ratio >= 1.0f ? a[k] : a[k] * ratio + b[k] * (1.0f - ratio);
}

int64_t oracle_remainder_i(int64_t a, int64_t b) { return remainder_i(a, b); }
double oracle_remainder_f(double a, double b) { return remainder_f(a, b); }

void oracle_blend(int fn, double ratio, const double* src, const double* dst, double* out) {
    Rgba r = blend_rgba(fn, ratio, Rgba{src[0], src[1], src[2], src[3]}, Rgba{dst[0], dst[1], dst[2], dst[3]});
    out[0] = r.r;
    out[1] = r.g;
    out[2] = r.b;
    out[3] = r.a;
}

void oracle_to_pixel(const double* rgba, uint8_t* out) {
    to_pixel4(Rgba{rgba[0], rgba[1], rgba[2], rgba[3]}, out, nullptr);
}

double oracle_perlin4(const uint8_t* perm, const double* point) { return perlin4(perm, point); }

// eucl_detmath functions (present in both builds), fn: 0 acos, 1 asin, 2 sin, 3 cos, 4 atan
void oracle_detmath_unary(int fn, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (fn) {
        case 0: out[i] = eucl_det::det_acos(x[i]); break;
        case 1: out[i] = eucl_det::det_asin(x[i]); break;
        case 2: out[i] = eucl_det::det_sin(x[i]); break;
        case 3: out[i] = eucl_det::det_cos(x[i]); break;
        default: out[i] = eucl_det::det_atan(x[i]); break;
        }
    }
}
void oracle_detmath_atan2(const double* y, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = eucl_det::det_atan2(y[i], x[i]);
}
#endif
int oracle_uses_detmath(void) {
#ifdef ORACLE_DETMATH
    return 1;
#else
    return 0;
#endif
}
// bytes of the scalar type this build traces in (8: f64, 4: the reference's `low_precision` f32)
int oracle_real_bytes(void) { return (int)sizeof(real); }
#ifndef ORACLE_F32

void oracle_hsv_to_rgb(double h, double s, double v, double* rgb) { hsv_to_rgb(h, s, v, rgb); }

void oracle_general_rotation(int dim, const double* self, const double* other, double angle, const double* v, double* out) {
    EuclFlatScene dummy{};
    EuclCamera cam{};
    if (dim == 3) {
        Vec<3> r = Tracer<3>(dummy, cam, 0.0).general_rotation(load<3>(self), load<3>(other), angle, load<3>(v));
        for (int k = 0; k < 3; ++k) out[k] = r[k];
    } else {
        Vec<4> r = Tracer<4>(dummy, cam, 0.0).general_rotation(load<4>(self), load<4>(other), angle, load<4>(v));
        for (int k = 0; k < 4; ++k) out[k] = r[k];
    }
}

// reflection ratio / directions of surface `surface` for a synthetic hit
void oracle_surface_probe(const EuclFlatScene* scene, int surface, const double* direction, const double* normal_closer,
                          int exiting, double* ratio, double* reflect_dir, double* threshold_dir) {
    EuclCamera cam{};
    const EuclSurface& sf = scene->surfaces[surface];
    if (scene->dim == 3) {
        Tracer<3> tr(*scene, cam, 0.0);
        Tracer<3>::Context c{};
        c.direction = load<3>(direction);
        c.normal_closer = load<3>(normal_closer);
        c.exiting = exiting != 0;
        *ratio = tr.reflection_ratio(sf, c);
        Vec<3> r = tr.reflection_direction(c), t = tr.threshold_direction(sf, c);
        for (int k = 0; k < 3; ++k) {
            reflect_dir[k] = r[k];
            threshold_dir[k] = t[k];
        }
    } else {
        Tracer<4> tr(*scene, cam, 0.0);
        Tracer<4>::Context c{};
        c.direction = load<4>(direction);
        c.normal_closer = load<4>(normal_closer);
        c.exiting = exiting != 0;
        *ratio = tr.reflection_ratio(sf, c);
        Vec<4> r = tr.reflection_direction(c), t = tr.threshold_direction(sf, c);
        for (int k = 0; k < 4; ++k) {
            reflect_dir[k] = r[k];
            threshold_dir[k] = t[k];
        }
    }
}

void oracle_mapped_color(const EuclFlatScene* scene, int mapped, const double* point, double* out) {
    EuclCamera cam{};
    Rgba r;
    if (scene->dim == 3) r = Tracer<3>(*scene, cam, 0.0).mapped_color(mapped, load<3>(point));
    else r = Tracer<4>(*scene, cam, 0.0).mapped_color(mapped, load<4>(point));
    out[0] = r.r;
    out[1] = r.g;
    out[2] = r.b;
    out[3] = r.a;
}

void oracle_ray_vector(const EuclFlatScene* scene, const EuclCamera* cam, int x, int y, int w, int h, double* out) {
    if (scene->dim == 3) {
        Vec<3> r = Tracer<3>(*scene, *cam, 0.0).ray_vector(x, y, w, h);
        for (int k = 0; k < 3; ++k) out[k] = r[k];
    } else {
        Vec<4> r = Tracer<4>(*scene, *cam, 0.0).ray_vector(x, y, w, h);
        for (int k = 0; k < 4; ++k) out[k] = r[k];
    }
}

#endif
} // extern "C"
// END_KEEP64
