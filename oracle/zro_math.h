// zro_math.h — ORACLE-side restatement of the "spec" transcendental kernels (DESIGN.md "Spec math").
// TEST INFRASTRUCTURE ONLY.  The reference calls Zig std.math.{sin,cos,acos,atan2,pow}; that library
// is not under /root/reference and no reference test pins its results to the ulp.  Any implementation
// that is accurate to ~1 ulp is an equally valid restatement, so the product fixes ONE such
// implementation (f32 only, + - * / sqrt, explicit evaluation order, no FMA) and the oracle restates it
// here independently, which makes GPU and oracle comparable draw for draw, bit for bit.
// ZRO_MATH_LIBM selects glibc instead (the independent cross-check used by the statistical tests).
//   sincos : Cephes single-precision sinf/cosf scheme (3-term Cody-Waite pi/4 reduction, degree-3 polys in z^2)
//   acos   : Abramowitz & Stegun 4.4.46 polynomial times sqrt(1-|x|), reflected for x < 0
//   atan2  : octant folding + Abramowitz & Stegun 4.4.49 odd polynomial on [0,1]
//   pow5   : Zig std.math.pow for an integer exponent is square-and-multiply on the mantissa
//            (x^5 = x * ((x*x)*(x*x))); powers of two scale exactly, so plain f32 products restate it.
#ifndef ZRO_MATH_H
#define ZRO_MATH_H
#include <cmath>
#include <cstdint>
#include <cstring>

namespace zro_math {

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// valid for 0 <= x <= 8192 (the path only needs [0, 2*pi])
inline void sincos(float x, float *s_out, float *c_out) {
    int j = (int)(x * 1.27323954473516f); // 4/pi
    if (j & 1) j += 1;
    const float y = (float)j;
    const float z = ((x - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    const float w = z * z;
    float ps = (-1.9515295891e-4f * w + 8.3321608736e-3f) * w - 1.6666654611e-1f;
    ps = ps * w;
    ps = ps * z;
    ps = ps + z;
    float pc = (2.443315711809948e-5f * w - 1.388731625493765e-3f) * w + 4.166664568298827e-2f;
    pc = pc * w;
    pc = pc * w;
    pc = pc - 0.5f * w;
    pc = pc + 1.0f;
    switch ((j >> 1) & 3) {
    case 0: *s_out = ps; *c_out = pc; break;
    case 1: *s_out = pc; *c_out = -ps; break;
    case 2: *s_out = -ps; *c_out = -pc; break;
    default: *s_out = -pc; *c_out = ps; break;
    }
}

// acos: Abramowitz & Stegun 4.4.46 (|error| <= 2e-8): for 0 <= x <= 1
//   acos(x) = sqrt(1 - x) * (a0 + a1 x + ... + a7 x^7), and acos(-x) = pi - acos(x).
// Branch-free on purpose: the device evaluates it inside a divergent region.
inline float acos(float x) {
    const float ax = std::fabs(x);
    float p = -0.0012624911f;
    p = p * ax + 0.0066700901f;
    p = p * ax + -0.0170881256f;
    p = p * ax + 0.0308918810f;
    p = p * ax + -0.0501743046f;
    p = p * ax + 0.0889789874f;
    p = p * ax + -0.2145988016f;
    p = p * ax + 1.5707963050f;
    const float r = std::sqrt(1.0f - ax) * p; // NaN for |x| > 1
    return (x < 0.0f) ? 3.14159265358979323846f - r : r;
}

// atan on [0, 1]: Abramowitz & Stegun 4.4.49 (|error| <= 2e-8), odd polynomial up to t^17
inline float atan01(float t) {
    const float s = t * t;
    float p = 0.0028662257f;
    p = p * s + -0.0161657367f;
    p = p * s + 0.0429096138f;
    p = p * s + -0.0752896400f;
    p = p * s + 0.1065626393f;
    p = p * s + -0.1420889944f;
    p = p * s + 0.1999355085f;
    p = p * s + -0.3333314528f;
    p = p * s + 1.0f;
    return p * t;
}

// atan2 by octant folding: t = min(|x|,|y|) / max(|x|,|y|) in [0,1], then three reflections
inline float atan2(float y, float x) {
    const float ax = std::fabs(x), ay = std::fabs(y);
    const float mx = (ax > ay) ? ax : ay, mn = (ax > ay) ? ay : ax;
    const float t = (mx == 0.0f) ? 0.0f : mn / mx;
    float p = atan01(t);
    if (ay > ax) p = 1.57079632679489661923f - p;
    if (x < 0.0f) p = 3.14159265358979323846f - p;
    if (y < 0.0f) p = -p;
    return p;
}

inline float pow5(float x) {
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x * x4;
}

} // namespace zro_math
#endif
