// zro_math.h — ORACLE-side restatement of the "spec" transcendental kernels (DESIGN.md "Spec math").
// TEST INFRASTRUCTURE ONLY.  The reference calls Zig std.math.{sin,cos,acos,atan2,pow}; that library
// is not under /root/reference and no reference test pins its results to the ulp.  Any implementation
// that is accurate to ~1 ulp is an equally valid restatement, so the product fixes ONE such
// implementation (f32 only, + - * / sqrt, explicit evaluation order, no FMA) and the oracle restates it
// here independently, which makes GPU and oracle comparable draw for draw, bit for bit.
// ZRO_MATH_LIBM selects glibc instead (the independent cross-check used by the statistical tests).
//   sincos : Cephes single-precision sinf/cosf scheme (3-term Cody-Waite pi/4 reduction, degree-3 polys in z^2)
//   acos   : FreeBSD/musl e_acosf.c scheme (rational R(z) = p/q, sqrt split for |x| > 0.5)
//   atan2  : FreeBSD/musl s_atanf.c + e_atan2f.c scheme (4 reduction ranges, hi/lo constants)
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

inline float acos_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f, qS1 = -7.0662963390e-01f;
    const float p = z * (pS0 + z * (pS1 + z * pS2));
    const float q = 1.0f + z * qS1;
    return p / q;
}
inline float acos(float x) {
    const float pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const uint32_t hx = f2u(x);
    const uint32_t ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) { // |x| >= 1 or NaN
        if (ix == 0x3f800000u) {
            if (hx >> 31) return 2 * pio2_hi + 7.5231638452626401e-37f; // 0x1p-120f
            return 0.0f;
        }
        return u2f(0x7fc00000u);
    }
    if (ix < 0x3f000000u) { // |x| < 0.5
        if (ix <= 0x32800000u) return pio2_hi + 7.5231638452626401e-37f; // |x| < 2^-26
        return pio2_hi - (x - (pio2_lo - x * acos_R(x * x)));
    }
    if (hx >> 31) { // x < -0.5
        const float z = (1 + x) * 0.5f;
        const float s = std::sqrt(z);
        const float w = acos_R(z) * s - pio2_lo;
        return 2 * (pio2_hi - (s + w));
    }
    const float z = (1 - x) * 0.5f; // x > 0.5
    const float s = std::sqrt(z);
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float w = acos_R(z) * s + c;
    return 2 * (df + w);
}

inline float atan(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[5] = {3.3333328366e-01f, -1.9999158382e-01f, 1.4253635705e-01f, -1.0648017377e-01f, 6.1687607318e-02f};
    uint32_t ix = f2u(x);
    const uint32_t sign = ix >> 31;
    ix &= 0x7fffffffu;
    int id;
    if (ix >= 0x4c800000u) { // |x| >= 2^26
        if (ix > 0x7f800000u) return x; // NaN
        const float z = atanhi[3] + 7.5231638452626401e-37f;
        return sign ? -z : z;
    }
    if (ix < 0x3ee00000u) { // |x| < 0.4375
        if (ix < 0x39800000u) return x; // |x| < 2^-12
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3f980000u) {     // |x| < 1.1875
            if (ix < 0x3f300000u) { // 7/16 <= |x| < 11/16
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else { // 11/16 <= |x| < 19/16
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000u) { // |x| < 2.4375
                id = 2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else { // 2.4375 <= |x| < 2^26
                id = 3;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * aT[4]));
    const float s2 = w * (aT[1] + w * aT[3]);
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return sign ? -r : r;
}

inline float atan2(float y, float x) {
    const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    if (std::isnan(x) || std::isnan(y)) return x + y;
    uint32_t ix = f2u(x), iy = f2u(y);
    if (ix == 0x3f800000u) return atan(y); // x = 1.0
    const uint32_t m = ((iy >> 31) & 1) | ((ix >> 30) & 2); // 2*sign(x) + sign(y)
    ix &= 0x7fffffffu;
    iy &= 0x7fffffffu;
    if (iy == 0) { // y = 0
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi;
        default: return -pi;
        }
    }
    if (ix == 0) return (m & 1) ? -pi / 2 : pi / 2; // x = 0
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
            case 0: return pi / 4;
            case 1: return -pi / 4;
            case 2: return 3 * pi / 4;
            default: return -3 * pi / 4;
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi;
        default: return -pi;
        }
    }
    if (ix + (26u << 23) < iy || iy == 0x7f800000u) return (m & 1) ? -pi / 2 : pi / 2; // |y/x| > 2^26
    float z;
    if ((m & 2) && iy + (26u << 23) < ix) z = 0.0f; // |y/x| < 2^-26, x < 0
    else z = atan(std::fabs(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

inline float pow5(float x) {
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x * x4;
}

} // namespace zro_math
#endif
