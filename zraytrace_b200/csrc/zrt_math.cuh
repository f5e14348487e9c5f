// zrt_math.cuh — the "spec" transcendental kernels of the device path (DESIGN.md "Spec math").
//
// The reference calls Zig std.math.{sin,cos,acos,atan2,pow} (sample.zig:50-52, sphere.zig:47-48,
// material.zig:127); that library is outside the reference tree and no reference test pins its
// results to the last bit.  The device path therefore fixes one ~1-ulp implementation per function,
// written with IEEE f32 + - * / sqrt only, in a fixed evaluation order, and the whole translation
// unit is compiled with -fmad=false so that nothing is contracted.  The CPU oracle restates the same
// kernels independently (oracle/zro_math.h), which is what makes GPU paths comparable with oracle
// paths draw for draw; tests/ also bound every kernel against float64 truth (<= 3 ulp).
//
//   sincos : Cephes single-precision scheme: j = trunc(x * 4/pi) rounded up to even, 3-term Cody-Waite
//            reduction by j*pi/4, degree-3 polynomials in z^2, octant fix-up.  Domain 0 <= x <= 8192.
//   acos   : FreeBSD msun e_acosf.c scheme: rational R(z)=p/q; sqrt split for |x| > 0.5.
//   atan2  : FreeBSD msun s_atanf.c / e_atan2f.c scheme: four reduction ranges with hi/lo constants.
//   pow5   : Zig's std.math.pow evaluates an integer power by square-and-multiply on the mantissa with
//            exact power-of-two scaling, i.e. x^5 = x * ((x*x) * (x*x)) in plain f32 products.
#pragma once
#include <cuda_runtime.h>

namespace zrt {
namespace dmath {

__device__ __forceinline__ void sincos_spec(float x, float *s_out, float *c_out) {
    int j = __float2int_rz(x * 1.27323954473516f); // 4/pi
    j += (j & 1);
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
    const int q = (j >> 1) & 3;
    const float a = (q & 1) ? pc : ps; // |sin|
    const float b = (q & 1) ? ps : pc; // |cos|
    *s_out = (q & 2) ? -a : a;
    *c_out = (q == 1 || q == 2) ? -b : b;
}

__device__ __forceinline__ float acos_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f, qS1 = -7.0662963390e-01f;
    const float p = z * (pS0 + z * (pS1 + z * pS2));
    const float q = 1.0f + z * qS1;
    return p / q;
}

__device__ __forceinline__ float acos_spec(float x) {
    const float pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const unsigned hx = __float_as_uint(x);
    const unsigned ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) {
        if (ix == 0x3f800000u) {
            if (hx >> 31) return 2 * pio2_hi + 7.5231638452626401e-37f;
            return 0.0f;
        }
        return __uint_as_float(0x7fc00000u);
    }
    if (ix < 0x3f000000u) {
        if (ix <= 0x32800000u) return pio2_hi + 7.5231638452626401e-37f;
        return pio2_hi - (x - (pio2_lo - x * acos_R(x * x)));
    }
    if (hx >> 31) {
        const float z = (1 + x) * 0.5f;
        const float s = sqrtf(z);
        const float w = acos_R(z) * s - pio2_lo;
        return 2 * (pio2_hi - (s + w));
    }
    const float z = (1 - x) * 0.5f;
    const float s = sqrtf(z);
    const float df = __uint_as_float(__float_as_uint(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float w = acos_R(z) * s + c;
    return 2 * (df + w);
}

__device__ __forceinline__ float atan_spec(float x) {
    const float aT0 = 3.3333328366e-01f, aT1 = -1.9999158382e-01f, aT2 = 1.4253635705e-01f, aT3 = -1.0648017377e-01f,
                aT4 = 6.1687607318e-02f;
    unsigned ix = __float_as_uint(x);
    const unsigned sign = ix >> 31;
    ix &= 0x7fffffffu;
    float hi = 0.0f, lo = 0.0f;
    bool reduced = true;
    if (ix >= 0x4c800000u) {
        if (ix > 0x7f800000u) return x;
        const float z = 1.5707962513e+00f + 7.5231638452626401e-37f;
        return sign ? -z : z;
    }
    if (ix < 0x3ee00000u) {
        if (ix < 0x39800000u) return x;
        reduced = false;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) {
                hi = 4.6364760399e-01f; lo = 5.0121582440e-09f;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                hi = 7.8539812565e-01f; lo = 3.7748947079e-08f;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000u) {
                hi = 9.8279368877e-01f; lo = 3.4473217170e-08f;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                hi = 1.5707962513e+00f; lo = 7.5497894159e-08f;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * aT4));
    const float s2 = w * (aT1 + w * aT3);
    if (!reduced) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return sign ? -r : r;
}

__device__ __forceinline__ float atan2_spec(float y, float x) {
    const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    if (isnan(x) || isnan(y)) return x + y;
    unsigned ix = __float_as_uint(x), iy = __float_as_uint(y);
    if (ix == 0x3f800000u) return atan_spec(y);
    const unsigned m = ((iy >> 31) & 1) | ((ix >> 30) & 2);
    ix &= 0x7fffffffu;
    iy &= 0x7fffffffu;
    if (iy == 0) return (m == 0 || m == 1) ? y : (m == 2 ? pi : -pi);
    if (ix == 0) return (m & 1) ? -pi / 2 : pi / 2;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) return m == 0 ? pi / 4 : (m == 1 ? -pi / 4 : (m == 2 ? 3 * pi / 4 : -3 * pi / 4));
        return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi : -pi));
    }
    if (ix + (26u << 23) < iy || iy == 0x7f800000u) return (m & 1) ? -pi / 2 : pi / 2;
    float z;
    if ((m & 2) && iy + (26u << 23) < ix) z = 0.0f;
    else z = atan_spec(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

__device__ __forceinline__ float pow5_spec(float x) {
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x * x4;
}

} // namespace dmath
} // namespace zrt
