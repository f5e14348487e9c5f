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
//   acos   : Abramowitz & Stegun 4.4.46: sqrt(1-|x|) * P7(|x|), reflected for x < 0.  Branch-free.
//   atan2  : octant folding t = min/max in [0,1], Abramowitz & Stegun 4.4.49 odd polynomial, three
//            reflections.  Branch-free (it runs inside a divergent region).
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

// ---- exact IEEE division helpers ---------------------------------------------------------------------
// a / b correctly rounded (round-to-nearest-even), given y = RN(1/b): two FMA residual corrections
// (Markstein).  Preconditions (callers guarantee them): b finite and normal, a == 0 or 2^-60 <= |a| and the
// quotient far from overflow/underflow, so that both residuals are exact.  The sign of a zero quotient
// is the IEEE one.  tests/ check it bit for bit against the compiler's own division.
__device__ __forceinline__ float div_exact(float a, float b, float y) {
    const float q0 = a * y;
    const float r0 = __fmaf_rn(-b, q0, a);
    const float q1 = __fmaf_rn(r0, y, q0);
    const float r1 = __fmaf_rn(-b, q1, a);
    const float q2 = __fmaf_rn(r1, y, q1);
    return (a == 0.0f) ? q0 : q2;
}

__device__ __forceinline__ float acos_spec(float x) {
    const float ax = fabsf(x);
    float p = -0.0012624911f;
    p = p * ax + 0.0066700901f;
    p = p * ax + -0.0170881256f;
    p = p * ax + 0.0308918810f;
    p = p * ax + -0.0501743046f;
    p = p * ax + 0.0889789874f;
    p = p * ax + -0.2145988016f;
    p = p * ax + 1.5707963050f;
    const float r = sqrtf(1.0f - ax) * p;
    return (x < 0.0f) ? 3.14159265358979323846f - r : r;
}

__device__ __forceinline__ float atan01_spec(float t) {
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

__device__ __forceinline__ float atan2_spec(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = (ax > ay) ? ax : ay, mn = (ax > ay) ? ay : ax;
    const float t = (mx == 0.0f) ? 0.0f : mn / mx;
    float p = atan01_spec(t);
    if (ay > ax) p = 1.57079632679489661923f - p;
    if (x < 0.0f) p = 3.14159265358979323846f - p;
    if (y < 0.0f) p = -p;
    return p;
}

__device__ __forceinline__ float pow5_spec(float x) {
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x * x4;
}

} // namespace dmath
} // namespace zrt
