// zro.cpp — CPU ORACLE for the zraytrace hot path.  TEST INFRASTRUCTURE ONLY (see zro.h).
//
// A restatement, function by function, of the reference's single-threaded CPU path
// (jsyrjala/zraytrace src/*.zig).  It keeps the reference's data structures on purpose (a
// pointer-linked BVH of tagged surfaces, recursion in rayColor, one shared sequential PRNG) and every
// quirk listed in SURVEY.md Appendix A.  Build with -ffp-contract=off -fno-fast-math: Zig's default
// float mode is strict (no FMA fusion, no reassociation).
//
// Third-party arithmetic that is not under /root/reference and is restated here from its published
// algorithm: Zig std.rand (pre-0.8 DefaultPrng = Xoroshiro128+ seeded by SplitMix64, Random.float /
// Random.boolean bit recipes; pinned by the sample.zig:70-118 goldens), std.sort.sort (stable ->
// std::stable_sort), std.math.min/max (`if (x < y) x else y` / `if (x > y) x else y`).
// Transcendentals (sin cos acos atan2 pow tan) come from glibc libm; no reference test pins them.
#include "zro.h"
#include "zro_math.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

using f32 = float; // base.zig:2

// ---------------------------------------------------------------- std.math.min / max / clamp
inline f32 zmin(f32 x, f32 y) { return (x < y) ? x : y; }
inline f32 zmax(f32 x, f32 y) { return (x > y) ? x : y; }

// ---------------------------------------------------------------- vector.zig:22-139
struct Vec3 {
    f32 x, y, z;
    f32 elem(int i) const { return i == 0 ? x : (i == 1 ? y : z); } // vector.zig:32-39
};
inline Vec3 v3(f32 x, f32 y, f32 z) { return Vec3{x, y, z}; }
inline Vec3 v3(const float *p) { return Vec3{p[0], p[1], p[2]}; }
inline Vec3 v3(const zrt_vec3 &p) { return Vec3{p.x, p.y, p.z}; }
inline f32 dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // :65
inline Vec3 cross(Vec3 u, Vec3 v) {                                          // :70-74
    return v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
inline f32 lengthSquared(Vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; } // :76
inline f32 length(Vec3 a) { return std::sqrt(lengthSquared(a)); }              // :80
inline Vec3 unitVector(Vec3 v) {                                               // :88-92
    const f32 len = length(v);
    return v3(v.x / len, v.y / len, v.z / len);
}
inline Vec3 negate(Vec3 a) { return v3(-a.x, -a.y, -a.z); }
inline Vec3 plus(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 minus(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 scale(Vec3 a, f32 s) { return v3(a.x * s, a.y * s, a.z * s); }
inline Vec3 reflect(Vec3 v, Vec3 n) { return minus(v, scale(n, 2 * dot(v, n))); } // :129-131
inline Vec3 refract(Vec3 v, Vec3 n, f32 ratio) {                                  // :134-139
    const f32 cos_theta = zmin(dot(negate(v), n), 1.0f);
    const Vec3 r_out_perp = scale(plus(v, scale(n, cos_theta)), ratio);
    const Vec3 r_out_parallel = scale(n, -std::sqrt(std::fabs(1.0f - lengthSquared(r_out_perp))));
    return plus(r_out_perp, r_out_parallel);
}
inline Vec3 center(const Vec3 *v, size_t n) { // vector.zig:149-160
    const f32 len_scale = 1.0f / (f32)n;
    f32 xs = 0, ys = 0, zs = 0;
    for (size_t i = 0; i < n; i++) {
        xs += v[i].x * len_scale;
        ys += v[i].y * len_scale;
        zs += v[i].z * len_scale;
    }
    return v3(xs, ys, zs);
}
struct Vec2 { f32 u, v; };

// ---------------------------------------------------------------- image.zig:9-72
struct Color { f32 r, g, b; };
inline Color cscale(Color c, f32 t) { return Color{c.r * t, c.g * t, c.b * t}; }
inline Color cmul(Color a, Color b) { return Color{a.r * b.r, a.g * b.g, a.b * b.b}; }
inline Color cadd(Color a, Color b) { return Color{a.r + b.r, a.g + b.g, a.b + b.b}; }

// ---------------------------------------------------------------- ray.zig:7-16
struct Ray {
    Vec3 origin, direction;
};
inline Ray rayInit(Vec3 o, Vec3 d) { return Ray{o, unitVector(d)}; } // ray.zig:11-13 normalises
inline Vec3 rayAt(const Ray &r, f32 t) { return plus(r.origin, scale(r.direction, t)); }

// ---------------------------------------------------------------- aabb.zig:10-127
struct AABB {
    Vec3 min, max, midpoint;
};
inline Vec3 minimumVec(Vec3 a, Vec3 b) { return v3(zmin(a.x, b.x), zmin(a.y, b.y), zmin(a.z, b.z)); }
inline Vec3 maximumVec(Vec3 a, Vec3 b) { return v3(zmax(a.x, b.x), zmax(a.y, b.y), zmax(a.z, b.z)); }
inline Vec3 midpointVec(Vec3 a, Vec3 b) {
    return v3((a.x + b.x) / 2.0f, (a.y + b.y) / 2.0f, (a.z + b.z) / 2.0f);
}
inline AABB aabbMinMax(Vec3 c1, Vec3 c2) { // aabb.zig:37-41
    return AABB{minimumVec(c1, c2), maximumVec(c1, c2), midpointVec(c1, c2)};
}
inline AABB aabbVertexes(const Vec3 *v, size_t n) { // aabb.zig:44-65
    const f32 inf = std::numeric_limits<f32>::infinity();
    f32 min_x = inf, min_y = inf, min_z = inf, max_x = -inf, max_y = -inf, max_z = -inf;
    for (size_t i = 0; i < n; i++) {
        min_x = zmin(min_x, v[i].x);
        min_y = zmin(min_y, v[i].y);
        min_z = zmin(min_z, v[i].z);
        max_x = zmax(max_x, v[i].x);
        max_y = zmax(max_y, v[i].y);
        max_z = zmax(max_z, v[i].z);
    }
    return aabbMinMax(v3(min_x, min_y, min_z), v3(max_x, max_y, max_z));
}
inline AABB aabbUnion(const AABB &a, const AABB &b) { // aabb.zig:68-71 initAabb
    return aabbMinMax(minimumVec(a.min, b.min), maximumVec(a.max, b.max));
}
inline f32 aabbVolume(const AABB &b) { // aabb.zig:84-87
    const Vec3 d = minus(b.min, b.max);
    return std::fabs(d.x) * std::fabs(d.y) * std::fabs(d.z);
}
inline f32 aabbSurfaceArea(const AABB &b) { // aabb.zig:99-105  (2*sum d^2, Q7)
    const Vec3 d = minus(b.min, b.max);
    const f32 dx = std::fabs(d.x), dy = std::fabs(d.y), dz = std::fabs(d.z);
    return 2 * (dx * dx + dy * dy + dz * dz);
}

struct Stats : zro_stats {
    Stats() { std::memset(static_cast<zro_stats *>(this), 0, sizeof(zro_stats)); }
};

// aabb.zig:109-127, literal: t_min/t_max are immutable parameters, tmin/tmax per-axis locals (Q4)
inline bool hitAabbRef(const AABB &box, const Ray &ray, f32 t_min, f32 t_max) {
    for (int i = 0; i < 3; i++) {
        const f32 inv_d = 1.0f / ray.direction.elem(i);
        f32 t0 = (box.min.elem(i) - ray.origin.elem(i)) * inv_d;
        f32 t1 = (box.max.elem(i) - ray.origin.elem(i)) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        const f32 tmin = zmax(t0, t_min);
        const f32 tmax = zmin(t1, t_max);
        if (tmax <= tmin) return false;
    }
    return true;
}
// the same loop with the interval carried from axis to axis (what the cited RTIOW method does)
inline bool hitAabbTight(const AABB &box, const Ray &ray, f32 t_min, f32 t_max) {
    for (int i = 0; i < 3; i++) {
        const f32 inv_d = 1.0f / ray.direction.elem(i);
        f32 t0 = (box.min.elem(i) - ray.origin.elem(i)) * inv_d;
        f32 t1 = (box.max.elem(i) - ray.origin.elem(i)) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        t_min = zmax(t0, t_min);
        t_max = zmin(t1, t_max);
        if (t_max <= t_min) return false;
    }
    return true;
}

// ---------------------------------------------------------------- surfaces
struct Surface;
struct HitRecord { // hit_record.zig:14-26
    Vec3 location, normal;
    f32 t;
    bool front_face;
    const Surface *surface;
    Vec2 texture_coords;
};
inline HitRecord hitRecordInit(const Ray &ray, Vec3 location, Vec3 outward_normal, f32 t,
                               const Surface *surface, Vec2 uv) { // hit_record.zig:28-41
    if (dot(ray.direction, outward_normal) > 0.0f)
        return HitRecord{location, negate(outward_normal), t, false, surface, uv};
    return HitRecord{location, outward_normal, t, true, surface, uv};
}

struct Sphere { // sphere.zig:15-20
    Vec3 center;
    f32 radius;
    uint32_t material;
    AABB aabb;
};
inline Sphere sphereInit(Vec3 c, f32 r, uint32_t m) { // sphere.zig:24-29
    return Sphere{c, r, m, aabbMinMax(minus(c, v3(r, r, r)), plus(c, v3(r, r, r)))};
}
struct Triangle { // triangle.zig:15-30
    Vec3 a, b, c, e1, e2, face_normal, face_unit_normal;
    uint32_t material;
    AABB aabb;
};
inline Triangle triangleInit(Vec3 a, Vec3 b, Vec3 c, uint32_t m) { // triangle.zig:32-44
    Triangle t;
    t.aabb = aabbUnion(aabbMinMax(a, b), aabbMinMax(a, c));
    t.a = a; t.b = b; t.c = c;
    t.e1 = minus(b, a);
    t.e2 = minus(c, a);
    t.face_normal = cross(t.e1, t.e2);
    t.face_unit_normal = unitVector(t.face_normal);
    t.material = m;
    return t;
}
struct BVHNode { // bvh.zig:32-35
    AABB aabb;
    Surface *left_child, *right_child;
};
enum { K_SPHERE = 0, K_TRIANGLE = 1, K_BVH = 2 };
struct Surface { // surface.zig:12-15 tagged union
    int kind;
    uint32_t id; // position in the caller's ArrayList(Surface) (not in the reference; for parity AOVs)
    Sphere sphere;
    Triangle triangle;
    BVHNode node;
    const AABB &aabb() const { // surface.zig:51-60
        return kind == K_SPHERE ? sphere.aabb : (kind == K_TRIANGLE ? triangle.aabb : node.aabb);
    }
    uint32_t material() const { return kind == K_SPHERE ? sphere.material : triangle.material; } // :39-48
};

struct Ctx {
    int traversal;
    Stats *stats;
    int math = ZRO_MATH_SPEC;
    bool roulette = false; // ZRT_FLAG_RUSSIAN_ROULETTE (sampler extension, see rrProbability below)
};
// transcendental dispatch: ZRO_MATH_SPEC = the kernels of zro_math.h, ZRO_MATH_LIBM = glibc
inline f32 mAcos(int m, f32 x) { return m == ZRO_MATH_LIBM ? std::acos(x) : zro_math::acos(x); }
inline f32 mAtan2(int m, f32 y, f32 x) { return m == ZRO_MATH_LIBM ? std::atan2(y, x) : zro_math::atan2(y, x); }
inline void mSinCos(int m, f32 x, f32 *s, f32 *c) {
    if (m == ZRO_MATH_LIBM) { *s = std::sin(x); *c = std::cos(x); } else zro_math::sincos(x, s, c);
}
inline f32 mPow5(int m, f32 x) { return m == ZRO_MATH_LIBM ? std::pow(x, 5.0f) : zro_math::pow5(x); }

bool surfaceHit(const Surface *s, const Ray &ray, f32 t_min, f32 t_max, const Ctx &cx, HitRecord *out);

// sphere.zig:31-71
inline bool sphereHit(const Sphere &sp, const Surface *surface, const Ray &ray, f32 t_min, f32 t_max,
                      const Ctx &cx, HitRecord *out) {
    cx.stats->sphere_tests++;
    const Vec3 oc = minus(ray.origin, sp.center);
    const f32 half_b = dot(oc, ray.direction);
    const f32 c = lengthSquared(oc) - (sp.radius * sp.radius);
    const f32 discriminant = half_b * half_b - c;
    if (discriminant < 0) return false;
    cx.stats->sphere_sqrt++;
    const f32 root = std::sqrt(discriminant);
    const f32 pi = (f32)3.14159265358979323846;
    const f32 two_pi = (f32)(2 * 3.14159265358979323846);
    const f32 ts[2] = {-half_b - root, -half_b + root};
    for (int k = 0; k < 2; k++) { // the reference writes the two candidate blocks out by hand
        const f32 t = ts[k];
        if (t < t_max && t > t_min) {
            const Vec3 location = rayAt(ray, t);
            const Vec3 outward_normal = scale(minus(location, sp.center), 1.0f / sp.radius);
            const f32 theta = mAcos(cx.math, -outward_normal.y);
            const f32 phi = mAtan2(cx.math, -outward_normal.z, -outward_normal.x) + pi;
            const f32 u = phi / two_pi;
            const f32 v = theta / pi;
            cx.stats->sphere_accepts++;
            *out = hitRecordInit(ray, location, outward_normal, t, surface, Vec2{u, v});
            return true;
        }
    }
    return false;
}

// triangle.zig:48-70 (single-sided, un-normalised det threshold, no early out: Q9)
inline bool triangleHit(const Triangle &tr, const Surface *surface, const Ray &ray, f32 t_min, f32 t_max,
                        const Ctx &cx, HitRecord *out) {
    cx.stats->triangle_tests++;
    const f32 det = -dot(ray.direction, tr.face_normal);
    const f32 inv_det = 1.0f / det;
    const Vec3 ao = minus(ray.origin, tr.a);
    const Vec3 dao = cross(ao, ray.direction);
    const f32 u = dot(tr.e2, dao) * inv_det;
    const f32 v = -dot(tr.e1, dao) * inv_det;
    const f32 t = dot(ao, tr.face_normal) * inv_det;
    const bool is_hit = det >= 1e-6f && t > t_min && t < t_max && u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f;
    if (is_hit) {
        const Vec3 location = plus(ray.origin, scale(ray.direction, t));
        cx.stats->triangle_accepts++;
        *out = hitRecordInit(ray, location, tr.face_unit_normal, t, surface, Vec2{u, v});
        return true;
    }
    return false;
}

// bvh.zig:187-205
inline bool bvhHit(const BVHNode &node, const Ray &ray, f32 t_min, f32 t_max, const Ctx &cx, HitRecord *out) {
    cx.stats->box_tests++;
    const bool box = cx.traversal == ZRO_TRAVERSAL_REF ? hitAabbRef(node.aabb, ray, t_min, t_max)
                                                       : hitAabbTight(node.aabb, ray, t_min, t_max);
    if (!box) return false;
    cx.stats->box_passes++;
    HitRecord hit_left;
    if (!surfaceHit(node.left_child, ray, t_min, t_max, cx, &hit_left))
        return surfaceHit(node.right_child, ray, t_min, t_max, cx, out);
    HitRecord hit_right;
    if (surfaceHit(node.right_child, ray, t_min, hit_left.t, cx, &hit_right)) {
        *out = hit_right;
        return true;
    }
    *out = hit_left;
    return true;
}

// surface.zig:28-36
bool surfaceHit(const Surface *s, const Ray &ray, f32 t_min, f32 t_max, const Ctx &cx, HitRecord *out) {
    switch (s->kind) {
    case K_BVH: return bvhHit(s->node, ray, t_min, t_max, cx, out);
    case K_TRIANGLE: return triangleHit(s->triangle, s, ray, t_min, t_max, cx, out);
    default: return sphereHit(s->sphere, s, ray, t_min, t_max, cx, out);
    }
}

// ---------------------------------------------------------------- bvh.zig:62-185 build
struct BvhBuilder {
    std::vector<std::unique_ptr<Surface>> pool; // the reference allocates nodes from an arena
    uint64_t max_depth = 0;

    static bool lessThanAxis(int axis, const Surface *a, const Surface *b) { // bvh.zig:38-56
        return a->aabb().midpoint.elem(axis) < b->aabb().midpoint.elem(axis);
    }
    static AABB surfacesToAabb(Surface **s, size_t n) { // bvh.zig:62-69 + aabb.zig:73-81
        std::vector<Vec3> pts;
        pts.reserve(2 * n);
        for (size_t i = 0; i < n; i++) {
            pts.push_back(s[i]->aabb().min);
            pts.push_back(s[i]->aabb().max);
        }
        return aabbVertexes(pts.data(), pts.size());
    }
    static void sortAxis(int axis, Surface **s, size_t n) { // bvh.zig:71-72 (std.sort.sort is stable)
        std::stable_sort(s, s + n, [axis](const Surface *a, const Surface *b) { return lessThanAxis(axis, a, b); });
    }
    // bvh.zig:85-120; returns the split position, leaves `s` sorted on the chosen axis
    static size_t optimalAxisDivide(Surface **s, size_t n) {
        int best_axis = 0;
        f32 best_ratio = std::numeric_limits<f32>::infinity();
        size_t best_split = n / 2;
        const f32 total_area = aabbSurfaceArea(surfacesToAabb(s, n));
        size_t splits[3] = {n / 2, 0, 0};
        int n_splits = 1;
        if (n >= 4) {
            splits[0] = n / 4; splits[1] = n / 2; splits[2] = n / 4 + n / 2;
            n_splits = 3;
        }
        for (int axis = 0; axis < 3; axis++) {
            for (int k = 0; k < n_splits; k++) {
                const size_t split = splits[k];
                sortAxis(axis, s, n); // make_axis_divide sorts in place every time
                const AABB right_aabb = surfacesToAabb(s + split, n - split);
                const AABB left_aabb = surfacesToAabb(s, split);
                const f32 area = aabbSurfaceArea(right_aabb) + aabbSurfaceArea(left_aabb);
                const f32 ratio = area / total_area;
                if (ratio < best_ratio) {
                    best_ratio = ratio;
                    best_axis = axis;
                    best_split = split;
                }
            }
        }
        sortAxis(best_axis, s, n); // "redo the best split"
        return best_split;
    }
    Surface *create(Surface *left, Surface *right) { // bvh.zig:162-169
        auto node = std::make_unique<Surface>();
        node->kind = K_BVH;
        node->id = ZRT_NO_HIT;
        node->node.aabb = aabbUnion(left->aabb(), right->aabb());
        node->node.left_child = left;
        node->node.right_child = right;
        pool.push_back(std::move(node));
        return pool.back().get();
    }
    Surface *divide(Surface **s, size_t n, uint64_t depth) { // bvh.zig:129-160
        if (depth > max_depth) max_depth = depth;
        if (n == 1) return create(s[0], s[0]);
        if (n == 2) return create(s[1], s[0]);
        const size_t split = optimalAxisDivide(s, n);
        Surface *left = divide(s, split, depth + 1);
        Surface *right = divide(s + split, n - split, depth + 1);
        return create(left, right);
    }
};

// ---------------------------------------------------------------- RNG
// Zig std.rand before 0.8: DefaultPrng = Xoroshiro128 (xoroshiro128+ 55/14/36) seeded via SplitMix64.
struct Xoroshiro128 {
    uint64_t s[2];
    explicit Xoroshiro128(uint64_t seed) {
        uint64_t sm = seed;
        auto splitmix = [&sm]() {
            sm += 0x9e3779b97f4a7c15ull;
            uint64_t z = sm;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            return z ^ (z >> 31);
        };
        s[0] = splitmix();
        s[1] = splitmix();
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t s0 = s[0];
        uint64_t s1 = s[1];
        const uint64_t r = s0 + s1;
        s1 ^= s0;
        s[0] = rotl(s0, 55) ^ s1 ^ (s1 << 14);
        s[1] = rotl(s1, 36);
        return r;
    }
    // Random.float(f32): low 32 bits of one u64, 23 mantissa bits, [1,2) - 1
    f32 float32() { return bitsToFloat((uint32_t)next()); }
    bool boolean() { return (next() & 1u) != 0; } // Random.boolean = int(u1) = low bit of one byte
    static f32 bitsToFloat(uint32_t s) {
        const uint32_t repr = (0x7fu << 23) | (s >> 9);
        f32 f;
        std::memcpy(&f, &repr, 4);
        return f - 1.0f;
    }
};

// Counter-based generator used by the GPU path: pcg4d (Jarzynski & Olano, "Hash Functions for GPU
// Rendering", JCGT 2020) of (pixel, sample, bounce, seed32).  Spec in DESIGN.md "RNG".
inline void rngCtr(uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t seed32, uint32_t out[4]) {
    uint32_t x = pixel, y = sample, z = bounce, w = seed32;
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    z = z * 1664525u + 1013904223u;
    w = w * 1664525u + 1013904223u;
    x += y * w; y += z * x; z += x * y; w += y * z;
    x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
    x += y * w; y += z * x; z += x * y; w += y * z;
    out[0] = x; out[1] = y; out[2] = z; out[3] = w;
}
inline uint32_t foldSeed(uint64_t seed) { return (uint32_t)seed ^ (uint32_t)(seed >> 32); }

struct Rng {
    int mode;
    Xoroshiro128 *seq; // ZRO_RNG_REF: the one shared stream
    uint32_t pixel = 0, sample = 0, seed32 = 0;
    const uint32_t *fixed = nullptr; // unit tests: use these four words for every draw

    void words(uint32_t bounce, uint32_t w[4]) const {
        if (fixed) { std::memcpy(w, fixed, 16); return; }
        rngCtr(pixel, sample, bounce, seed32, w);
    }
    void jitter(f32 *xi_u, f32 *xi_v) { // raytrace.zig:173-174 (u first, then v)
        if (mode == ZRO_RNG_REF) { *xi_u = seq->float32(); *xi_v = seq->float32(); return; }
        uint32_t w[4]; words(0, w);
        *xi_u = Xoroshiro128::bitsToFloat(w[0]);
        *xi_v = Xoroshiro128::bitsToFloat(w[1]);
    }
    void lambertian(uint32_t bounce, f32 *r1, f32 *r2, bool *coin) { // sample.zig:47-61
        if (mode == ZRO_RNG_REF) { *r1 = seq->float32(); *r2 = seq->float32(); *coin = seq->boolean(); return; }
        uint32_t w[4]; words(bounce, w);
        *r1 = Xoroshiro128::bitsToFloat(w[0]);
        *r2 = Xoroshiro128::bitsToFloat(w[1]);
        *coin = (w[2] >> 31) != 0;
    }
    f32 roulette(uint32_t bounce) { // sampler extension: 4th word of the scatter's draw
        if (mode == ZRO_RNG_REF) return seq->float32();
        uint32_t w[4]; words(bounce, w);
        return Xoroshiro128::bitsToFloat(w[3]);
    }
    f32 dielectric(uint32_t bounce) { // material.zig:117
        if (mode == ZRO_RNG_REF) return seq->float32();
        uint32_t w[4]; words(bounce, w);
        return Xoroshiro128::bitsToFloat(w[0]);
    }
};

// sample.zig:47-61
inline Vec3 randomUnitVectorFrom(f32 r1, f32 r2, bool coin, int math) {
    const f32 r = std::sqrt(1.0f - r1 * r1);
    const f32 phi = (f32)(2.0 * 3.14159265358979323846) * r2;
    f32 sn, cs;
    mSinCos(math, phi, &sn, &cs);
    const Vec3 v = v3(cs * r, sn * r, r1);
    if (coin) return v;
    return v3(v.x, v.y, v.z * -1.0f);
}

// ---------------------------------------------------------------- texture.zig
struct Scene {
    const zrt_scene_desc *desc;
    std::vector<Surface> surfaces;         // caller's list, in order
    std::vector<Surface *> render_list;    // what rayColor loops over (raytrace.zig:71-81)
    BvhBuilder bvh;
};

inline uint64_t floatToU64(f32 f) { // @floatToInt(u64, f); negative/NaN is UB in the reference -> 0 here
    if (!(f >= 0.0f)) return 0;
    if (f >= 18446744073709551615.0f) return UINT64_MAX;
    return (uint64_t)f;
}
inline Color textureAlbedo(const zrt_texture &tex, Vec2 uv, Stats *stats) { // texture.zig:20-28
    if (tex.kind == ZRT_TEXTURE_COLOR) return Color{tex.r, tex.g, tex.b}; // :36-40
    // ImageTexture.albedo texture.zig:52-74 (Q17: the second else-if tests uu_first, not vv_first)
    if (stats) stats->texture_lookups++;
    const f32 uu_first = (1.0f - uv.u + tex.u_offset);
    f32 uu = uu_first;
    if (uu_first > 1.0f) uu = uu_first - 1.0f;
    else if (uu_first < 0) uu = uu_first + 1.0f;
    const f32 vv_first = uv.v + tex.v_offset;
    f32 vv = vv_first;
    if (vv_first > 1.0f) vv = vv_first - 1.0f;
    else if (uu_first < 0) vv = vv_first + 1.0f;
    const uint64_t w = tex.width, h = tex.height;
    uint64_t img_x = floatToU64(uu * (f32)tex.width);
    uint64_t img_y = floatToU64(vv * (f32)tex.height);
    img_x = std::max<uint64_t>(0, std::min<uint64_t>(img_x, w - 1));
    img_y = std::max<uint64_t>(0, std::min<uint64_t>(img_y, h - 1));
    const uint8_t *p = tex.pixels + (img_y * w + img_x) * tex.channels;
    // png_image.zig:87: @intToFloat(f32, px)/0xff
    return Color{(f32)p[0] / 255.0f, (f32)p[1] / 255.0f, (f32)p[2] / 255.0f};
}

// ---------------------------------------------------------------- material.zig
struct Scattering {
    Ray scattered_ray;
    Color attenuation;
};
inline bool materialScatter(const zrt_scene_desc *desc, uint32_t mat_index, const Ray &ray, const HitRecord &hit,
                            Rng &rng, uint32_t bounce, Stats *stats, int math, Scattering *out) { // material.zig:43-51
    const zrt_material &m = desc->materials[mat_index];
    switch (m.kind) {
    case ZRT_MATERIAL_LAMBERTIAN: { // material.zig:71-76
        f32 r1, r2; bool coin;
        rng.lambertian(bounce, &r1, &r2, &coin);
        const Vec3 scatter_direction = plus(hit.normal, randomUnitVectorFrom(r1, r2, coin, math));
        out->scattered_ray = rayInit(hit.location, scatter_direction);
        out->attenuation = textureAlbedo(desc->textures[m.texture], hit.texture_coords, stats);
        stats->lambertian++;
        return true;
    }
    case ZRT_MATERIAL_METAL: { // material.zig:87-96
        const Vec3 reflected = reflect(unitVector(ray.direction), hit.normal);
        const Ray scattered = rayInit(hit.location, reflected);
        const bool produce_ray = dot(scattered.direction, hit.normal) > 0;
        if (produce_ray) {
            out->scattered_ray = scattered;
            out->attenuation = textureAlbedo(desc->textures[m.texture], hit.texture_coords, stats);
            stats->metal++;
            return true;
        }
        stats->metal_absorbed++;
        return false;
    }
    default: { // Dielectric material.zig:109-128
        const f32 ior = m.index_of_refraction;
        const f32 refraction_ratio = hit.front_face ? (1.0f / ior) : ior;
        const Vec3 unit_direction = unitVector(ray.direction);
        const f32 cos_theta = zmin(dot(negate(unit_direction), hit.normal), 1.0f);
        const f32 sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        const bool cannot_refract = refraction_ratio * sin_theta > 1.0f;
        bool do_reflect = cannot_refract;
        if (!do_reflect) { // `or` short-circuits: xi is drawn only when refraction is possible (Q15)
            const f32 r0 = (1.0f - refraction_ratio) / (1.0f + refraction_ratio); // NOT squared (Q15)
            const f32 reflectance = r0 + (1.0f - r0) * mPow5(math, 1 - cos_theta);
            do_reflect = (double)reflectance > (double)rng.dielectric(bounce);
        }
        out->attenuation = Color{1.0f, 1.0f, 1.0f};
        if (do_reflect) {
            out->scattered_ray = rayInit(hit.location, reflect(unit_direction, hit.normal));
            stats->dielectric_reflect++;
        } else {
            out->scattered_ray = rayInit(hit.location, refract(unit_direction, hit.normal, refraction_ratio));
            stats->dielectric_refract++;
        }
        return true;
    }
    }
}

// ---------------------------------------------------------------- raytrace.zig
inline Color backgroundColor(const Ray &ray) { // raytrace.zig:53-58
    const Vec3 unit_direction = unitVector(ray.direction);
    const f32 t = 0.5f * (unit_direction.y + 1.0f);
    return cadd(cscale(Color{1.0f, 1.0f, 1.0f}, 1.0f - t), cscale(Color{0.5f, 0.7f, 1.0f}, t));
}

inline bool closestHit(const Scene &sc, const Ray &ray, const Ctx &cx, HitRecord *closest) { // raytrace.zig:71-81
    const f32 t_min = 0.001f;
    f32 t_max = std::numeric_limits<f32>::infinity();
    bool any = false;
    for (const Surface *s : sc.render_list) {
        HitRecord h;
        if (surfaceHit(s, ray, t_min, t_max, cx, &h)) {
            *closest = h;
            t_max = h.t;
            any = true;
        }
    }
    return any;
}

// ---- sampler extensions (the reference's TODO list src/README.md:5-13; spec in include/zrt.h and DESIGN.md) ----
constexpr uint32_t kRouletteStart = 3;
inline f32 rrProbability(const Color &thr) { // continuation probability: max channel, clamped to [0.05, 1]
    f32 p = thr.r;
    if (thr.g > p) p = thr.g;
    if (thr.b > p) p = thr.b;
    if (p > 1.0f) p = 1.0f;
    if (p < 0.05f) p = 0.05f;
    return p;
}
inline f32 halton2(uint32_t i) { // radical inverse base 2: bit reversal, exact in f32 after dropping 8 bits
    uint32_t r = 0;
    for (int b = 0; b < 32; b++) r |= ((i >> b) & 1u) << (31 - b);
    return (f32)(r >> 8) * 5.9604644775390625e-8f;
}
inline f32 halton3(uint32_t i) { // digits of i in base 3, reversed, over 3^k
    uint32_t r = 0, d = 1;
    while (i) {
        r = r * 3u + i % 3u;
        d *= 3u;
        i /= 3u;
    }
    return (f32)r / (f32)d;
}

// thr: product of the attenuations from the camera to this ray, front to back (only the roulette reads it)
Color rayColor(const Scene &sc, const Ray &ray, uint32_t depth, uint32_t max_depth, zrt_counters *progress,
               Rng &rng, const Ctx &cx, Color thr = Color{1.0f, 1.0f, 1.0f}) { // raytrace.zig:62-100
    if (depth <= 0) {
        progress->recursion_depth_hits += 1;
        return Color{0, 0, 0};
    }
    progress->rays_processed += 1;
    HitRecord hit;
    if (!closestHit(sc, ray, cx, &hit)) {
        progress->background_hits += 1;
        cx.stats->background++;
        return backgroundColor(ray);
    }
    Scattering sct;
    const uint32_t bounce = max_depth - depth + 1; // 1 for the primary ray's hit
    if (!materialScatter(sc.desc, hit.surface->material(), ray, hit, rng, bounce, cx.stats, cx.math, &sct))
        return Color{0, 0, 0};
    progress->reflections += 1;
    if (cx.roulette) {
        thr = cmul(thr, sct.attenuation);
        if (bounce >= kRouletteStart) {
            const f32 p = rrProbability(thr);
            if (rng.roulette(bounce) >= p) return Color{0, 0, 0}; // ends here: no further ray, nothing counted
            thr = Color{thr.r / p, thr.g / p, thr.b / p};
            const Color c = rayColor(sc, sct.scattered_ray, depth - 1, max_depth, progress, rng, cx, thr);
            return cmul(sct.attenuation, Color{c.r / p, c.g / p, c.b / p});
        }
        return cmul(sct.attenuation, rayColor(sc, sct.scattered_ray, depth - 1, max_depth, progress, rng, cx, thr));
    }
    return cmul(sct.attenuation, rayColor(sc, sct.scattered_ray, depth - 1, max_depth, progress, rng, cx));
}

inline Ray cameraGetRay(const zrt_camera &cam, f32 u, f32 v) { // camera.zig:46-52
    const Vec3 dir = minus(plus(plus(v3(cam.lower_left_corner), scale(v3(cam.horizontal), u)),
                                scale(v3(cam.vertical), v)),
                           v3(cam.origin));
    return rayInit(v3(cam.origin), dir);
}

bool validate(const zrt_scene_desc *d) {
    if (!d) return false;
    for (uint32_t i = 0; i < d->n_surfaces; i++) {
        const zrt_surface &s = d->surfaces[i];
        if (s.kind == ZRT_SURFACE_SPHERE ? s.index >= d->n_spheres : s.index >= d->n_triangles) return false;
    }
    return true;
}

// raytrace.zig:111-133 preprocessSufraces + boundedVolumeHierarchy
void buildScene(Scene &sc, const zrt_scene_desc *desc, bool use_bvh) {
    sc.desc = desc;
    sc.surfaces.resize(desc->n_surfaces);
    for (uint32_t i = 0; i < desc->n_surfaces; i++) {
        Surface &s = sc.surfaces[i];
        s.id = i;
        if (desc->surfaces[i].kind == ZRT_SURFACE_SPHERE) {
            const zrt_sphere &p = desc->spheres[desc->surfaces[i].index];
            s.kind = K_SPHERE;
            s.sphere = sphereInit(v3(p.center), p.radius, p.material);
        } else {
            const zrt_triangle &p = desc->triangles[desc->surfaces[i].index];
            s.kind = K_TRIANGLE;
            s.triangle = triangleInit(v3(p.a), v3(p.b), v3(p.c), p.material);
        }
    }
    sc.render_list.clear();
    if (use_bvh && desc->n_surfaces > 10) { // raytrace.zig:127
        std::vector<Surface *> ptrs(desc->n_surfaces);
        for (uint32_t i = 0; i < desc->n_surfaces; i++) ptrs[i] = &sc.surfaces[i];
        Surface *root = sc.bvh.divide(ptrs.data(), ptrs.size(), 1); // bvh.zig:171-185
        sc.render_list.push_back(root);
    } else {
        for (auto &s : sc.surfaces) sc.render_list.push_back(&s);
    }
}

void sampleRange(const zrt_params *p, uint32_t *b, uint32_t *e) {
    *b = p->sample_begin;
    *e = p->sample_end;
    if (*b == 0 && *e == 0) *e = p->samples_per_pixel;
}

uint32_t xLimit(const zrt_params *p) { // raytrace.zig:168 `while (x < image.height)` (Q1)
    if (p->x_limit == ZRT_XLIMIT_WIDTH) return p->width;
    return std::min(p->height, p->width); // the reference would write out of row when height > width
}

void countBvh(const Surface *s, uint64_t depth, Stats *st) {
    if (s->kind != K_BVH) return;
    st->bvh_nodes++;
    if (depth > st->bvh_max_depth) st->bvh_max_depth = depth;
    countBvh(s->node.left_child, depth + 1, st);
    if (s->node.right_child != s->node.left_child) countBvh(s->node.right_child, depth + 1, st);
}

void renderRows(const Scene &sc, const zrt_camera *camera, const zrt_params *p, int rng_mode, int traversal, int math,
                uint32_t y0, uint32_t y1, Xoroshiro128 *seq, float *out_rgb, zrt_counters *progress, Stats *stats) {
    const f32 f_width = (f32)p->width, f_height = (f32)p->height;
    const f32 color_scale = (p->flags & ZRT_FLAG_RAW_SUM) ? 1.0f : 1.0f / (f32)p->samples_per_pixel; // :157
    uint32_t s_begin, s_end;
    sampleRange(p, &s_begin, &s_end);
    const uint32_t x_end = xLimit(p);
    Ctx cx{traversal, stats, math};
    cx.roulette = (p->flags & ZRT_FLAG_RUSSIAN_ROULETTE) != 0;
    const bool halton = (p->flags & ZRT_FLAG_SAMPLER_HALTON) != 0;
    Rng rng{rng_mode, seq};
    rng.seed32 = foldSeed(p->seed);
    for (uint32_t y = y0; y < y1; y++) { // raytrace.zig:162-187
        const f32 f_y = (f32)y;
        for (uint32_t x = 0; x < x_end; x++) {
            Color color_acc{0, 0, 0};
            rng.pixel = y * p->width + x;
            for (uint32_t sample = s_begin; sample < s_end; sample++) {
                rng.sample = sample;
                f32 xi_u, xi_v;
                if (halton) { // Halton (2,3) point of the global sample index, rotated by the pixel's offsets
                    uint32_t w[4];
                    rngCtr(rng.pixel, 0xFFFFFFFFu, 0, rng.seed32, w);
                    xi_u = Xoroshiro128::bitsToFloat(w[0]) + halton2(sample + 1);
                    xi_v = Xoroshiro128::bitsToFloat(w[1]) + halton3(sample + 1);
                    if (xi_u >= 1.0f) xi_u -= 1.0f;
                    if (xi_v >= 1.0f) xi_v -= 1.0f;
                } else {
                    rng.jitter(&xi_u, &xi_v);
                }
                const f32 u = ((f32)x + xi_u - 0.5f) / f_width;
                const f32 v = (f_y + xi_v - 0.5f) / f_height;
                const Ray ray = cameraGetRay(*camera, u, v);
                const Color color = rayColor(sc, ray, p->max_depth, p->max_depth, progress, rng, cx);
                color_acc = cadd(color_acc, color);
                progress->samples_processed += 1;
            }
            if (s_begin == 0) progress->pixels_processed += 1; // a partial sample range (multi-GPU split) counts the
                                                               // pixel once, on the call that owns sample 0
            const Color px = cscale(color_acc, color_scale);
            float *o = out_rgb + ((size_t)y * p->width + x) * 3;
            o[0] = px.r; o[1] = px.g; o[2] = px.b;
        }
    }
}

void addCounters(zrt_counters *a, const zrt_counters &b) {
    a->recursion_depth_hits += b.recursion_depth_hits;
    a->reflections += b.reflections;
    a->background_hits += b.background_hits;
    a->pixels_processed += b.pixels_processed;
    a->samples_processed += b.samples_processed;
    a->rays_processed += b.rays_processed;
}
void addStats(zro_stats *a, const zro_stats &b) {
    uint64_t *pa = reinterpret_cast<uint64_t *>(a);
    const uint64_t *pb = reinterpret_cast<const uint64_t *>(&b);
    for (size_t i = 0; i < sizeof(zro_stats) / 8; i++) pa[i] += pb[i];
}

} // namespace

extern "C" {

int zro_render(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *p, int rng_mode,
               int traversal_mode, int math_mode, int n_threads, float *out_rgb, zrt_counters *counters,
               zro_stats *stats_out) {
    if (!validate(desc) || !camera || !p || !out_rgb || p->width == 0 || p->height == 0) return ZRT_ERR_INVALID;
    if (n_threads > 1 && rng_mode != ZRO_RNG_CTR) return ZRT_ERR_INVALID;
    if ((p->flags & ZRT_FLAG_SAMPLER_HALTON) && rng_mode != ZRO_RNG_CTR) return ZRT_ERR_INVALID; // keyed on (pixel, sample)
    if (n_threads < 1) n_threads = 1;
    Scene sc;
    buildScene(sc, desc, p->bounded_volume_hierarchy != 0);
    std::memset(out_rgb, 0, sizeof(float) * 3 * (size_t)p->width * p->height); // image.zig:80-90
    zrt_counters total{};
    Stats stats_total;
    if (!sc.render_list.empty() && sc.render_list[0]->kind == K_BVH) countBvh(sc.render_list[0], 1, &stats_total);
    if (n_threads == 1) {
        Xoroshiro128 seq(p->seed);
        renderRows(sc, camera, p, rng_mode, traversal_mode, math_mode, 0, p->height, &seq, out_rgb, &total, &stats_total);
    } else {
        std::vector<std::thread> th;
        std::vector<zrt_counters> pc(n_threads);
        std::vector<Stats> ps(n_threads);
        for (int t = 0; t < n_threads; t++) {
            pc[t] = zrt_counters{};
            th.emplace_back([&, t]() {
                // interleaved scanlines balance sky/ground/glass rows across threads
                for (uint32_t y = t; y < p->height; y += n_threads)
                    renderRows(sc, camera, p, rng_mode, traversal_mode, math_mode, y, y + 1, nullptr, out_rgb, &pc[t], &ps[t]);
            });
        }
        for (auto &t : th) t.join();
        for (int t = 0; t < n_threads; t++) { addCounters(&total, pc[t]); addStats(&stats_total, ps[t]); }
    }
    if (counters) *counters = total;
    if (stats_out) *stats_out = stats_total;
    return ZRT_OK;
}

// rows [y0, y1) of the first-hit AOV; each row only reads the scene, so rows can run on separate threads
static void primaryRows(const Scene &sc, const zrt_camera *camera, const zrt_params *p, int jitter, int traversal_mode,
                        uint32_t y0, uint32_t y1, uint32_t y_step, uint32_t *surface_id, float *t_out) {
    Stats stats;
    Ctx cx{traversal_mode, &stats};
    const f32 f_width = (f32)p->width, f_height = (f32)p->height;
    Rng rng{ZRO_RNG_CTR, nullptr};
    rng.seed32 = foldSeed(p->seed);
    rng.sample = p->sample_begin;
    for (uint32_t y = y0; y < y1; y += y_step)
        for (uint32_t x = 0; x < p->width; x++) {
            f32 xi_u = 0.0f, xi_v = 0.0f;
            rng.pixel = y * p->width + x;
            if (jitter) rng.jitter(&xi_u, &xi_v);
            const f32 u = ((f32)x + xi_u - 0.5f) / f_width;
            const f32 v = ((f32)y + xi_v - 0.5f) / f_height;
            const Ray ray = cameraGetRay(*camera, u, v);
            HitRecord hit;
            const size_t o = (size_t)y * p->width + x;
            if (closestHit(sc, ray, cx, &hit)) {
                surface_id[o] = hit.surface->id;
                t_out[o] = hit.t;
            } else {
                surface_id[o] = ZRT_NO_HIT;
                t_out[o] = std::numeric_limits<f32>::infinity();
            }
        }
}

int zro_primary_hits_mt(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *p, int jitter,
                        int traversal_mode, int n_threads, uint32_t *surface_id, float *t_out) {
    if (!validate(desc) || !camera || !p || !surface_id || !t_out) return ZRT_ERR_INVALID;
    Scene sc;
    buildScene(sc, desc, p->bounded_volume_hierarchy != 0);
    if (n_threads <= 1) {
        primaryRows(sc, camera, p, jitter, traversal_mode, 0, p->height, 1, surface_id, t_out);
        return ZRT_OK;
    }
    std::vector<std::thread> th; // interleaved scanlines, as zro_render
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&, t]() { primaryRows(sc, camera, p, jitter, traversal_mode, (uint32_t)t, p->height, (uint32_t)n_threads, surface_id, t_out); });
    for (auto &t : th) t.join();
    return ZRT_OK;
}

int zro_primary_hits(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *p, int jitter,
                     int traversal_mode, uint32_t *surface_id, float *t_out) {
    return zro_primary_hits_mt(desc, camera, p, jitter, traversal_mode, 1, surface_id, t_out);
}

static void dfsOrder(const Surface *s, bool under_flat, std::vector<uint32_t> &order, std::vector<uint8_t> &vis,
                     std::vector<uint8_t> &seen) {
    if (s->kind != K_BVH) {
        if (!seen[s->id]) { seen[s->id] = 1; order.push_back(s->id); }
        if (!under_flat) vis[s->id] = 1;
        return;
    }
    const AABB &b = s->node.aabb;
    const bool flat = under_flat || b.min.x == b.max.x || b.min.y == b.max.y || b.min.z == b.max.z;
    dfsOrder(s->node.left_child, flat, order, vis, seen);
    dfsOrder(s->node.right_child, flat, order, vis, seen);
}

int zro_bvh_order(const zrt_scene_desc *desc, uint32_t *order, uint8_t *visible, zro_stats *stats_out) {
    if (!validate(desc) || !order || !visible || desc->n_surfaces == 0) return ZRT_ERR_INVALID;
    Scene sc;
    buildScene(sc, desc, true);
    std::vector<uint32_t> ord;
    std::vector<uint8_t> vis(desc->n_surfaces, 0), seen(desc->n_surfaces, 0);
    Stats st;
    if (sc.render_list[0]->kind == K_BVH) {
        dfsOrder(sc.render_list[0], false, ord, vis, seen);
        countBvh(sc.render_list[0], 1, &st);
    } else {
        for (uint32_t i = 0; i < desc->n_surfaces; i++) { ord.push_back(i); vis[i] = 1; }
    }
    std::memcpy(order, ord.data(), sizeof(uint32_t) * ord.size());
    std::memcpy(visible, vis.data(), vis.size());
    if (stats_out) *stats_out = st;
    return ZRT_OK;
}

void zro_camera_init(const float look_from[3], const float look_at[3], const float vup[3], float vfov,
                     float aspect_ratio, zrt_camera *out) { // camera.zig:7-35
    const f32 theta = (f32)3.14159265358979323846 * vfov / 180.0f; // deg2rad camera.zig:7-9
    const f32 h = std::tan(theta / 2.0f);
    const f32 viewport_height = 2.0f * h;
    const f32 viewport_width = aspect_ratio * viewport_height;
    const Vec3 w = unitVector(minus(v3(look_from), v3(look_at)));
    const Vec3 u = unitVector(cross(v3(vup), w));
    const Vec3 v = cross(w, u);
    const Vec3 horizontal = scale(u, viewport_width);
    const Vec3 vertical = scale(v, viewport_height);
    const Vec3 llc = minus(minus(minus(v3(look_from), scale(horizontal, 1 / 2.0f)), scale(vertical, 1 / 2.0f)), w);
    out->origin = zrt_vec3{look_from[0], look_from[1], look_from[2]};
    out->lower_left_corner = zrt_vec3{llc.x, llc.y, llc.z};
    out->horizontal = zrt_vec3{horizontal.x, horizontal.y, horizontal.z};
    out->vertical = zrt_vec3{vertical.x, vertical.y, vertical.z};
}

void zro_ray_at(const float origin[3], const float direction[3], float t, float out[3]) {
    const Vec3 p = rayAt(rayInit(v3(origin), v3(direction)), t);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void zro_vec3_unit(const float v[3], float out[3]) {
    const Vec3 u = unitVector(v3(v));
    out[0] = u.x; out[1] = u.y; out[2] = u.z;
}
float zro_vec3_dot(const float a[3], const float b[3]) { return dot(v3(a), v3(b)); }
void zro_vec3_center(const float *xyz, uint32_t n, float out[3]) {
    std::vector<Vec3> v(n);
    for (uint32_t i = 0; i < n; i++) v[i] = v3(xyz + 3 * i);
    const Vec3 c = center(v.data(), n);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

static int exportHit(bool ok, const HitRecord &h, float *t, float location[3], float normal[3], int *front_face,
                     float uv[2]) {
    if (!ok) return 0;
    if (t) *t = h.t;
    if (location) { location[0] = h.location.x; location[1] = h.location.y; location[2] = h.location.z; }
    if (normal) { normal[0] = h.normal.x; normal[1] = h.normal.y; normal[2] = h.normal.z; }
    if (front_face) *front_face = h.front_face ? 1 : 0;
    if (uv) { uv[0] = h.texture_coords.u; uv[1] = h.texture_coords.v; }
    return 1;
}
int zro_triangle_hit(const float a[3], const float b[3], const float c[3], const float origin[3],
                     const float direction[3], float t_min, float t_max, float *t, float location[3],
                     float normal[3], int *front_face, float uv[2]) {
    Surface s;
    s.kind = K_TRIANGLE; s.id = 0;
    s.triangle = triangleInit(v3(a), v3(b), v3(c), 0);
    Stats st; Ctx cx{ZRO_TRAVERSAL_REF, &st};
    HitRecord h{};
    const bool ok = triangleHit(s.triangle, &s, rayInit(v3(origin), v3(direction)), t_min, t_max, cx, &h);
    return exportHit(ok, h, t, location, normal, front_face, uv);
}
int zro_sphere_hit(const float center_[3], float radius, const float origin[3], const float direction[3],
                   float t_min, float t_max, float *t, float location[3], float normal[3], int *front_face,
                   float uv[2]) {
    Surface s;
    s.kind = K_SPHERE; s.id = 0;
    s.sphere = sphereInit(v3(center_), radius, 0);
    Stats st; Ctx cx{ZRO_TRAVERSAL_REF, &st};
    HitRecord h{};
    const bool ok = sphereHit(s.sphere, &s, rayInit(v3(origin), v3(direction)), t_min, t_max, cx, &h);
    return exportHit(ok, h, t, location, normal, front_face, uv);
}
static void put(float o[3], Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
void zro_aabb_min_max(const float c1[3], const float c2[3], float out_min[3], float out_max[3], float out_mid[3]) {
    const AABB b = aabbMinMax(v3(c1), v3(c2));
    put(out_min, b.min); put(out_max, b.max); put(out_mid, b.midpoint);
}
void zro_aabb_vertexes(const float *xyz, uint32_t n, float out_min[3], float out_max[3]) {
    std::vector<Vec3> v(n);
    for (uint32_t i = 0; i < n; i++) v[i] = v3(xyz + 3 * i);
    const AABB b = aabbVertexes(v.data(), n);
    put(out_min, b.min); put(out_max, b.max);
}
void zro_aabb_union(const float min1[3], const float max1[3], const float min2[3], const float max2[3],
                    float out_min[3], float out_max[3]) {
    const AABB b = aabbUnion(aabbMinMax(v3(min1), v3(max1)), aabbMinMax(v3(min2), v3(max2)));
    put(out_min, b.min); put(out_max, b.max);
}
float zro_aabb_surface_area(const float mn[3], const float mx[3]) { return aabbSurfaceArea(aabbMinMax(v3(mn), v3(mx))); }
float zro_aabb_volume(const float mn[3], const float mx[3]) { return aabbVolume(aabbMinMax(v3(mn), v3(mx))); }
int zro_aabb_hit(const float mn[3], const float mx[3], const float origin[3], const float direction[3],
                 float t_min, float t_max) {
    return hitAabbRef(aabbMinMax(v3(mn), v3(mx)), rayInit(v3(origin), v3(direction)), t_min, t_max) ? 1 : 0;
}
void zro_texture_albedo(const zrt_texture *tex, float u, float v, float out[3]) {
    const Color c = textureAlbedo(*tex, Vec2{u, v}, nullptr);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}

void zro_sample(int which, uint64_t seed, float out[3]) { // sample.zig:10-61
    Xoroshiro128 rng(seed);
    auto randomVector = [&]() { // :10-17
        const f32 x = rng.float32() * 2.0f - 1.0f;
        const f32 y = rng.float32() * 2.0f - 1.0f;
        const f32 z = rng.float32() * 2.0f - 1.0f;
        return v3(x, y, z);
    };
    auto inUnitSphere = [&]() { // :19-28
        while (true) {
            const Vec3 p = randomVector();
            if (lengthSquared(p) > 1.0f) continue;
            return p;
        }
    };
    Vec3 r{0, 0, 0};
    switch (which) {
    case 0: r = randomVector(); break;
    case 1: r = inUnitSphere(); break;
    case 2: // randomUnitVector_old :30-39
        while (true) {
            const Vec3 u = unitVector(inUnitSphere());
            if (std::isnan(u.x)) continue;
            r = u;
            break;
        }
        break;
    default: { // randomUnitVector :55-61
        const f32 r1 = rng.float32();
        const f32 r2 = rng.float32();
        r = randomUnitVectorFrom(r1, r2, rng.boolean(), ZRO_MATH_LIBM);
    }
    }
    put(out, r);
}

int zro_scatter(const zrt_scene_desc *desc, uint32_t material, const float origin[3], const float direction[3],
                const float location[3], const float normal[3], int front_face, const float uv[2],
                const uint32_t rnd[4], int math_mode, float out_origin[3], float out_direction[3],
                float attenuation[3]) {
    Ray ray{v3(origin), v3(direction)}; // direction taken as given (already a Ray)
    HitRecord hit{v3(location), v3(normal), 0.0f, front_face != 0, nullptr, Vec2{uv[0], uv[1]}};
    Rng rng{ZRO_RNG_CTR, nullptr};
    rng.fixed = rnd;
    Stats st;
    Scattering s;
    if (!materialScatter(desc, material, ray, hit, rng, 1, &st, math_mode, &s)) return 0;
    put(out_origin, s.scattered_ray.origin);
    put(out_direction, s.scattered_ray.direction);
    attenuation[0] = s.attenuation.r; attenuation[1] = s.attenuation.g; attenuation[2] = s.attenuation.b;
    return 1;
}

uint64_t zro_bvh_random_test(uint32_t n_spheres, uint32_t n_rays, uint64_t seed, int traversal_mode) { // bvh.zig:224-291
    Xoroshiro128 rng(seed);
    std::vector<Surface> surfaces(n_spheres);
    for (uint32_t i = 0; i < n_spheres; i++) { // createSurfaces bvh.zig:224-237
        const f32 x = (rng.float32() - 0.5f) * 100.0f;
        const f32 y = (rng.float32() - 0.5f) * 100.0f;
        const f32 z = (rng.float32() - 0.5f) * 100.0f;
        const f32 radius = rng.float32() * 10 + 0.01f;
        surfaces[i].kind = K_SPHERE;
        surfaces[i].id = i;
        surfaces[i].sphere = sphereInit(v3(x, y, z), radius, 0);
    }
    std::vector<Surface *> ptrs(n_spheres);
    for (uint32_t i = 0; i < n_spheres; i++) ptrs[i] = &surfaces[i];
    BvhBuilder b;
    Surface *root = b.divide(ptrs.data(), ptrs.size(), 1);
    Stats st; Ctx cx{traversal_mode, &st};
    uint64_t hits = 0;
    auto unit = [&]() {
        const f32 r1 = rng.float32();
        const f32 r2 = rng.float32();
        return randomUnitVectorFrom(r1, r2, rng.boolean(), ZRO_MATH_LIBM);
    };
    for (uint32_t i = 0; i < n_rays; i++) {
        const Vec3 o = scale(unit(), 100.0f);
        const Vec3 d = unit();
        HitRecord h;
        if (surfaceHit(root, rayInit(o, d), 0.0001f, std::numeric_limits<f32>::infinity(), cx, &h)) hits++;
    }
    return hits;
}

void zro_rng_ctr(uint32_t pixel, uint32_t sample, uint32_t bounce, uint64_t seed, uint32_t out[4]) {
    rngCtr(pixel, sample, bounce, foldSeed(seed), out);
}

uint8_t zro_quantize(float c) { // png_image.zig:136-140: clamp(255.999*c, 0, 255) truncated
    f32 v = 255.999f * c;
    v = zmax(0.0f, zmin(v, 255.0f));
    if (!(v >= 0.0f)) v = 0.0f;
    return (uint8_t)v;
}

} // extern "C"

extern "C" void zro_math_eval(int which, const float *x, const float *y, float *out, float *out2, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        switch (which) {
        case 0: zro_math::sincos(x[i], &out[i], &out2[i]); break;
        case 1: out[i] = zro_math::acos(x[i]); break;
        case 2: out[i] = zro_math::atan2(y[i], x[i]); break;
        default: out[i] = zro_math::pow5(x[i]); break;
        }
    }
}
