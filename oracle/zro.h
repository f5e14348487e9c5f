/*
 * zro.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a single-threaded-by-default C++ restatement of the reference's CPU path
 * (jsyrjala/zraytrace, the .zig files under src/; each function in zro.cpp cites the file:line it follows).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * it.  libzrt never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED by the reference's own known-answer tests (ray.zig:32-39,
 * triangle.zig:84-118, aabb.zig:151-254, vector.zig:169-255, texture.zig:90-103, sample.zig:70-118),
 * by the counters published in README.md:49-61 and by showcase/7-spheres.png — see
 * tests/test_oracle_kat.py.  The reference itself cannot be compiled here (no Zig toolchain).
 *
 * It shares only the POD scene description of include/zrt.h with the product (the interface, not
 * an implementation).
 */
#ifndef ZRO_H
#define ZRO_H
#include "../include/zrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* RNG modes */
enum {
    ZRO_RNG_REF = 0, /* Xoroshiro128+ seeded by SplitMix64, ONE sequential stream shared by jitter,
                        Lambertian and Dielectric in program order (scenes.zig:60-61) */
    ZRO_RNG_CTR = 1  /* counter-based pcg4d keyed (pixel, sample, bounce, seed): the generator the
                        GPU path uses (DESIGN.md "RNG"); makes paths comparable draw for draw */
};
/* BVH traversal modes (results are identical; only the visit counts differ) */
enum {
    ZRO_TRAVERSAL_REF = 0,  /* literal aabb.zig:109-127 (interval not carried between axes, Q4) */
    ZRO_TRAVERSAL_TIGHT = 1 /* interval-carrying slab test on the same tree; flat boxes still rejected */
};

/* transcendental kernels (sin cos acos atan2 pow): the reference uses Zig std.math, unpinned */
enum {
    ZRO_MATH_SPEC = 0, /* the f32 kernels specified in DESIGN.md "Spec math" (zro_math.h), which the GPU
                          path also implements: oracle and GPU agree draw for draw, bit for bit */
    ZRO_MATH_LIBM = 1  /* glibc libm: independent cross-check, statistical agreement only */
};

/* event counts used for the algorithmic-work side of the roofline */
typedef struct zro_stats {
    uint64_t sphere_tests, sphere_sqrt, sphere_accepts;
    uint64_t triangle_tests, triangle_accepts;
    uint64_t box_tests, box_passes;
    uint64_t lambertian, metal, metal_absorbed, dielectric_reflect, dielectric_refract;
    uint64_t texture_lookups, background;
    uint64_t bvh_nodes, bvh_max_depth;
} zro_stats;

/* raytrace.render() restated.  n_threads > 1 is allowed only with ZRO_RNG_CTR (scanline stripes,
 * results independent of the thread count).  out_rgb: width*height*3, row 0 = bottom. */
int zro_render(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *params,
               int rng_mode, int traversal_mode, int math_mode, int n_threads,
               float *out_rgb, zrt_counters *counters, zro_stats *stats);

/* first rayColor iteration per pixel; jitter 0 = xi 0, 1 = ctr RNG draw of sample params->sample_begin */
int zro_primary_hits(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *params,
                     int jitter, int traversal_mode, uint32_t *surface_id, float *t);

/* the same rows on n_threads host threads (scanline-interleaved; every pixel is independent, results identical) */
int zro_primary_hits_mt(const zrt_scene_desc *desc, const zrt_camera *camera, const zrt_params *params,
                        int jitter, int traversal_mode, int n_threads, uint32_t *surface_id, float *t);

/* DFS (left-first) order of the surfaces in the reference tree and which of them can never be hit
 * because they sit under a zero-thickness box (Q4).  order/visible have n_surfaces entries. */
int zro_bvh_order(const zrt_scene_desc *desc, uint32_t *order, uint8_t *visible, zro_stats *stats);

/* Camera.init camera.zig:17-35 */
void zro_camera_init(const float look_from[3], const float look_at[3], const float vup[3],
                     float vfov, float aspect_ratio, zrt_camera *out);

/* ---- unit-level entry points used by the known-answer tests ---- */
void zro_ray_at(const float origin[3], const float direction[3], float t, float out[3]);
void zro_vec3_unit(const float v[3], float out[3]);
float zro_vec3_dot(const float a[3], const float b[3]);
void zro_vec3_center(const float *xyz, uint32_t n, float out[3]);
int zro_triangle_hit(const float a[3], const float b[3], const float c[3], const float origin[3],
                     const float direction[3], float t_min, float t_max, float *t, float location[3],
                     float normal[3], int *front_face, float uv[2]);
int zro_sphere_hit(const float center[3], float radius, const float origin[3], const float direction[3],
                   float t_min, float t_max, float *t, float location[3], float normal[3],
                   int *front_face, float uv[2]);
void zro_aabb_min_max(const float c1[3], const float c2[3], float out_min[3], float out_max[3], float out_mid[3]);
void zro_aabb_vertexes(const float *xyz, uint32_t n, float out_min[3], float out_max[3]);
void zro_aabb_union(const float min1[3], const float max1[3], const float min2[3], const float max2[3],
                    float out_min[3], float out_max[3]);
float zro_aabb_surface_area(const float mn[3], const float mx[3]);
float zro_aabb_volume(const float mn[3], const float mx[3]);
int zro_aabb_hit(const float mn[3], const float mx[3], const float origin[3], const float direction[3],
                 float t_min, float t_max);
void zro_texture_albedo(const zrt_texture *tex, float u, float v, float out[3]);
/* which: 0 randomVector, 1 randomVectorInUnitSphere, 2 randomUnitVector_old, 3 randomUnitVector */
void zro_sample(int which, uint64_t seed, float out[3]);
/* material.scatter: returns 1 and the scattered ray/attenuation, or 0 if absorbed.  rnd[4] are the
 * ctr-RNG words (x,y,z,w) the scatter may consume. */
int zro_scatter(const zrt_scene_desc *desc, uint32_t material, const float origin[3], const float direction[3],
                const float location[3], const float normal[3], int front_face, const float uv[2],
                const uint32_t rnd[4], int math_mode, float out_origin[3], float out_direction[3],
                float attenuation[3]);
/* bvh.zig:262-291: n random spheres (seed), n_rays random rays -> number of rays that hit */
uint64_t zro_bvh_random_test(uint32_t n_spheres, uint32_t n_rays, uint64_t seed, int traversal_mode);
/* spec math kernels, elementwise: which 0 sincos(x)->(out,out2), 1 acos(x), 2 atan2(y,x), 3 pow5(x) */
void zro_math_eval(int which, const float *x, const float *y, float *out, float *out2, uint64_t n);
/* the counter-based generator itself (spec: DESIGN.md "RNG") */
void zro_rng_ctr(uint32_t pixel, uint32_t sample, uint32_t bounce, uint64_t seed, uint32_t out[4]);
/* PNG 8-bit quantisation of png_image.zig:136-140 */
uint8_t zro_quantize(float c);

#ifdef __cplusplus
}
#endif
#endif
