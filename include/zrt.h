/*
 * zrt.h — C ABI of libzrt, the B200-native path-tracing core for zraytrace scenes.
 *
 * This header is the drop-in boundary for ONE function of the reference:
 *
 *     pub fn render(allocator, random, camera: Camera, surfaces: ArrayList(Surface),
 *                   render_params: RenderParams) !*Image          (src/raytrace.zig:136-138)
 *
 * Everything the reference does below that call (pixel/sample loop raytrace.zig:162-187, rayColor
 * :62-100, BVH bvh.zig:187-205, sphere.zig:31-71, triangle.zig:48-70, material.zig:43-128,
 * texture.zig:52-74) runs as hand-written sm_100a CUDA behind these entry points.  There is no CPU
 * fallback: without a CUDA device every compute entry point returns ZRT_ERR_NO_DEVICE.
 *
 * Plain C, POD structs, caller-owned host buffers in and out, no C++/torch types.
 * A Zig caller binds it exactly like the reference binds libpng (png_image.zig:6-9, build.zig:17-19):
 *     const c = @cImport({ @cInclude("zrt.h"); });   exe.linkSystemLibrary("zrt");
 */
#ifndef ZRT_H
#define ZRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZRT_ABI_VERSION 1

/* ---- status codes (Zig error union of render() -> int; never throws/aborts across the ABI) ---- */
enum {
    ZRT_OK = 0,
    ZRT_ERR_INVALID = -1,   /* bad argument / inconsistent scene description */
    ZRT_ERR_NO_DEVICE = -2, /* no CUDA device: there is deliberately no CPU fallback */
    ZRT_ERR_CUDA = -3,      /* a CUDA runtime call failed, see zrt_last_error() */
    ZRT_ERR_OOM = -4,       /* host allocation failed (error.OutOfMemory in the reference) */
    ZRT_ERR_IO = -5,        /* asset file could not be read / written (host helpers only) */
    ZRT_ERR_NCCL = -6       /* multi-GPU group: libnccl.so.2 could not be loaded, or an NCCL call failed */
};

/* ---- math PODs: vector.zig:22-26 (Vec3), base.zig:2 (BaseFloat = f32) ---- */
typedef struct zrt_vec3 { float x, y, z; } zrt_vec3;

/* ---- shapes ---- */
/* sphere.zig:15-20.  Negative radius is legal and flips the normal (scenes.zig:96). */
typedef struct zrt_sphere {
    zrt_vec3 center;
    float radius;
    uint32_t material; /* index into zrt_scene_desc.materials (reference: *const Material) */
} zrt_sphere;

/* triangle.zig:15-30.  Only a,b,c are given; e1,e2,face_normal are derived as triangle.zig:32-44. */
typedef struct zrt_triangle {
    zrt_vec3 a, b, c;
    uint32_t material;
} zrt_triangle;

/* surface.zig:12-15: the caller's ArrayList(Surface) in order.  The position in this list is the
 * surface id reported by zrt_primary_hits and decides ties exactly like raytrace.zig:75-81. */
enum { ZRT_SURFACE_SPHERE = 0, ZRT_SURFACE_TRIANGLE = 1 };
typedef struct zrt_surface {
    uint32_t kind;  /* ZRT_SURFACE_* */
    uint32_t index; /* into spheres[] or triangles[] */
} zrt_surface;

/* ---- materials and textures ---- */
/* material.zig:16-29 */
enum { ZRT_MATERIAL_LAMBERTIAN = 0, ZRT_MATERIAL_METAL = 1, ZRT_MATERIAL_DIELECTRIC = 2 };
typedef struct zrt_material {
    uint32_t kind;             /* ZRT_MATERIAL_* */
    uint32_t texture;          /* lambertian/metal: index into textures[]; ignored for dielectric */
    float index_of_refraction; /* dielectric only (material.zig:99-103) */
} zrt_material;

/* texture.zig:7-16,30-50 */
enum { ZRT_TEXTURE_COLOR = 0, ZRT_TEXTURE_IMAGE = 1 };
typedef struct zrt_texture {
    uint32_t kind;         /* ZRT_TEXTURE_* */
    float r, g, b;         /* ColorTexture.color */
    /* ImageTexture: 8-bit texels, `channels` (3 or 4) bytes per texel, row 0 = BOTTOM scanline,
     * i.e. already flipped the way png_image.readFile stores it (png_image.zig:82-87).  The device
     * converts byte/255.0f on lookup, bit-identical to png_image.zig:87. */
    uint32_t width, height, channels;
    const uint8_t *pixels;
    float u_offset, v_offset; /* texture.zig:14-16 default (0.19, 0.1) */
} zrt_texture;

/* ---- the scene: what the reference passes as `surfaces` plus everything reachable from it ---- */
typedef struct zrt_scene_desc {
    uint32_t n_surfaces;   const zrt_surface *surfaces;
    uint32_t n_spheres;    const zrt_sphere *spheres;
    uint32_t n_triangles;  const zrt_triangle *triangles;
    uint32_t n_materials;  const zrt_material *materials;
    uint32_t n_textures;   const zrt_texture *textures;
} zrt_scene_desc;

/* camera.zig:11-15 (the result of Camera.init :17-35, computed by the caller on the host) */
typedef struct zrt_camera {
    zrt_vec3 origin, lower_left_corner, horizontal, vertical;
} zrt_camera;

/* raytrace.zig:102-108 RenderParams, widened to u32 (main.zig parses u16), plus the extensions the
 * B200 path needs.  Zero in an extension field selects the reference behaviour. */
enum { ZRT_XLIMIT_HEIGHT = 0, ZRT_XLIMIT_WIDTH = 1 };
typedef struct zrt_params {
    uint32_t width, height, samples_per_pixel, max_depth;
    uint32_t bounded_volume_hierarchy; /* BVH is used iff this != 0 AND n_surfaces > 10 (raytrace.zig:127) */
    /* --- extensions --- */
    uint32_t x_limit;      /* ZRT_XLIMIT_HEIGHT reproduces `while (x < image.height)` raytrace.zig:168 */
    uint64_t seed;         /* key of the counter-based RNG that replaces the shared PRNG (scenes.zig:60) */
    uint32_t sample_begin; /* global sample range [begin,end) traced by this call; 0,0 = all */
    uint32_t sample_end;   /*   (used to split samples-per-pixel across GPUs) */
    uint32_t flags;        /* ZRT_FLAG_* */
    uint32_t sample_chunks;/* 0 = auto.  N > 0: split each pixel's samples over N device threads whose
                              partial sums are added in chunk order; 1 keeps the reference's sequential
                              f32 accumulation order raytrace.zig:177 */
} zrt_params;

enum {
    ZRT_FLAG_RAW_SUM = 1u << 0, /* output the un-normalised sum over the traced samples (for multi-GPU
                                   reduction) instead of sum * (1/samples_per_pixel) raytrace.zig:157,182 */
    ZRT_FLAG_BVH_REFERENCE = 1u << 1, /* traverse the flattened topology of the reference's own tree
                                   (bvh.zig:62-185) instead of the binned-SAH tree libzrt builds by default over
                                   the same primitives; hits are identical either way (ties break on the
                                   reference DFS order, unreachable surfaces are pruned), only speed differs */
    ZRT_FLAG_KERNEL_THREAD = 1u << 2, /* force the one-thread-per-path megakernel k_trace (wins over SORTED) */
    ZRT_FLAG_KERNEL_WARP = 1u << 4,   /* BVH scenes: the warp-scheduled state machine k_trace_ws (bit-identical
                                   output; 13 % / 7 % faster than k_trace on configs 2 / 4, 3 % slower on config 3).
                                   With neither flag a BVH launch of >= 2^24 samples runs k_trace_ws, a smaller one
                                   k_trace; sphere-only and list scenes always run k_trace */
    ZRT_FLAG_RUSSIAN_ROULETTE = 1u << 5, /* src/README.md:5-6 TODO of the reference: from the 3rd ray of a path on,
                                   continue with probability p = clamp(max(throughput), 0.05, 1) and divide the
                                   throughput by p.  Changes the estimator's variance, not its expectation */
    ZRT_FLAG_SAMPLER_HALTON = 1u << 6,   /* src/README.md:8-13 TODO: pixel jitter from the Halton (2,3) sequence over
                                   the global sample index, Cranley-Patterson rotated per pixel, instead of two
                                   independent uniforms.  Both extensions run on k_trace only and are restated by
                                   the oracle (draw-for-draw parity); zrt_trace_statistics ignores them */
    ZRT_FLAG_KERNEL_X2 = 1u << 7,     /* spheres-only scenes: k_trace_x2, two paths per thread in packed f32x2 (opt-in: ties) */
    ZRT_FLAG_KERNEL_POOL = 1u << 8,   /* the slot-pool kernels: every warp keeps a pool of work items in shared memory and runs
                                   batches of up to 32 paths that need the same thing next (new sample, Lambertian, metal,
                                   glass, image-textured variants); bit-identical output.  Sphere-only scenes: k_trace_pool3,
                                   which is also what launches of >= 2^24 samples run without any flag (36 ms against
                                   40.4 ms of k_trace on the README headline).  BVH scenes: k_trace_bpool (traversal lanes
                                   re-armed from a ring), measured SLOWER than k_trace_ws and therefore only on request.
                                   Ignored on surface-list scenes with triangles */
    ZRT_FLAG_KERNEL_SORTED = 1u << 3  /* the block-sorted-shading kernel k_trace_sorted (shared-memory wavefront
                                   inside a thread block).  Images and counters are bit-identical between the
                                   two kernels, only speed differs; the thread kernel measured faster */
};

/* raytrace.zig:20-34 Progress (the six u64 counters), same semantics */
typedef struct zrt_counters {
    uint64_t recursion_depth_hits;
    uint64_t reflections;
    uint64_t background_hits;
    uint64_t pixels_processed;
    uint64_t samples_processed;
    uint64_t rays_processed;
} zrt_counters;

/* device timings of the last call, milliseconds from CUDA events on the launching stream */
typedef struct zrt_timing {
    float prepare_ms; /* host flatten / BVH build + H2D upload ("Prepare runtime" raytrace.zig:200) */
    float kernel_ms;  /* path-tracing kernel(s) only */
    float resolve_ms; /* chunk sum + 1/spp scale */
    float total_ms;   /* first launch to last D2H copy done */
    uint32_t launches;/* kernels launched by this call */
    uint32_t bvh_nodes;
} zrt_timing;

typedef struct zrt_scene zrt_scene; /* opaque: flattened scene resident in HBM of one device */

/* ---- entry points ---- */

/* Number of CUDA devices visible (0 = none; compute entry points will fail with ZRT_ERR_NO_DEVICE). */
int zrt_device_count(void);

/* Message of the last error on the calling thread ("" if none). */
const char *zrt_last_error(void);

/* Copy + flatten the caller's surface list (raytrace.zig:150 preprocessSufraces, bvh.zig:171-185)
 * and upload it to `device`.  The description may be freed as soon as this returns. */
int zrt_scene_create(const zrt_scene_desc *desc, int device, zrt_scene **out);
void zrt_scene_destroy(zrt_scene *scene);

/* raytrace.render(): out_rgb is width*height*3 floats, pixel (x,y) at (y*width+x)*3, row 0 = bottom
 * scanline (image.zig:74-103 as written by raytrace.zig:180-182).  counters/timing may be NULL. */
int zrt_render(zrt_scene *scene, const zrt_camera *camera, const zrt_params *params,
               float *out_rgb, zrt_counters *counters, zrt_timing *timing);

/* Page-locked host memory for the image that zrt_render / zrt_render_rgb8 / zrt_primary_hits fill (what the reference
 * takes from its allocator in Image.init, image.zig:74-84): the copy back is then one DMA instead of a staged copy
 * through the driver's bounce buffer into pages the host touches for the first time (41.4 instead of 44.4 ms end to
 * end on the 12 MB headline image).  Any host pointer works with the render calls; these two are for hosts that do
 * not link the CUDA runtime themselves.  zrt_pinned_free(NULL) is a no-op. */
int zrt_pinned_alloc(size_t bytes, void **out);
void zrt_pinned_free(void *ptr);

/* raytrace.render() + the quantisation of png_image.writeFile (png_image.zig:131-142) fused on the device:
 * out_rgb8 is width*height*3 bytes, u8 = clamp(255.999 * c, 0, 255) truncated, row 0 = TOP scanline (the order
 * a PNG stores).  Moves a quarter of the bytes of zrt_render back to the host; the bytes are identical to
 * quantising zrt_render's float image on the host. */
int zrt_render_rgb8(zrt_scene *scene, const zrt_camera *camera, const zrt_params *params,
                    uint8_t *out_rgb8, zrt_counters *counters, zrt_timing *timing);

/* Same work, but results stay on the device: d_rgb (width*height*3 floats) and d_counters (6 u64)
 * are DEVICE pointers on the scene's device, `stream` is a cudaStream_t (NULL = default stream).
 * Asynchronous: returns after the launches are enqueued.  This is what the multi-GPU driver uses so
 * that the accumulators can be summed with one NCCL reduce without touching the host. */
int zrt_render_device(zrt_scene *scene, const zrt_camera *camera, const zrt_params *params,
                      float *d_rgb, uint64_t *d_counters, void *stream);

/* ---- multi-GPU: raytrace.render() over W devices (BASELINE.json north_star) -------------------------
 * The scene is replicated on every device; rank g of W traces the global samples [g*spp/W, (g+1)*spp/W) of every
 * pixel; the f32 accumulators are summed with ONE ncclReduce (plus one of the six u64 counters) over NVLink to
 * rank 0, which applies the reference's 1/samples_per_pixel (raytrace.zig:157,182).  The RNG is keyed on the global
 * sample index: the counters do not depend on W, the image only through the association of the f32 sum.
 * NCCL is loaded at run time (dlopen "libnccl.so.2") and only by these calls.
 *   zrt_multi_create       one process drives n_devices GPUs (devices = NULL: 0..n-1): ncclCommInitAll.
 *   zrt_multi_create_rank  one process per GPU: every process calls it with the same id (made by zrt_comm_id on one of
 *                          them and distributed by the caller - MPI, torch.distributed, a file) and its own rank.
 *   zrt_multi_render       zrt_render for the group.  Every process of the group calls it with the same arguments; image
 *                          and counters arrive on the process that owns rank 0 (out_rgb may be NULL elsewhere, and on
 *                          rank 0 to leave the image on the device: zrt_multi_image_device).  timing: kernel_ms = slowest
 *                          local rank's trace, total_ms = first launch -> reduced, scaled image on the device.
 * Callable from one host thread at a time per group. */
typedef struct zrt_multi zrt_multi;
#define ZRT_COMM_ID_BYTES 128
int zrt_comm_id(uint8_t id[ZRT_COMM_ID_BYTES]);
int zrt_multi_create(const zrt_scene_desc *desc, const int *devices, int n_devices, zrt_multi **out);
int zrt_multi_create_rank(const zrt_scene_desc *desc, int device, const uint8_t id[ZRT_COMM_ID_BYTES], int rank, int world,
                          zrt_multi **out);
void zrt_multi_destroy(zrt_multi *group);
/* A new scene for the same group (communicator and accumulators are kept): zrt_scene_create on every local device. */
int zrt_multi_reload(zrt_multi *group, const zrt_scene_desc *desc);
int zrt_multi_render(zrt_multi *group, const zrt_camera *camera, const zrt_params *params, float *out_rgb,
                     zrt_counters *counters, zrt_timing *timing);
int zrt_multi_world_size(const zrt_multi *group);
uint64_t zrt_multi_launch_count(const zrt_multi *group);     /* libzrt kernels launched by the local ranks */
const float *zrt_multi_image_device(const zrt_multi *group); /* rank 0's device copy of the last image (NULL elsewhere) */
int zrt_nccl_version(void);                                  /* NCCL_VERSION_CODE of the loaded library, 0 if none */

/* Optional pieces compiled into this build of the library. */
enum { ZRT_FEATURE_EXPERIMENTS = 1u << 0 /* k_trace_sorted / k_trace_x2 (make EXPERIMENTS=1); without it their flags fail */ };
uint32_t zrt_build_features(void);

/* Number of libzrt kernels launched on behalf of this scene since it was created (trace, resolve and
 * primary-hit kernels; driver memsets and copies are not counted). */
uint64_t zrt_scene_launch_count(const zrt_scene *scene);

/* Parity AOV: first iteration of rayColor (raytrace.zig:71-81) for every pixel.  surface_id[y*w+x] is
 * the index in desc.surfaces of the closest hit (0xFFFFFFFF = background), t its ray parameter.
 * jitter = 0 traces (u,v) = ((x-0.5)/w,(y-0.5)/h), i.e. raytrace.zig:173-174 with xi = 0;
 * jitter = 1 uses the RNG draw of sample index params->sample_begin. */
#define ZRT_NO_HIT 0xFFFFFFFFu
int zrt_primary_hits(zrt_scene *scene, const zrt_camera *camera, const zrt_params *params,
                     int jitter, uint32_t *surface_id, float *t);

/* Statistics of the flattened acceleration structure (0 nodes when the list path is used). */
typedef struct zrt_bvh_info {
    uint32_t nodes, leaves, max_depth; /* flattened tree that the device traverses */
    uint32_t pruned_surfaces;          /* surfaces under a zero-thickness box: unreachable in the reference (Q4) */
    uint32_t reference_nodes, reference_max_depth; /* the bvh.zig tree it was derived from */
} zrt_bvh_info;
int zrt_scene_bvh_info(zrt_scene *scene, uint32_t flags, zrt_bvh_info *out);

/* Inspection hook for the parity tests: order[k] = surface id at left-first DFS position k of the
 * reference tree (the tie-break key), visible[id] = 0 for pruned surfaces.  Both n_surfaces long.
 * Works on a scene created with device = -1 (host only; such a scene cannot render). */
int zrt_scene_bvh_order(zrt_scene *scene, uint32_t *order, uint8_t *visible);

/* Event counts of one render, measured by an instrumented build of the same kernel (slower; the image is
 * discarded).  They are the byte side of the roofline: 64 B per node visit, 48 B per triangle test, 32 B per
 * sphere test, 3-4 B per texture lookup, 12 B per sample (DESIGN.md "Measurement"). */
typedef struct zrt_trace_stats {
    uint64_t rays, samples, node_visits, triangle_tests, sphere_tests, texture_lookups;
} zrt_trace_stats;
int zrt_trace_statistics(zrt_scene *scene, const zrt_camera *camera, const zrt_params *params, zrt_trace_stats *out);

/* Device self-test: the exact-quotient fast paths used by the kernels (shared reciprocal + FMA residuals)
 * against the compiler's IEEE division, over every raytrace.zig:173 numerator of ten image widths and
 * 2^23 random vectors.  *mismatches must come back 0. */
int zrt_selftest(int device, uint64_t *mismatches);

/* K0 microbenchmarks that measure the roofline denominators this path is judged against
 * (MEASURED_PEAKS.json has no FP32-issue or L2 number): results in out[0..n).
 *   out[0] fp32 non-FMA op/s (FMUL+FADD chains), out[1] FFMA op/s (1 per instr),
 *   out[2] L2-resident read GB/s, out[3] HBM streaming read GB/s, out[4] SM clock MHz seen */
int zrt_measure_peaks(int device, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* ZRT_H */
