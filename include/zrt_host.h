/*
 * zrt_host.h — C entry points of the host-side mirror of the reference's scene / camera / OBJ / PNG API.
 * These live in the same libzrt.so; they never touch the GPU except zrt_host_render_scene, which calls
 * zrt_scene_create + zrt_render.  They exist so that the reference's scenes can be produced on the caller's
 * side of the boundary (SURVEY §2 "boundary (host)") and are what the CLI (zrt_cli) is made of.
 */
#ifndef ZRT_HOST_H
#define ZRT_HOST_H
#include "zrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Camera.init (camera.zig:17-35): look_from/look_at/vup, vertical fov in degrees, aspect ratio. */
int zrt_host_camera_init(const float look_from[3], const float look_at[3], const float vup[3], float vfov,
                         float aspect_ratio, zrt_camera *out);

/* ObjReader.readObjFile (obj_reader.zig:114-198): `v` / `f` lines, 1-based indices with optional
 * /vt/vn suffixes (ignored), faces of 3..6 vertices fan-triangulated (obj_reader.zig:64-111).
 * Reads plain or gzip-compressed files.  *triangles is malloc'ed; free with zrt_host_free. */
int zrt_host_read_obj(const char *path, uint32_t material, zrt_triangle **triangles, uint32_t *n_triangles);

/* png_image.readFile (png_image.zig:19-94): 8-bit RGB or RGBA, non-interlaced.  Rows are flipped so
 * that row 0 is the bottom scanline (png_image.zig:82-87); texels stay bytes (the device divides by 255).
 * *pixels is malloc'ed; free with zrt_host_free. */
int zrt_host_png_read(const char *path, uint8_t **pixels, uint32_t *width, uint32_t *height, uint32_t *channels);

/* png_image.writeFile (png_image.zig:96-148): 8-bit RGB, u8 = clamp(255.999 * c, 0, 255) truncated, no
 * gamma, image row 0 (bottom) written last. */
int zrt_host_png_write(const char *path, const float *rgb, uint32_t width, uint32_t height);

/* Same file format, from the 8-bit top-down image zrt_render_rgb8 returns (no further arithmetic). */
int zrt_host_png_write_rgb8(const char *path, const uint8_t *rgb8_top_down, uint32_t width, uint32_t height);

void zrt_host_free(void *p);

/* The reference's scene builders (scenes.zig:26-277), by the CLI's scene index:
 *   0 manAndBall  1 threeBalls (7-spheres)  2 bunnyAndBall  3 teapotAndBall  4 teapotAndBallCircle  5 goat
 * assets_dir holds either the reference layout (models/man/Man.obj, models/images/earthmap.png, ...) or
 * this repository's (models/Man.obj.gz, images/earthmap.png).  variant selects BASELINE.json's tweaks. */
enum {
    ZRT_HOST_VARIANT_REFERENCE = 0,
    ZRT_HOST_VARIANT_BUNNY_GLASS = 1,     /* scene 2 with the bunny material = Dielectric(1.52) (config 3) */
    ZRT_HOST_VARIANT_GOAT_SUBSTITUTE = 2  /* scene 5 with the missing goat replaced, see DESIGN.md (config 4) */
};
typedef struct zrt_host_scene zrt_host_scene;
int zrt_host_scene_load(uint32_t scene_index, const char *assets_dir, uint32_t variant, float aspect_ratio,
                        zrt_host_scene **out);
const zrt_scene_desc *zrt_host_scene_desc(const zrt_host_scene *scene);
const zrt_camera *zrt_host_scene_camera(const zrt_host_scene *scene);
void zrt_host_scene_free(zrt_host_scene *scene);
/* Move the scene's texel arrays into page-locked memory (zrt_pinned_alloc) on the calling thread's current device, so
 * that every later zrt_scene_create uploads them with one DMA each.  Needs a device (ZRT_ERR_NO_DEVICE otherwise, the
 * scene is left as it was); the description returned by zrt_host_scene_desc stays valid. */
int zrt_host_scene_pin(zrt_host_scene *scene);

/* scenes.render_scene (scenes.zig:267-277) followed by nothing else: build scene `scene_index`, render
 * it on `device` through zrt_render, return image and counters. */
int zrt_host_render_scene(uint32_t scene_index, const char *assets_dir, uint32_t variant, const zrt_params *params,
                          int device, float *out_rgb, zrt_counters *counters, zrt_timing *timing);

#ifdef __cplusplus
}
#endif
#endif
