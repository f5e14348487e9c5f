// scenes.cpp — host mirror of camera.zig:17-35 and scenes.zig:26-277 (the six scene builders and
// render_scene).  Scene data is produced on the host exactly as the reference does and handed to the
// device path through the C ABI (zrt_scene_create / zrt_render).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>

#include <algorithm>
#include <cstring>

#include "../../../include/zrt_host.h"

struct zrt_host_scene {
    std::vector<zrt_surface> surfaces;
    std::vector<zrt_sphere> spheres;
    std::vector<zrt_triangle> triangles;
    std::vector<zrt_material> materials;
    std::vector<zrt_texture> textures;
    std::vector<uint8_t *> owned_pixels;
    std::vector<uint8_t *> pinned_pixels; // after zrt_host_scene_pin
    zrt_scene_desc desc{};
    zrt_camera camera{};
    ~zrt_host_scene() {
        for (uint8_t *p : owned_pixels) zrt_host_free(p);
        for (uint8_t *p : pinned_pixels) zrt_pinned_free(p);
    }
};

namespace {

struct V {
    float x, y, z;
};
V sub(V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; }
V scale(V a, float s) { return V{a.x * s, a.y * s, a.z * s}; }
V cross(V u, V v) { return V{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
V unit(V v) {
    const float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return V{v.x / len, v.y / len, v.z / len};
}

bool exists(const std::string &p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}
// reference layout first (./models/man/Man.obj), then this repository's (models/Man.obj.gz)
std::string findModel(const std::string &dir, const std::string &sub, const std::string &name) {
    const std::string cands[] = {dir + "/models/" + sub + "/" + name + ".obj", dir + "/models/" + name + ".obj",
                                 dir + "/models/" + name + ".obj.gz", dir + "/" + name + ".obj.gz"};
    for (const auto &c : cands)
        if (exists(c)) return c;
    return "";
}
std::string findImage(const std::string &dir, const std::string &name) {
    const std::string cands[] = {dir + "/models/images/" + name, dir + "/images/" + name, dir + "/" + name};
    for (const auto &c : cands)
        if (exists(c)) return c;
    return "";
}

// image.zig:14-20
const float SILVER[3] = {0.752f, 0.752f, 0.752f};
const float GREEN[3] = {0.01f, 1.0f, 0.01f};
const float BLUE[3] = {0.01f, 0.01f, 1.0f};

struct Builder {
    zrt_host_scene *s;
    std::string dir;
    int rc = ZRT_OK;

    uint32_t colorTexture(const float c[3]) { // texture.zig:11-13
        zrt_texture t{};
        t.kind = ZRT_TEXTURE_COLOR;
        t.r = c[0]; t.g = c[1]; t.b = c[2];
        s->textures.push_back(t);
        return (uint32_t)s->textures.size() - 1;
    }
    uint32_t imageTexture(const std::string &file) { // png_image.readFile + Texture.initImage (texture.zig:14-16)
        zrt_texture t{};
        t.kind = ZRT_TEXTURE_IMAGE;
        t.u_offset = 0.19f;
        t.v_offset = 0.1f;
        const std::string path = findImage(dir, file);
        uint8_t *px = nullptr;
        const int r = path.empty() ? ZRT_ERR_IO : zrt_host_png_read(path.c_str(), &px, &t.width, &t.height, &t.channels);
        if (r != ZRT_OK) { rc = r; t.kind = ZRT_TEXTURE_COLOR; }
        t.pixels = px;
        if (px) s->owned_pixels.push_back(px);
        s->textures.push_back(t);
        return (uint32_t)s->textures.size() - 1;
    }
    uint32_t material(uint32_t kind, uint32_t tex, float ior) {
        s->materials.push_back(zrt_material{kind, tex, ior});
        return (uint32_t)s->materials.size() - 1;
    }
    uint32_t metal(uint32_t tex) { return material(ZRT_MATERIAL_METAL, tex, 0.0f); }
    uint32_t lambertian(uint32_t tex) { return material(ZRT_MATERIAL_LAMBERTIAN, tex, 0.0f); }
    uint32_t dielectric(float ior) { return material(ZRT_MATERIAL_DIELECTRIC, 0, ior); }
    void sphere(float x, float y, float z, float r, uint32_t m) { // Surface.initSphere(Sphere.init(..))
        s->surfaces.push_back(zrt_surface{ZRT_SURFACE_SPHERE, (uint32_t)s->spheres.size()});
        s->spheres.push_back(zrt_sphere{zrt_vec3{x, y, z}, r, m});
    }
    void triangle(const zrt_triangle &t) {
        s->surfaces.push_back(zrt_surface{ZRT_SURFACE_TRIANGLE, (uint32_t)s->triangles.size()});
        s->triangles.push_back(t);
    }
    bool loadModel(const std::string &sub, const std::string &name, uint32_t m, std::vector<zrt_triangle> *out) {
        const std::string path = findModel(dir, sub, name);
        zrt_triangle *tris = nullptr;
        uint32_t n = 0;
        const int r = path.empty() ? ZRT_ERR_IO : zrt_host_read_obj(path.c_str(), m, &tris, &n);
        if (r != ZRT_OK) { rc = r; return false; }
        out->assign(tris, tris + n);
        zrt_host_free(tris);
        return true;
    }
    void ground(float top) { // "earth" sphere shared by the OBJ scenes, e.g. scenes.zig:42-46
        const float radius = 100.0f;
        sphere(1.66445508e-01f, top - radius, 7.37018966e+00f, radius, lambertian(colorTexture(GREEN)));
    }
    void camera(float fx, float fy, float fz, float aspect) { // every scene: look at z_unit, vup y, vfov 45
        const float from[3] = {fx, fy, fz}, at[3] = {0.0f, 0.0f, 1.0f}, up[3] = {0.0f, 1.0f, 0.0f};
        zrt_host_camera_init(from, at, up, 45.0f, aspect, &s->camera);
    }
};

// one midpoint subdivision step: each triangle becomes four (deterministic, used by the goat substitute)
void subdivide(std::vector<zrt_triangle> *tris) {
    std::vector<zrt_triangle> out;
    out.reserve(tris->size() * 4);
    auto mid = [](zrt_vec3 p, zrt_vec3 q) { return zrt_vec3{(p.x + q.x) * 0.5f, (p.y + q.y) * 0.5f, (p.z + q.z) * 0.5f}; };
    for (const zrt_triangle &t : *tris) {
        const zrt_vec3 ab = mid(t.a, t.b), bc = mid(t.b, t.c), ca = mid(t.c, t.a);
        out.push_back(zrt_triangle{t.a, ab, ca, t.material});
        out.push_back(zrt_triangle{ab, t.b, bc, t.material});
        out.push_back(zrt_triangle{ca, bc, t.c, t.material});
        out.push_back(zrt_triangle{ab, bc, ca, t.material});
    }
    tris->swap(out);
}

int buildScene(uint32_t index, uint32_t variant, float aspect, Builder &b) {
    std::vector<zrt_triangle> model;
    switch (index) {
    case 0: { // manAndBall scenes.zig:26-52
        const uint32_t blue = b.metal(b.colorTexture(BLUE));
        if (!b.loadModel("man", "Man", blue, &model)) return b.rc;
        b.ground(-2.33f);
        for (const auto &t : model) b.triangle(t);
        b.camera(0.0f, 0.0f, -30.0f, aspect);
        break;
    }
    case 1: { // threeBalls scenes.zig:54-100 — the 7-spheres showcase
        const uint32_t mirror = b.metal(b.colorTexture(SILVER));
        const uint32_t nitor = b.lambertian(b.imageTexture("nitor-logo-25.png"));
        const uint32_t green = b.lambertian(b.colorTexture(GREEN));
        const uint32_t glass = b.dielectric(1.52f);
        const uint32_t earth = b.metal(b.imageTexture("earthmap.png"));
        b.sphere(1.0f, -102.5f, 4.0f, 100.0f, green);
        b.sphere(0.0f, 0.0f, 8.0f, 2.0f, nitor);
        b.sphere(-3.0f, -1.5f, 3.0f, 1.0f, mirror);
        b.sphere(3.0f, -1.0f, 4.0f, 1.5f, earth);
        b.sphere(-1.0f, -1.0f, 2.0f, 0.7f, glass);
        b.sphere(0.85f, -0.7f, 1.5f, 0.9f, glass);
        b.sphere(0.85f, -0.7f, 1.5f, -0.8f, glass); // -(radius - thickness): hollow bubble
        b.camera(0.0f, 0.0f, -7.0f, aspect);
        break;
    }
    case 2: { // bunnyAndBall scenes.zig:102-128
        const uint32_t mat = (variant == ZRT_HOST_VARIANT_BUNNY_GLASS) ? b.dielectric(1.52f) : b.metal(b.colorTexture(SILVER));
        if (!b.loadModel("bunny", "bunny", mat, &model)) return b.rc;
        b.ground(-0.33f);
        for (const auto &t : model) b.triangle(t);
        b.camera(0.0f, 0.0f, -0.5f, aspect);
        break;
    }
    case 3: { // teapotAndBall scenes.zig:206-232
        const uint32_t blue = b.metal(b.colorTexture(BLUE));
        if (!b.loadModel("teapot", "teapot", blue, &model)) return b.rc;
        b.ground(-2.33f);
        for (const auto &t : model) b.triangle(t);
        b.camera(0.0f, 0.0f, -10.0f, aspect);
        break;
    }
    case 4: { // teapotAndBallCircle scenes.zig:130-204
        const uint32_t blue = b.metal(b.colorTexture(BLUE));
        const uint32_t silver = b.metal(b.colorTexture(SILVER));
        const uint32_t purple = b.lambertian(b.imageTexture("earthmap.png"));
        if (!b.loadModel("teapot", "teapot", blue, &model)) return b.rc;
        b.sphere(0.0f, 0.0f, 6.0f, -2.0f, silver);
        b.sphere(3.0f, -1.0f, 4.0f, 1.0f, purple);
        b.ground(-2.33f);
        for (const auto &t : model) b.triangle(t);
        b.camera(-8.0f, 0.0f, -10.0f, aspect);
        break;
    }
    case 5: { // goat scenes.zig:234-260; models/high_poly_goat.obj is not shipped with the reference
        if (variant != ZRT_HOST_VARIANT_GOAT_SUBSTITUTE) {
            const uint32_t silver = b.metal(b.colorTexture(SILVER));
            if (!b.loadModel("", "high_poly_goat", silver, &model)) return b.rc;
            b.ground(-2.33f);
            for (const auto &t : model) b.triangle(t);
            b.camera(0.0f, 0.0f, -1.7f, aspect);
            break;
        }
        // BASELINE config 4 "high_poly_goat.obj + man model with image textures": the goat is replaced by
        // bunny.obj subdivided three times (x64 triangles), scaled x40 and set on the ground next to Man.obj.
        const uint32_t man_mat = b.lambertian(b.imageTexture("earthmap.png"));
        const uint32_t goat_mat = b.metal(b.imageTexture("nitor-logo-25.png"));
        std::vector<zrt_triangle> man, goat;
        if (!b.loadModel("man", "Man", man_mat, &man)) return b.rc;
        if (!b.loadModel("bunny", "bunny", goat_mat, &goat)) return b.rc;
        for (int k = 0; k < 3; k++) subdivide(&goat);
        auto place = [](zrt_vec3 p) { return zrt_vec3{p.x * 40.0f + 11.0f, p.y * 40.0f - 3.66f, p.z * 40.0f + 8.0f}; };
        for (auto &t : goat) { t.a = place(t.a); t.b = place(t.b); t.c = place(t.c); }
        b.ground(-2.33f);
        for (const auto &t : man) b.triangle(t);
        for (const auto &t : goat) b.triangle(t);
        b.camera(0.0f, 0.0f, -30.0f, aspect);
        break;
    }
    default: return ZRT_ERR_INVALID; // SceneError.UnkownSceneIndex
    }
    return b.rc;
}

} // namespace

extern "C" {

int zrt_host_camera_init(const float look_from[3], const float look_at[3], const float vup[3], float vfov,
                         float aspect_ratio, zrt_camera *out) { // camera.zig:7-35
    if (!look_from || !look_at || !vup || !out) return ZRT_ERR_INVALID;
    const float theta = 3.14159265358979323846f * vfov / 180.0f; // deg2rad
    const float h = std::tan(theta / 2.0f);
    const float viewport_height = 2.0f * h;
    const float viewport_width = aspect_ratio * viewport_height;
    const V from{look_from[0], look_from[1], look_from[2]}, at{look_at[0], look_at[1], look_at[2]}, up{vup[0], vup[1], vup[2]};
    const V w = unit(sub(from, at));
    const V u = unit(cross(up, w));
    const V v = cross(w, u);
    const V horizontal = scale(u, viewport_width);
    const V vertical = scale(v, viewport_height);
    const V llc = sub(sub(sub(from, scale(horizontal, 1 / 2.0f)), scale(vertical, 1 / 2.0f)), w);
    if (std::isnan(w.x) || std::isnan(u.x)) return ZRT_ERR_INVALID; // std.debug.assert camera.zig:31-32
    out->origin = zrt_vec3{from.x, from.y, from.z};
    out->lower_left_corner = zrt_vec3{llc.x, llc.y, llc.z};
    out->horizontal = zrt_vec3{horizontal.x, horizontal.y, horizontal.z};
    out->vertical = zrt_vec3{vertical.x, vertical.y, vertical.z};
    return ZRT_OK;
}

int zrt_host_scene_load(uint32_t scene_index, const char *assets_dir, uint32_t variant, float aspect_ratio,
                        zrt_host_scene **out) {
    if (!out || !assets_dir) return ZRT_ERR_INVALID;
    *out = nullptr;
    zrt_host_scene *s = new zrt_host_scene();
    Builder b{s, assets_dir};
    const int rc = buildScene(scene_index, variant, aspect_ratio > 0.0f ? aspect_ratio : 1.0f, b);
    if (rc != ZRT_OK) {
        delete s;
        return rc;
    }
    s->desc.n_surfaces = (uint32_t)s->surfaces.size();   s->desc.surfaces = s->surfaces.data();
    s->desc.n_spheres = (uint32_t)s->spheres.size();     s->desc.spheres = s->spheres.data();
    s->desc.n_triangles = (uint32_t)s->triangles.size(); s->desc.triangles = s->triangles.data();
    s->desc.n_materials = (uint32_t)s->materials.size(); s->desc.materials = s->materials.data();
    s->desc.n_textures = (uint32_t)s->textures.size();   s->desc.textures = s->textures.data();
    *out = s;
    return ZRT_OK;
}

const zrt_scene_desc *zrt_host_scene_desc(const zrt_host_scene *scene) { return scene ? &scene->desc : nullptr; }
const zrt_camera *zrt_host_scene_camera(const zrt_host_scene *scene) { return scene ? &scene->camera : nullptr; }
void zrt_host_scene_free(zrt_host_scene *scene) { delete scene; }

int zrt_host_scene_pin(zrt_host_scene *scene) {
    if (!scene) return ZRT_ERR_INVALID;
    for (zrt_texture &t : scene->textures) {
        if (t.kind != ZRT_TEXTURE_IMAGE || !t.pixels) continue;
        if (std::find(scene->pinned_pixels.begin(), scene->pinned_pixels.end(), t.pixels) != scene->pinned_pixels.end()) continue;
        const size_t bytes = (size_t)t.width * t.height * t.channels;
        void *locked = nullptr;
        const int rc = zrt_pinned_alloc(bytes, &locked);
        if (rc != ZRT_OK) return rc;
        std::memcpy(locked, t.pixels, bytes);
        auto it = std::find(scene->owned_pixels.begin(), scene->owned_pixels.end(), t.pixels);
        if (it != scene->owned_pixels.end()) {
            zrt_host_free(*it);
            scene->owned_pixels.erase(it);
        }
        t.pixels = (const uint8_t *)locked;
        scene->pinned_pixels.push_back((uint8_t *)locked);
    }
    return ZRT_OK;
}

int zrt_host_render_scene(uint32_t scene_index, const char *assets_dir, uint32_t variant, const zrt_params *params,
                          int device, float *out_rgb, zrt_counters *counters, zrt_timing *timing) { // scenes.zig:267-277
    if (!params) return ZRT_ERR_INVALID;
    zrt_host_scene *hs = nullptr;
    int rc = zrt_host_scene_load(scene_index, assets_dir, variant, 1.0f, &hs); // every reference scene uses aspect 1.0
    if (rc != ZRT_OK) return rc;
    zrt_scene *sc = nullptr;
    rc = zrt_scene_create(&hs->desc, device, &sc);
    if (rc == ZRT_OK) {
        rc = zrt_render(sc, &hs->camera, params, out_rgb, counters, timing);
        zrt_scene_destroy(sc);
    }
    zrt_host_scene_free(hs);
    return rc;
}

} // extern "C"
