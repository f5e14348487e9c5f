// zrt_api.cu — implementation of the C ABI declared in include/zrt.h.
// Replaces raytrace.render() (raytrace.zig:136-203): validates and copies the caller's scene, flattens it
// (list / BVH), keeps it resident in HBM, launches the sm_100a kernels and returns image + counters.
// There is no CPU path in this file: without a device the compute entry points fail with
// ZRT_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <mutex>

#include "zrt_scene.h"

namespace zrt {
uint32_t launch_trace(const KParams &P, int mode, cudaStream_t st);
void launch_primary(const KParams &P, int mode, cudaStream_t st);
void launch_resolve(const float *part, float *out, uint32_t n, uint32_t chunks, float scale, cudaStream_t st);
void launch_resolve_rgb8(const float *part, uint8_t *out, uint32_t width, uint32_t height, uint32_t chunks, float scale,
                         cudaStream_t st);
void launch_selftest_div(unsigned long long *mismatch, uint32_t width, uint32_t seed, cudaStream_t st);
void launch_peak_fp32(float *out, int blocks, int threads, int iters, cudaStream_t st);
void launch_peak_ffma(float *out, int blocks, int threads, int iters, cudaStream_t st);
void launch_peak_read(const float4 *src, size_t n4, int passes, float *out, int blocks, int threads, cudaStream_t st);
} // namespace zrt

using namespace zrt;

namespace zrt {

static thread_local std::string g_error;

int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

// Device buffers come from the device's default stream-ordered memory pool (cudaMallocAsync) with the release
// threshold raised, so creating and destroying scenes or scratch images does not pay cudaMalloc/cudaFree
// (the 100+ ms spikes seen in the first end-to-end measurements) after the first use.
cudaStream_t g_alloc_stream(int device) {
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[device]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking);
        cudaSetDevice(cur);
    }
    return streams[device];
}

} // namespace zrt


namespace zrt {

uint32_t packMaterial(const HostScene &hs, uint32_t m) {
    const zrt_material &mat = hs.materials[m];
    uint32_t w = (m & MAT_INDEX_MASK) | (mat.kind << MAT_KIND_SHIFT);
    if (mat.kind != ZRT_MATERIAL_DIELECTRIC && hs.textures[mat.texture].kind == ZRT_TEXTURE_IMAGE) w |= MAT_IMAGE_BIT;
    return w;
}

DevSphere makeSphere(const HostScene &hs, uint32_t surface, uint32_t slot) {
    const zrt_sphere &s = hs.spheres[hs.surfaces[surface].index];
    DevSphere d;
    d.cx = s.center.x; d.cy = s.center.y; d.cz = s.center.z;
    d.r2 = s.radius * s.radius;  // sphere.zig:34
    d.inv_r = 1.0f / s.radius;   // sphere.zig:46
    d.material = packMaterial(hs, s.material);
    d.surface_id = surface;
    d.slot = slot;
    return d;
}

void makeTriangle(const HostScene &hs, uint32_t surface, float4 *A, float4 *E1, float4 *E2, TriMeta *meta) {
    const zrt_triangle &t = hs.triangles[hs.surfaces[surface].index];
    // triangle.zig:32-44: e1 = b - a, e2 = c - a, face_normal = e1 x e2 (vector.zig:70-74)
    const float e1x = t.b.x - t.a.x, e1y = t.b.y - t.a.y, e1z = t.b.z - t.a.z;
    const float e2x = t.c.x - t.a.x, e2y = t.c.y - t.a.y, e2z = t.c.z - t.a.z;
    const float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    *A = float4{t.a.x, t.a.y, t.a.z, nx};
    *E1 = float4{e1x, e1y, e1z, ny};
    *E2 = float4{e2x, e2y, e2z, nz};
    meta->material = packMaterial(hs, t.material);
    meta->surface_id = surface;
}

int validate(const zrt_scene_desc *d) {
    if (!d) return fail(ZRT_ERR_INVALID, "scene description is NULL");
    if (d->n_surfaces && !d->surfaces) return fail(ZRT_ERR_INVALID, "surfaces is NULL");
    if (d->n_spheres && !d->spheres) return fail(ZRT_ERR_INVALID, "spheres is NULL");
    if (d->n_triangles && !d->triangles) return fail(ZRT_ERR_INVALID, "triangles is NULL");
    if (d->n_materials && !d->materials) return fail(ZRT_ERR_INVALID, "materials is NULL");
    if (d->n_textures && !d->textures) return fail(ZRT_ERR_INVALID, "textures is NULL");
    if (d->n_surfaces > REF_INDEX_MASK || d->n_materials > MAT_INDEX_MASK) return fail(ZRT_ERR_INVALID, "scene too large");
    for (uint32_t i = 0; i < d->n_surfaces; i++) {
        const zrt_surface &s = d->surfaces[i];
        if (s.kind == ZRT_SURFACE_SPHERE) {
            if (s.index >= d->n_spheres) return fail(ZRT_ERR_INVALID, "surface refers to a missing sphere");
            if (d->spheres[s.index].material >= d->n_materials) return fail(ZRT_ERR_INVALID, "sphere material out of range");
        } else if (s.kind == ZRT_SURFACE_TRIANGLE) {
            if (s.index >= d->n_triangles) return fail(ZRT_ERR_INVALID, "surface refers to a missing triangle");
            if (d->triangles[s.index].material >= d->n_materials) return fail(ZRT_ERR_INVALID, "triangle material out of range");
        } else {
            return fail(ZRT_ERR_INVALID, "unknown surface kind");
        }
    }
    // non-finite coordinates break the strict weak ordering of the BVH builder's sorts (undefined behaviour)
    auto finite3 = [](const zrt_vec3 &v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); };
    for (uint32_t i = 0; i < d->n_spheres; i++)
        if (!finite3(d->spheres[i].center) || !std::isfinite(d->spheres[i].radius))
            return fail(ZRT_ERR_INVALID, "sphere with a non-finite centre or radius");
    for (uint32_t i = 0; i < d->n_triangles; i++)
        if (!finite3(d->triangles[i].a) || !finite3(d->triangles[i].b) || !finite3(d->triangles[i].c))
            return fail(ZRT_ERR_INVALID, "triangle with a non-finite vertex");
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const zrt_material &m = d->materials[i];
        if (m.kind > ZRT_MATERIAL_DIELECTRIC) return fail(ZRT_ERR_INVALID, "unknown material kind");
        if (m.kind != ZRT_MATERIAL_DIELECTRIC && m.texture >= d->n_textures) return fail(ZRT_ERR_INVALID, "material texture out of range");
    }
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const zrt_texture &t = d->textures[i];
        if (t.kind == ZRT_TEXTURE_IMAGE) {
            if (!t.pixels || t.width == 0 || t.height == 0 || (t.channels != 3 && t.channels != 4))
                return fail(ZRT_ERR_INVALID, "image texture needs 8-bit RGB/RGBA pixels");
        } else if (t.kind != ZRT_TEXTURE_COLOR) {
            return fail(ZRT_ERR_INVALID, "unknown texture kind");
        }
    }
    return ZRT_OK;
}

int uploadMaterials(zrt_scene *sc) {
    const HostScene &hs = sc->host;
    std::vector<DevMaterial> mats(hs.materials.size());
    sc->d_texels.resize(hs.textures.size());
    for (size_t i = 0; i < hs.textures.size(); i++) {
        const zrt_texture &t = hs.textures[i];
        if (t.kind != ZRT_TEXTURE_IMAGE) continue;
        // straight from the caller's pixels (page-locked or not): a device scene keeps no host copy of the texels
        CUDA_TRY(sc->d_texels[i].upload(t.pixels, (size_t)t.width * t.height * t.channels, sc->stream));
    }
    for (size_t i = 0; i < mats.size(); i++) {
        const zrt_material &m = hs.materials[i];
        DevMaterial d{};
        d.kind = m.kind;
        d.ior = m.index_of_refraction;
        d.inv_ior = 1.0f / m.index_of_refraction; // material.zig:111
        d.r0_front = (1.0f - d.inv_ior) / (1.0f + d.inv_ior); // material.zig:126 with ratio = 1/ior
        d.r0_back = (1.0f - d.ior) / (1.0f + d.ior);          //                   and ratio = ior
        if (m.kind != ZRT_MATERIAL_DIELECTRIC) {
            const zrt_texture &t = hs.textures[m.texture];
            d.tex_kind = t.kind;
            d.r = t.r; d.g = t.g; d.b = t.b;
            d.u_off = t.u_offset; d.v_off = t.v_offset;
            d.w = t.width; d.h = t.height; d.ch = t.channels;
            d.pixels = sc->d_texels[m.texture].p;
        }
        mats[i] = d;
    }
    CUDA_TRY(sc->mats.upload(mats, sc->stream));
    CUDA_TRY(cudaStreamSynchronize(sc->stream)); // visible to any stream a later zrt_render_device is given
    return ZRT_OK;
}

// ---- flattening (host, once per scene and representation) and upload (once per device) ----------------------------
static std::shared_ptr<HostRep> flattenList(const HostScene &hs, bool all_spheres) {
    auto h = std::make_shared<HostRep>();
    const uint32_t n = (uint32_t)hs.surfaces.size();
    h->list.resize(n);
    if (hs.triangles.size()) { h->A.resize(n); h->E1.resize(n); h->E2.resize(n); h->meta.resize(n); }
    for (uint32_t i = 0; i < n; i++) {
        if (hs.surfaces[i].kind == ZRT_SURFACE_SPHERE) {
            h->list[i] = REF_LEAF | REF_SPHERE | (uint32_t)h->spheres.size();
            h->spheres.push_back(makeSphere(hs, i, i));
            if (hs.triangles.size()) { h->A[i] = h->E1[i] = h->E2[i] = float4{0, 0, 0, 0}; h->meta[i] = TriMeta{0, i}; }
        } else {
            h->list[i] = REF_LEAF | i; // triangle planes are indexed by list position
            makeTriangle(hs, i, &h->A[i], &h->E1[i], &h->E2[i], &h->meta[i]);
        }
    }
    h->n_list = n;
    h->mode = (all_spheres && n <= MAX_INLINE_SPHERES && n > 0) ? MODE_SPHERES : MODE_LIST;
    return h;
}

static std::shared_ptr<HostRep> flattenBvh(const HostScene &hs, bool sah) {
    auto h = std::make_shared<HostRep>();
    build_flat_bvh(hs, sah, &h->info);
    BuildLap lap;
    const uint32_t slots = (uint32_t)h->info.slot_surface.size();
    h->A.resize(slots); h->E1.resize(slots); h->E2.resize(slots); h->meta.resize(slots);
    // sphere_seq in zrt_flatten.cpp numbers spheres in surface-list order
    std::vector<uint32_t> slot_of(hs.surfaces.size(), 0);
    for (uint32_t s = 0; s < slots; s++) slot_of[h->info.slot_surface[s]] = s;
    for (uint32_t i = 0; i < hs.surfaces.size(); i++)
        if (hs.surfaces[i].kind == ZRT_SURFACE_SPHERE) h->spheres.push_back(makeSphere(hs, i, slot_of[i]));
    HostRep *hp = h.get();
    parallelFor(slots, 16384, [&hs, hp](size_t begin, size_t end) {
        for (size_t s = begin; s < end; s++) {
            const uint32_t surf = hp->info.slot_surface[s];
            if (hs.surfaces[surf].kind == ZRT_SURFACE_TRIANGLE) makeTriangle(hs, surf, &hp->A[s], &hp->E1[s], &hp->E2[s], &hp->meta[s]);
            else { hp->A[s] = hp->E1[s] = hp->E2[s] = float4{0, 0, 0, 0}; hp->meta[s] = TriMeta{0, surf}; }
        }
    });
    lap("pack primitives");
    h->root = h->info.root;
    h->mode = MODE_BVH;
    return h;
}

static int uploadRep(zrt_scene *sc, DevRep &r) {
    const HostRep &h = *r.host;
    if (h.mode == MODE_BVH && h.info.max_depth + 2 >= (uint32_t)TRAVERSAL_STACK)
        return fail(ZRT_ERR_INVALID, "BVH deeper than the traversal stack; drop ZRT_FLAG_BVH_REFERENCE");
    BuildLap lap;
    cudaStream_t st = sc->stream;
    CUDA_TRY(r.spheres.upload(h.spheres, st));
    CUDA_TRY(r.list.upload(h.list, st));
    CUDA_TRY(r.triA.upload(h.A.data(), h.A.size(), st));
    CUDA_TRY(r.triE1.upload(h.E1.data(), h.E1.size(), st));
    CUDA_TRY(r.triE2.upload(h.E2.data(), h.E2.size(), st));
    CUDA_TRY(r.triMeta.upload(h.meta.data(), h.meta.size(), st));
    CUDA_TRY(r.nodes.upload(h.info.nodes.data(), h.info.nodes.size(), st));
    CUDA_TRY(cudaStreamSynchronize(st)); // visible to every stream; the host arrays may go away
    lap("upload");
    r.mode = h.mode;
    r.n_spheres = (uint32_t)h.spheres.size();
    r.n_list = h.n_list;
    r.root = h.root;
    r.ready = true;
    return ZRT_OK;
}

DevRep *repFor(zrt_scene *sc, const zrt_params *p) {
    const bool use_bvh = p->bounded_volume_hierarchy != 0 && sc->host.surfaces.size() > 10;
    return !use_bvh ? &sc->rep_list : ((p->flags & ZRT_FLAG_BVH_REFERENCE) ? &sc->rep_bvh : &sc->rep_sah);
}

// raytrace.zig:111-133 preprocessSufraces: BVH iff flag and more than 10 surfaces.  A representation whose host half
// was flattened elsewhere (r->host set by zrt_multi: one flattening for all replicas) is only uploaded here.
int selectRep(zrt_scene *sc, const zrt_params *p, DevRep **out) {
    DevRep *r = repFor(sc, p);
    if (!r->ready) {
        const auto t0 = std::chrono::steady_clock::now();
        int rc;
        try { // nothing may throw across the C ABI
            if (!r->host) r->host = (r == &sc->rep_list) ? flattenList(sc->host, sc->all_spheres) : flattenBvh(sc->host, r == &sc->rep_sah);
            rc = uploadRep(sc, *r);
        } catch (const std::bad_alloc &) {
            return fail(ZRT_ERR_OOM, "out of host memory while flattening the scene");
        } catch (const std::exception &e) {
            return fail(ZRT_ERR_OOM, std::string("scene flattening failed: ") + e.what());
        }
        if (rc != ZRT_OK) return rc;
        r->prepare_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    *out = r;
    return ZRT_OK;
}

int makePlan(zrt_scene *sc, const zrt_camera *cam, const zrt_params *p, DevRep *r, Plan *plan) {
    if (!cam || !p) return fail(ZRT_ERR_INVALID, "camera/params is NULL");
    if (p->width == 0 || p->height == 0) return fail(ZRT_ERR_INVALID, "empty image");
    if ((uint64_t)p->width * p->height > 0x7FFFFFFFull / 3) return fail(ZRT_ERR_INVALID, "image too large");
    KParams &P = plan->P;
    std::memset(&P, 0, sizeof(P));
    P.ox = cam->origin.x; P.oy = cam->origin.y; P.oz = cam->origin.z;
    P.llx = cam->lower_left_corner.x; P.lly = cam->lower_left_corner.y; P.llz = cam->lower_left_corner.z;
    P.hx = cam->horizontal.x; P.hy = cam->horizontal.y; P.hz = cam->horizontal.z;
    P.vx = cam->vertical.x; P.vy = cam->vertical.y; P.vz = cam->vertical.z;
    P.width = p->width; P.height = p->height;
    P.f_width = (float)p->width; P.f_height = (float)p->height; // raytrace.zig:153-154
    P.rcp_width = 1.0f / P.f_width; P.rcp_height = 1.0f / P.f_height;
    // raytrace.zig:168 `while (x < image.height)`; when height > width the reference writes
    // past the end of the row, which is clamped here
    P.x_end = (p->x_limit == ZRT_XLIMIT_WIDTH) ? p->width : (p->height < p->width ? p->height : p->width);
    uint32_t sb = p->sample_begin, se = p->sample_end;
    if (sb == 0 && se == 0) se = p->samples_per_pixel;
    if (se < sb || se > p->samples_per_pixel) return fail(ZRT_ERR_INVALID, "bad sample range");
    P.s_begin = sb; P.s_end = se;
    plan->n_samples = se - sb;
    const uint64_t pixels = (uint64_t)P.x_end * P.height;
    // L slices per pixel (see k_trace).  Two costs pull in opposite directions: every item costs ~68 warp instructions
    // to hand over (fewer, longer items are cheaper), and the launch cannot end before its longest item, which sits on
    // the most expensive pixels (glass: ~7x the average rays per sample), so items must stay a small fraction of the
    // launch: L ~ 54 x resident lanes / pixels, independent of the sample count.  Measured on the 7-spheres scene
    // (best L): 2000^2 -> 2 (16.96 ms against 18.12 with 8), 1400^2 -> 4, 1000^2 -> 8 at 125..1000 spp, 600^2 -> 16,
    // 500^2 and below -> 32.  BVH / surface-list scenes: 32, i.e. the 32 lanes of a warp trace 32 slices of ONE pixel,
    // so their primary rays walk the same nodes (measured against 8: C2 13.2 -> 11.0 ms, C3 36.0 -> 34.2, C4 119 -> 101).
    // The caller can pin it (sample_chunks = 1 reproduces the reference's sequential f32 sum per pixel).
    // k_trace_pool (K1q) holds 4 x P.pool items per block instead of 128: the same rule over its slot count.
    // the slot word of k_trace_pool packs the sample index in 17 bits, the bounce count in 8, pixel coordinates in 16 each
    const bool ext = (p->flags & (ZRT_FLAG_SAMPLER_HALTON | ZRT_FLAG_RUSSIAN_ROULETTE)) != 0;
#if defined(ZRT_EXPERIMENTS) || defined(ZRT_EMU)
    const bool pool_modes = r->mode == MODE_SPHERES || r->mode == MODE_BVH; // k_trace_bpool is an experiment (slower than k_trace_ws)
#else
    const bool pool_modes = r->mode == MODE_SPHERES;
#endif
    const bool pool_ok = pool_modes && p->max_depth < 255u && p->samples_per_pixel < 65536u &&
                         p->width < 65536u && p->height < 65536u && !ext && !(p->flags & ZRT_FLAG_KERNEL_SORTED);
    uint32_t pool = 0;
    // Sphere-only scenes: launches that are mostly steady state (>= 2^24 samples) run k_trace_pool unless a flag says otherwise
    // (C5: 35.95 ms against 40.4 ms of k_trace, profiles/r2_a_kernel_ab.log); BVH scenes only on request (k_trace_bpool is
    // the slower kernel there).
    const bool pool_auto = r->mode == MODE_SPHERES && !(p->flags & (ZRT_FLAG_KERNEL_X2 | ZRT_FLAG_KERNEL_WARP)) &&
                           pixels * (uint64_t)plan->n_samples >= (1ull << 24);
    if (((p->flags & ZRT_FLAG_KERNEL_POOL) || pool_auto) && !(p->flags & ZRT_FLAG_KERNEL_THREAD) && pool_ok) {
        pool = 128;
        if (const char *e = std::getenv("ZRT_POOL_SLOTS")) {
            const uint32_t v = (uint32_t)std::atoi(e);
            pool = v >= 128u ? 128u : (v >= 96u ? 96u : 64u);
        }
    }
    uint32_t lanes = p->sample_chunks;
    if (lanes == 0) {
        lanes = 32u;
        if (r->mode == MODE_SPHERES) {
            static int sms = 0;
            if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device) != cudaSuccess) sms = 148;
            const double resident = pool == 128 ? 7.0 * 512.0 : (pool == 64 ? 8.0 * 256.0 : 8.0 * 128.0); // items per SM
            const double target = 54.0 * resident * (double)sms / (double)(pixels ? pixels : 1);
            lanes = 1u;
            while (lanes < 32u && (double)lanes * 1.41421356 < target) lanes <<= 1; // nearest power of two
        }
    }
    if (lanes > 32u) lanes = 32u;
    // k_trace_pool3 parks a finished item until 32 of them hand over together, which pays when items are long: below ~12
    // samples per item it takes 16 slices (125 spp per GPU of the 8-GPU split: 4.63 ms against 4.91 ms with 32; 250 spp:
    // 8.85 against 8.97; profiles/r2_f_pool3_handover_policy_ab.log)
    if (pool && p->sample_chunks == 0 && lanes == 32u && plan->n_samples < 384u) lanes = 16u;
    if (p->sample_chunks == 0) // the slice buffer is lanes x 12 B per pixel: keep the automatic choice under 2 GiB
        while (lanes > 1u && (uint64_t)p->width * p->height * 12ull * lanes > (2ull << 30)) lanes >>= 1;
    while (lanes & (lanes - 1u)) lanes &= lanes - 1u;          // round down to a power of two
    while (lanes > 1u && lanes > plan->n_samples) lanes >>= 1; // every slice gets at least one sample
    if (lanes < 1u) lanes = 1u;
    if (pixels * lanes > 0xFFF00000ull) return fail(ZRT_ERR_INVALID, "image too large"); // the item counter is 32 bits and every resident warp overshoots it once by one window
    P.lanes = lanes;
    P.lanes_log2 = 0;
    while ((1u << P.lanes_log2) < lanes) P.lanes_log2++;
    P.x_end_magic = P.x_end > 1 ? (uint32_t)((0x100000000ull + P.x_end - 1) / P.x_end) : 0xFFFFFFFFu; // see item_decode
    P.max_depth = p->max_depth;
    P.seed32 = (uint32_t)p->seed ^ (uint32_t)(p->seed >> 32);
    P.color_scale = (p->flags & ZRT_FLAG_RAW_SUM) ? 1.0f : 1.0f / (float)p->samples_per_pixel; // raytrace.zig:157
    P.count_pixels = (sb == 0) ? 1u : 0u;
    P.n_spheres = r->n_spheres; P.n_list = r->n_list; P.root = r->root;
    P.spheres = r->spheres.p;
    P.triA = r->triA.p; P.triE1 = r->triE1.p; P.triE2 = r->triE2.p; P.triMeta = r->triMeta.p;
    P.list = r->list.p; P.nodes = r->nodes.p; P.mats = sc->mats.p;
    if (r->mode == MODE_SPHERES)
        for (uint32_t i = 0; i < MAX_INLINE_SPHERES; i++) {
            KParams::SpherePair &pr = P.inl[i / 2];
            if (i < r->n_spheres) {
                const DevSphere &sp = r->host->spheres[i];
                pr.ncx[i & 1] = -sp.cx; pr.ncy[i & 1] = -sp.cy; pr.ncz[i & 1] = -sp.cz; pr.nr2[i & 1] = -sp.r2;
            } else {
                pr.ncx[i & 1] = pr.ncy[i & 1] = pr.ncz[i & 1] = 0.0f;
                pr.nr2[i & 1] = 1e30f; // c = |oc|^2 + 1e30 > half_b^2: the discriminant is always negative
            }
        }
    // kernel choice: the one-thread-per-path megakernel measured faster on every workload (C5: 43.2 ms against
    // 55.9 ms, DESIGN.md "Megakernel vs wavefront"), so block-sorted shading is opt-in.  It packs (px, py) into
    // 16 bits each and the material index into 18 bits.
#ifndef ZRT_EXPERIMENTS
    if (p->flags & (ZRT_FLAG_KERNEL_SORTED | ZRT_FLAG_KERNEL_X2))
        return fail(ZRT_ERR_INVALID, "ZRT_FLAG_KERNEL_SORTED / ZRT_FLAG_KERNEL_X2 are experiments: build libzrt with EXPERIMENTS=1");
#endif
    const bool sorted_ok = p->width < 65536u && p->height < 65536u && sc->host.materials.size() < (1u << 18);
    if ((p->flags & ZRT_FLAG_KERNEL_SORTED) && !sorted_ok)
        return fail(ZRT_ERR_INVALID, "ZRT_FLAG_KERNEL_SORTED needs width, height < 65536 and < 262144 materials");
    const bool sorted = (p->flags & ZRT_FLAG_KERNEL_SORTED) && !(p->flags & ZRT_FLAG_KERNEL_THREAD);
    P.sorted_shading = (sorted && sorted_ok) ? 1u : 0u;
    // BVH scenes: the warp-scheduled state machine k_trace_ws is opt-in as well: at 8 blocks per SM it runs C2 / C3 / C4
    // at 0.87 / 1.03 / 0.93 of k_trace's time, so neither kernel wins everywhere.  (Tried: letting the first render of a
    // scene time both on a 1/4 x 1/4 image at 32 spp and keep the faster one.  Launches of 0.2-0.5 ms are all ramp-up and
    // tail, where k_trace_ws is the slower one: it picked k_trace on all three.  A calibration long enough to be
    // representative costs more than the 3-13 % at stake for a scene that is rendered once.)
    // What is affordable is a rule: launches large enough to be mostly steady state (>= 2^24 samples) take k_trace_ws
    // unless a flag says otherwise, everything smaller k_trace.  On the three BVH configurations that is -13 %, +3 %, -7 %.
    const bool ws_ok = r->mode == MODE_BVH && !sorted && !(p->flags & (ZRT_FLAG_SAMPLER_HALTON | ZRT_FLAG_RUSSIAN_ROULETTE));
    const bool ws_auto = !(p->flags & (ZRT_FLAG_KERNEL_WARP | ZRT_FLAG_KERNEL_THREAD)) && pixels * plan->n_samples >= (1ull << 24);
    P.warp_scheduled = (ws_ok && ((p->flags & ZRT_FLAG_KERNEL_WARP) || ws_auto)) ? 1u : 0u;
    // thresholds from tools/ws_sweep.py at 8 resident blocks per SM (profiles/r1_v10_ws_sweep.log): (12, 2, 20) runs C2 /
    // C3 / C4 at 0.908 / 1.044 / 0.966 of the default kernel's time, the former (8, 4, 24) at 0.927 / 1.061 / 0.969
    P.ws_node_min = 12;  // keep stepping nodes while >= 12 lanes can
    P.ws_leaf_min = 2;   // run postponed leaves once 2 lanes hold one
    P.ws_shade_min = 20; // shade / regenerate once 20 lanes wait for it
    if (const char *e = std::getenv("ZRT_WS_THRESHOLDS")) { // "node,leaf,shade": tuning sweeps (tools/ws_sweep.sh)
        unsigned a = 0, b = 0, c = 0;
        if (std::sscanf(e, "%u,%u,%u", &a, &b, &c) == 3) { P.ws_node_min = a; P.ws_leaf_min = b; P.ws_shade_min = c; }
    }
    if (pool && r->mode == MODE_BVH) { // k_trace_bpool (K1p): node steps at >= 12 lanes, leaves at >= 2, refill at >= 8 idle, dry batch >= 16
        P.ws_node_min = 12; P.ws_leaf_min = 2; P.ws_shade_min = 8; P.ws_batch_min = 16; P.ws_burst_num = 3;
        if (const char *e = std::getenv("ZRT_POOL_THRESHOLDS")) { // "node,leaf,idle,batch[,burst]": tuning sweeps
            unsigned a = 0, b = 0, c = 0, d = 0, f = 3;
            if (std::sscanf(e, "%u,%u,%u,%u,%u", &a, &b, &c, &d, &f) >= 4 && c >= 1 && c <= 32 && f <= 4) {
                P.ws_node_min = a; P.ws_leaf_min = b; P.ws_shade_min = c; P.ws_batch_min = d; P.ws_burst_num = f;
            }
        }
        P.warp_scheduled = 0;
    }
    P.halton = (p->flags & ZRT_FLAG_SAMPLER_HALTON) ? 1u : 0u;
    P.roulette = (p->flags & ZRT_FLAG_RUSSIAN_ROULETTE) ? 1u : 0u;
    if ((P.halton || P.roulette) && (p->flags & (ZRT_FLAG_KERNEL_SORTED | ZRT_FLAG_KERNEL_WARP)))
        return fail(ZRT_ERR_INVALID, "the sampler extensions run on the thread kernel only");
    P.neg_zero[0] = P.neg_zero[1] = -0.0f;
    P.pk_ll[0] = P.llx; P.pk_ll[1] = P.lly; P.pk_h[0] = P.hx; P.pk_h[1] = P.hy; P.pk_v[0] = P.vx; P.pk_v[1] = P.vy;
    P.pk_no[0] = -P.ox; P.pk_no[1] = -P.oy; P.pk_nwh[0] = -P.f_width; P.pk_nwh[1] = -P.f_height;
    P.pk_rcp[0] = P.rcp_width; P.pk_rcp[1] = P.rcp_height;
    if (r->mode == MODE_SPHERES)
        for (uint32_t i = 0; i < r->n_spheres && i < MAX_INLINE_SPHERES; i++) {
            const DevSphere &sp = r->host->spheres[i];
            KParams::SphereX2 &e = P.inl2[i];
            e.ncx[0] = e.ncx[1] = -sp.cx; e.ncy[0] = e.ncy[1] = -sp.cy; e.ncz[0] = e.ncz[1] = -sp.cz;
            e.nr2[0] = e.nr2[1] = -sp.r2;
        }
    // k_trace_pool (K1q): spheres-only scenes, bounce count and pixel coordinates packed in 16 bits each
    P.inl_kinds = 0;
    P.row_order = 1; // the item queue runs from the top scanline down: 0.8-1.3 % faster than bottom up on C5 with either sphere
                     // kernel, 10.5 / 2.7 / 3.0 % on C2 / C3 / C4 (profiles/r2_k_tail_and_row_order_ab.log); results do not depend on it
    if (const char *e = std::getenv("ZRT_ROW_ORDER")) P.row_order = (uint32_t)std::atoi(e) & 3u; // A/B hook
    // k_trace_pool3 draws windows of 4 pixels' worth of items (128 at 32 slices, 64 at 16): the 128 slots of a warp then sit on
    // neighbouring pixels instead of on whatever 4 windows the ~4000 resident warps left it, and its batches of primary rays
    // stay coherent.  Measured on C5 (profiles/r2_x_pool3_queue_window_fine_ab.log): 32 / 64 / 96 / 128 / 160 items -> 32.88 /
    // 32.45 / 32.18 / 32.12 / 32.02 ms at 1000 spp, smooth; beyond ~200 items (more than a pool holds) it turns erratic and
    // slower (256: 33.7, 384: 35).  The last round of windows before the queue ends is 32 items again.
    P.queue_window = 32;
    P.queue_taper = 0;
    if (r->mode == MODE_BVH && !pool) { // k_trace_ws / k_trace on BVH scenes: two pixels' worth per atomic.  C2 / C3 / C4: 7.97 / 32.85 /
        // 86.4 ms at 32 items, 7.81 / 32.19 / 85.0 at 64; 128 and more make the end of the launch depend on which warp drew the
        // heaviest pixels (C4: 87-95 ms at 128, 103 at 256; profiles/r2_aa_bvh_queue_window_ab.log)
        uint32_t win = 64;
        if (const char *e = std::getenv("ZRT_QUEUE_WINDOW_ALL")) win = (uint32_t)std::atoi(e) & ~31u; // A/B hook
        const uint64_t items = pixels * lanes, warps = 148ull * 32ull;
        if (win > 32u && lanes == 32u && items >= warps * win * 24ull) { P.queue_window = win; P.queue_taper = (uint32_t)(items - warps * win * 2ull); }
    }
    if (pool && r->mode == MODE_SPHERES) {
        uint32_t win = 4u * lanes;
        win = win < 64u ? 64u : (win > 128u ? 128u : win); // 8 slices at 2000^2: 64 items 32.4 ms, 32 items 32.9 ms
        const uint64_t items = pixels * lanes, warps = 148ull * 28ull;
        bool on = items >= warps * win * 48ull; // a warp should see ~50 windows or the coarser hand-out costs balance (500^2: 2.8 against 2.4 ms)
        if (const char *e = std::getenv("ZRT_QUEUE_WINDOW")) { win = (uint32_t)std::atoi(e) & ~31u; on = win > 32u && items > warps * win * 2ull; } // A/B hook
        if (on) { P.queue_window = win; P.queue_taper = (uint32_t)(items - warps * win * 2ull); }
    }
    P.pool_split = 0; // C5: 34.34 ms without the image rings, 34.60 ms with them (profiles/r2_c_pool3_ab.log)
    if (const char *e = std::getenv("ZRT_POOL_SPLIT")) P.pool_split = std::atoi(e) ? 1u : 0u; // A/B hook
    if (r->mode == MODE_SPHERES)
        for (uint32_t i = 0; i < r->n_spheres && i < MAX_INLINE_SPHERES; i++) { // PoolKind: 1 + kind, image variants at 4, 5
            const uint32_t mat = r->host->spheres[i].material, kind = (mat >> MAT_KIND_SHIFT) & 3u;
            uint32_t ring = 1u + kind;
            if (P.pool_split && (mat & MAT_IMAGE_BIT) && kind != ZRT_MATERIAL_DIELECTRIC) ring = 4u + kind;
            P.inl_kinds |= ring << (3u * i);
        }
    if (r->mode == MODE_SPHERES)
        for (uint32_t i = 0; i < MAX_INLINE_SPHERES; i++) { // the operations of closest_spheres_inline, once, on the host
            const KParams::SpherePair &s = P.inl[i / 2];
            const float ocx = P.ox + s.ncx[i & 1], ocy = P.oy + s.ncy[i & 1], ocz = P.oz + s.ncz[i & 1];
            const float c = ((ocx * ocx + ocy * ocy) + ocz * ocz) + s.nr2[i & 1];
            KParams::SpherePair &q = P.inl_prim[i / 2];
            q.ncx[i & 1] = ocx; q.ncy[i & 1] = ocy; q.ncz[i & 1] = ocz; q.nr2[i & 1] = -c;
        }
    P.pool = pool;
    P.two_paths = (r->mode == MODE_SPHERES && (p->flags & ZRT_FLAG_KERNEL_X2) && !P.sorted_shading && !P.halton && !P.roulette && !P.pool) ? 1u : 0u;
    plan->mode = r->mode;
    plan->n_floats = (size_t)p->width * p->height * 3;
    return ZRT_OK;
}

// the host waits until the last render enqueued on a caller's stream has finished (before scene buffers are freed or
// reallocated: cudaFreeAsync on the allocation stream is not ordered after the caller's stream)
void quiesce(zrt_scene *sc) {
    if (sc->user_pending && sc->ev_user) cudaEventSynchronize(sc->ev_user);
    sc->user_pending = false;
}
// `st` waits for the last render on a caller's stream: two renders of one scene share its scratch buffers
cudaError_t orderAfterUser(zrt_scene *sc, cudaStream_t st) {
    if (!sc->user_pending) return cudaSuccess;
    return cudaStreamWaitEvent(st, sc->ev_user, 0);
}

// enqueue everything for one render on `st`; d_rgb receives the final image
int enqueueRender(zrt_scene *sc, Plan &plan, float *d_rgb, unsigned long long *d_counters, cudaStream_t st,
                  cudaEvent_t e_k0, cudaEvent_t e_k1, cudaEvent_t e_r1, uint32_t *launches, uint8_t *d_rgb8) {
    KParams &P = plan.P;
    if ((P.lanes > 1 && plan.n_floats * P.lanes > sc->part.n) || sc->work.n < 1) quiesce(sc); // about to reallocate scratch
    CUDA_TRY(orderAfterUser(sc, st));
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, 6 * sizeof(unsigned long long), st));
    float *trace_out = d_rgb;
    if (P.lanes > 1) {
        CUDA_TRY(sc->part.reserve(plan.n_floats * P.lanes));
        trace_out = sc->part.p;
    }
    // pixels the reference never writes (x >= height, Q1) stay black: image.zig:80-90
    if (P.x_end < P.width) CUDA_TRY(cudaMemsetAsync(trace_out, 0, plan.n_floats * P.lanes * sizeof(float), st));
    P.out = trace_out;
    P.counters = d_counters;
    CUDA_TRY(sc->work.reserve(1));
    CUDA_TRY(cudaMemsetAsync(sc->work.p, 0, sizeof(uint32_t), st));
    P.work_counter = sc->work.p;
    *launches = 0;
    if (e_k0) CUDA_TRY(cudaEventRecord(e_k0, st));
    bool traced = false;
    if (plan.n_samples > 0 && P.max_depth > 0) {
        *launches += launch_trace(P, plan.mode, st); // k_trace (+ k_finish_counters)
        traced = true;
    } else {
        // no samples, or max_depth = 0: every sample ends at the recursion limit without casting a ray
        // (raytrace.zig:64-68); the image is black and the counters are known on the host
        CUDA_TRY(cudaMemsetAsync(d_rgb, 0, plan.n_floats * sizeof(float), st));
        const unsigned long long n_px = (unsigned long long)P.x_end * P.height, n_s = n_px * plan.n_samples;
        const unsigned long long c[6] = {n_s, 0, 0, P.count_pixels ? n_px : 0ull, n_s, 0};
        CUDA_TRY(cudaMemcpyAsync(d_counters, c, sizeof(c), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st)); // `c` lives on this stack frame
    }
    if (e_k1) CUDA_TRY(cudaEventRecord(e_k1, st));
    if (d_rgb8) { // fused output stage: chunk sum + 1/spp + 8-bit quantisation + row flip in one pass
        if (traced) launch_resolve_rgb8(trace_out, d_rgb8, P.width, P.height, P.lanes, P.lanes > 1 ? P.color_scale : 1.0f, st);
        else launch_resolve_rgb8(d_rgb, d_rgb8, P.width, P.height, 1, 1.0f, st);
        (*launches)++;
    } else if (traced && P.lanes > 1) {
        launch_resolve(sc->part.p, d_rgb, (uint32_t)plan.n_floats, P.lanes, P.color_scale, st);
        (*launches)++;
    }
    if (e_r1) CUDA_TRY(cudaEventRecord(e_r1, st));
    CUDA_TRY(cudaGetLastError());
    sc->launch_count += *launches;
    return ZRT_OK;
}

int requireDevice(zrt_scene *sc) {
    if (!sc) return fail(ZRT_ERR_INVALID, "scene is NULL");
    if (sc->device < 0) return fail(ZRT_ERR_NO_DEVICE, "scene was created without a device; libzrt has no CPU path");
    CUDA_TRY(cudaSetDevice(sc->device));
    return ZRT_OK;
}

} // namespace zrt

extern "C" {

int zrt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *zrt_last_error(void) { return g_error.c_str(); }

int zrt_scene_create(const zrt_scene_desc *desc, int device, zrt_scene **out) {
    if (!out) return fail(ZRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate(desc);
    if (rc != ZRT_OK) return rc;
    if (device >= 0) {
        const int n = zrt_device_count();
        if (n == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible; libzrt has no CPU path");
        if (device >= n) return fail(ZRT_ERR_INVALID, "device index out of range");
    }
    zrt_scene *sc = new (std::nothrow) zrt_scene();
    if (!sc) return fail(ZRT_ERR_OOM, "out of memory");
    try {
        HostScene &hs = sc->host;
        hs.surfaces.assign(desc->surfaces, desc->surfaces + desc->n_surfaces);
        hs.spheres.assign(desc->spheres, desc->spheres + desc->n_spheres);
        hs.triangles.assign(desc->triangles, desc->triangles + desc->n_triangles);
        hs.materials.assign(desc->materials, desc->materials + desc->n_materials);
        hs.textures.assign(desc->textures, desc->textures + desc->n_textures);
        hs.texels.resize(desc->n_textures);
        for (uint32_t i = 0; i < desc->n_textures && device < 0; i++) { // host-only scenes own a copy of the texels
            zrt_texture &t = hs.textures[i];
            if (t.kind != ZRT_TEXTURE_IMAGE) continue;
            const size_t bytes = (size_t)t.width * t.height * t.channels;
            hs.texels[i].assign(t.pixels, t.pixels + bytes);
            t.pixels = hs.texels[i].data();
        }
    } catch (const std::bad_alloc &) {
        delete sc;
        return fail(ZRT_ERR_OOM, "out of memory");
    }
    sc->all_spheres = true;
    for (const auto &s : sc->host.surfaces) sc->all_spheres &= (s.kind == ZRT_SURFACE_SPHERE);
    sc->device = device;
    if (device >= 0) {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking);
        for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&sc->ev[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sc->ev_user, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            const std::string msg = cudaGetErrorString(e);
            zrt_scene_destroy(sc);
            return fail(ZRT_ERR_CUDA, "device setup failed: " + msg);
        }
        rc = uploadMaterials(sc);
        for (auto &t : sc->host.textures) t.pixels = nullptr; // borrowed from the caller for the upload only
        if (rc != ZRT_OK) {
            zrt_scene_destroy(sc);
            return rc;
        }
    }
    *out = sc;
    return ZRT_OK;
}

int zrt_pinned_alloc(size_t bytes, void **out) {
    if (!out) return fail(ZRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (bytes == 0) return fail(ZRT_ERR_INVALID, "zero-sized allocation");
    if (zrt_device_count() == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible: page-locking needs the driver");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return ZRT_OK;
}

void zrt_pinned_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

void zrt_scene_destroy(zrt_scene *sc) {
    if (!sc) return;
    if (sc->device >= 0) {
        cudaSetDevice(sc->device);
        quiesce(sc); // renders enqueued on a caller's stream may still be reading the buffers freed below
        if (sc->stream) cudaStreamSynchronize(sc->stream);
        sc->rep_list.release(); sc->rep_bvh.release(); sc->rep_sah.release();
        sc->mats.release(); sc->part.release(); sc->image.release(); sc->image8.release(); sc->counters.release();
        sc->hit_id.release(); sc->hit_t.release(); sc->work.release();
        for (auto &t : sc->d_texels) t.release();
        for (auto &e : sc->ev)
            if (e) cudaEventDestroy(e);
        if (sc->ev_user) cudaEventDestroy(sc->ev_user);
        if (sc->stream) cudaStreamDestroy(sc->stream);
    }
    delete sc;
}

// raytrace.render() with host buffers out: the float image (out_rgb) or the 8-bit image of the fused output stage (out_rgb8)
static int renderToHost(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, float *out_rgb, uint8_t *out_rgb8,
                        zrt_counters *counters, zrt_timing *timing) {
    int rc = requireDevice(sc);
    if (rc != ZRT_OK) return rc;
    if (!out_rgb && !out_rgb8) return fail(ZRT_ERR_INVALID, out_rgb8 ? "out_rgb8 is NULL" : "out_rgb is NULL");
    if (!params) return fail(ZRT_ERR_INVALID, "params is NULL");
    if (out_rgb8 && (params->flags & ZRT_FLAG_RAW_SUM)) return fail(ZRT_ERR_INVALID, "ZRT_FLAG_RAW_SUM has no 8-bit form");
    DevRep *rep = nullptr;
    const auto t_prep0 = std::chrono::steady_clock::now();
    rc = selectRep(sc, params, &rep);
    if (rc != ZRT_OK) return rc;
    const float prep_now = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_prep0).count();
    Plan plan;
    rc = makePlan(sc, camera, params, rep, &plan);
    if (rc != ZRT_OK) return rc;
    if (plan.n_floats > sc->image.n || (out_rgb8 && plan.n_floats > sc->image8.n)) quiesce(sc);
    CUDA_TRY(sc->image.reserve(plan.n_floats));
    if (out_rgb8) CUDA_TRY(sc->image8.reserve(plan.n_floats));
    CUDA_TRY(sc->counters.reserve(10));
    uint32_t launches = 0;
    rc = enqueueRender(sc, plan, sc->image.p, sc->counters.p, sc->stream, sc->ev[0], sc->ev[1], sc->ev[2], &launches,
                       out_rgb8 ? sc->image8.p : nullptr);
    if (rc != ZRT_OK) return rc;
    if (out_rgb8) CUDA_TRY(cudaMemcpyAsync(out_rgb8, sc->image8.p, plan.n_floats, cudaMemcpyDeviceToHost, sc->stream));
    else CUDA_TRY(cudaMemcpyAsync(out_rgb, sc->image.p, plan.n_floats * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
    unsigned long long h_counters[6];
    CUDA_TRY(cudaMemcpyAsync(h_counters, sc->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, sc->stream));
    CUDA_TRY(cudaEventRecord(sc->ev[3], sc->stream));
    CUDA_TRY(cudaStreamSynchronize(sc->stream));
    if (counters) {
        counters->recursion_depth_hits = h_counters[0];
        counters->reflections = h_counters[1];
        counters->background_hits = h_counters[2];
        counters->pixels_processed = h_counters[3];
        counters->samples_processed = h_counters[4];
        counters->rays_processed = h_counters[5];
    }
    if (timing) {
        float k = 0, r = 0, tot = 0;
        cudaEventElapsedTime(&k, sc->ev[0], sc->ev[1]);
        cudaEventElapsedTime(&r, sc->ev[1], sc->ev[2]);
        cudaEventElapsedTime(&tot, sc->ev[0], sc->ev[3]);
        timing->prepare_ms = prep_now;
        timing->kernel_ms = k;
        timing->resolve_ms = r;
        timing->total_ms = tot;
        timing->launches = launches;
        timing->bvh_nodes = (uint32_t)rep->host->info.nodes.size();
    }
    return ZRT_OK;
}

int zrt_render(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, float *out_rgb,
               zrt_counters *counters, zrt_timing *timing) {
    if (!out_rgb) return fail(ZRT_ERR_INVALID, "out_rgb is NULL");
    return renderToHost(sc, camera, params, out_rgb, nullptr, counters, timing);
}

int zrt_render_rgb8(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, uint8_t *out_rgb8,
                    zrt_counters *counters, zrt_timing *timing) {
    if (!out_rgb8) return fail(ZRT_ERR_INVALID, "out_rgb8 is NULL");
    return renderToHost(sc, camera, params, nullptr, out_rgb8, counters, timing);
}

int zrt_render_device(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, float *d_rgb,
                      uint64_t *d_counters, void *stream) {
    int rc = requireDevice(sc);
    if (rc != ZRT_OK) return rc;
    if (!d_rgb || !d_counters) return fail(ZRT_ERR_INVALID, "device output pointers are NULL");
    if (!params) return fail(ZRT_ERR_INVALID, "params is NULL");
    DevRep *rep = nullptr;
    rc = selectRep(sc, params, &rep);
    if (rc != ZRT_OK) return rc;
    Plan plan;
    rc = makePlan(sc, camera, params, rep, &plan);
    if (rc != ZRT_OK) return rc;
    uint32_t launches = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = enqueueRender(sc, plan, d_rgb, reinterpret_cast<unsigned long long *>(d_counters), st, nullptr, nullptr, nullptr,
                       &launches);
    if (rc != ZRT_OK) return rc;
    CUDA_TRY(cudaEventRecord(sc->ev_user, st));
    sc->user_pending = true;
    return ZRT_OK;
}

int zrt_primary_hits(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, int jitter,
                     uint32_t *surface_id, float *t) {
    int rc = requireDevice(sc);
    if (rc != ZRT_OK) return rc;
    quiesce(sc);
    if (!surface_id || !t) return fail(ZRT_ERR_INVALID, "output pointers are NULL");
    if (!params) return fail(ZRT_ERR_INVALID, "params is NULL");
    DevRep *rep = nullptr;
    rc = selectRep(sc, params, &rep);
    if (rc != ZRT_OK) return rc;
    Plan plan;
    zrt_params p2 = *params;
    if (p2.samples_per_pixel == 0) p2.samples_per_pixel = 1;
    if (p2.sample_begin == 0 && p2.sample_end == 0) p2.sample_end = p2.samples_per_pixel;
    rc = makePlan(sc, camera, &p2, rep, &plan);
    if (rc != ZRT_OK) return rc;
    const size_t n = (size_t)params->width * params->height;
    CUDA_TRY(sc->hit_id.reserve(n));
    CUDA_TRY(sc->hit_t.reserve(n));
    plan.P.hit_id = sc->hit_id.p;
    plan.P.hit_t = sc->hit_t.p;
    plan.P.jitter = jitter ? 1u : 0u;
    launch_primary(plan.P, plan.mode, sc->stream);
    CUDA_TRY(cudaGetLastError());
    sc->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(surface_id, sc->hit_id.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc->stream));
    CUDA_TRY(cudaMemcpyAsync(t, sc->hit_t.p, n * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
    CUDA_TRY(cudaStreamSynchronize(sc->stream));
    return ZRT_OK;
}

static const FlatBvh *hostBvh(zrt_scene *sc, uint32_t flags) {
    const int k = (flags & ZRT_FLAG_BVH_REFERENCE) ? 0 : 1;
    DevRep &r = k ? sc->rep_sah : sc->rep_bvh;
    if (r.host) return &r.host->info;
    if (!sc->host_bvh_ready[k]) {
        try {
            build_flat_bvh(sc->host, k != 0, &sc->host_bvh[k]);
        } catch (const std::exception &e) {
            fail(ZRT_ERR_OOM, std::string("scene flattening failed: ") + e.what());
            return nullptr;
        }
        sc->host_bvh_ready[k] = true;
    }
    return &sc->host_bvh[k];
}

uint64_t zrt_scene_launch_count(const zrt_scene *sc) { return sc ? sc->launch_count : 0; }

uint32_t zrt_build_features(void) {
#ifdef ZRT_EXPERIMENTS
    return ZRT_FEATURE_EXPERIMENTS;
#else
    return 0;
#endif
}

int zrt_scene_bvh_info(zrt_scene *sc, uint32_t flags, zrt_bvh_info *out) {
    if (!sc || !out) return fail(ZRT_ERR_INVALID, "NULL argument");
    const FlatBvh *b = hostBvh(sc, flags);
    if (!b) return ZRT_ERR_OOM;
    out->nodes = (uint32_t)b->nodes.size();
    out->leaves = b->leaves;
    out->max_depth = b->max_depth;
    out->pruned_surfaces = b->pruned;
    out->reference_nodes = b->ref_nodes;
    out->reference_max_depth = b->ref_max_depth;
    return ZRT_OK;
}

int zrt_scene_bvh_order(zrt_scene *sc, uint32_t *order, uint8_t *visible) {
    if (!sc || !order || !visible) return fail(ZRT_ERR_INVALID, "NULL argument");
    const FlatBvh *b = hostBvh(sc, 0);
    if (!b) return ZRT_ERR_OOM;
    for (size_t s = 0; s < b->slot_surface.size(); s++) order[s] = b->slot_surface[s];
    std::memset(visible, 0, sc->host.surfaces.size());
    for (size_t s = 0; s < b->slot_surface.size(); s++) visible[b->slot_surface[s]] = b->slot_visible[s];
    return ZRT_OK;
}

int zrt_trace_statistics(zrt_scene *sc, const zrt_camera *camera, const zrt_params *params, zrt_trace_stats *out) {
    int rc = requireDevice(sc);
    if (rc != ZRT_OK) return rc;
    quiesce(sc);
    if (!out || !params) return fail(ZRT_ERR_INVALID, "NULL argument");
    DevRep *rep = nullptr;
    rc = selectRep(sc, params, &rep);
    if (rc != ZRT_OK) return rc;
    Plan plan;
    rc = makePlan(sc, camera, params, rep, &plan);
    if (rc != ZRT_OK) return rc;
    CUDA_TRY(sc->image.reserve(plan.n_floats));
    CUDA_TRY(sc->counters.reserve(6 + 4));
    CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 10 * sizeof(unsigned long long), sc->stream));
    plan.P.stats = sc->counters.p + 6;
    uint32_t launches = 0;
    rc = enqueueRender(sc, plan, sc->image.p, sc->counters.p, sc->stream, nullptr, nullptr, nullptr, &launches);
    if (rc != ZRT_OK) return rc;
    unsigned long long h[10];
    CUDA_TRY(cudaMemcpyAsync(h, sc->counters.p, sizeof(h), cudaMemcpyDeviceToHost, sc->stream));
    CUDA_TRY(cudaStreamSynchronize(sc->stream));
    out->rays = h[5];
    out->samples = h[4];
    out->node_visits = h[6];
    out->triangle_tests = h[7];
    out->sphere_tests = h[8];
    out->texture_lookups = h[9];
    return ZRT_OK;
}

int zrt_selftest(int device, uint64_t *mismatches) {
    if (!mismatches) return fail(ZRT_ERR_INVALID, "mismatches is NULL");
    if (zrt_device_count() == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible");
    CUDA_TRY(cudaSetDevice(device));
    unsigned long long *d = nullptr, h = 0;
    CUDA_TRY(cudaMalloc(&d, sizeof(h)));
    CUDA_TRY(cudaMemset(d, 0, sizeof(h)));
    const uint32_t widths[] = {1, 7, 200, 512, 1000, 1024, 1080, 1920, 4096, 65535};
    uint32_t seed = 1;
    for (uint32_t w : widths) launch_selftest_div(d, w, seed++, 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = h;
    return ZRT_OK;
}

int zrt_measure_peaks(int device, double *out, int n) {
    if (!out || n < 5) return fail(ZRT_ERR_INVALID, "need room for 5 results");
    if (zrt_device_count() == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float *d_out = nullptr;
    const int blocks = sms * 8, threads = 256, iters = 8192;
    CUDA_TRY(cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(float)));
    auto best_of = [&](auto &&launch, int reps, float *best_ms) -> cudaError_t {
        *best_ms = 1e30f;
        for (int i = 0; i < reps + 2; i++) {
            cudaEventRecord(e0, 0);
            launch();
            cudaEventRecord(e1, 0);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) return e;
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2 && ms < *best_ms) *best_ms = ms;
        }
        return cudaGetLastError();
    };
    float ms;
    CUDA_TRY(best_of([&] { launch_peak_fp32(d_out, blocks, threads, iters, 0); }, 5, &ms));
    out[0] = (double)blocks * threads * iters * 16.0 / (ms * 1e-3);
    CUDA_TRY(best_of([&] { launch_peak_ffma(d_out, blocks, threads, iters, 0); }, 5, &ms));
    out[1] = (double)blocks * threads * iters * 16.0 / (ms * 1e-3);
    // L2-resident read: 32 MiB buffer read 16 times per launch
    float4 *buf = nullptr;
    const size_t l2_bytes = 32ull << 20, hbm_bytes = 1ull << 30;
    CUDA_TRY(cudaMalloc(&buf, hbm_bytes));
    CUDA_TRY(cudaMemset(buf, 0, hbm_bytes));
    CUDA_TRY(best_of([&] { launch_peak_read(buf, l2_bytes / 16, 16, d_out, sms * 8, 512, 0); }, 5, &ms));
    out[2] = (double)l2_bytes * 16 / (ms * 1e-3) / 1e9;
    CUDA_TRY(best_of([&] { launch_peak_read(buf, hbm_bytes / 16, 1, d_out, sms * 8, 512, 0); }, 5, &ms));
    out[3] = (double)hbm_bytes / (ms * 1e-3) / 1e9;
    out[4] = prop.clockRate / 1000.0;
    cudaFree(buf);
    cudaFree(d_out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ZRT_OK;
}

} // extern "C"
