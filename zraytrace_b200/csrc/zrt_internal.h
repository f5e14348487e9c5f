// zrt_internal.h — flattened scene layout shared by the host flattener and the sm_100a kernels.
// Everything here is resident in HBM (in practice L1/L2: a scene is at most a few MB).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <memory>
#include <string>
#include <system_error>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/zrt.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
#endif

namespace zrt {

// ---- primitive references (BVH child refs and list entries) -------------------------------------
// ref >= 0            : inner node index
// ref & 0x80000000    : leaf; bit 30 = sphere, low 30 bits = triangle slot / sphere index
constexpr uint32_t REF_LEAF = 0x80000000u;
constexpr uint32_t REF_SPHERE = 0x40000000u;
constexpr uint32_t REF_INDEX_MASK = 0x3FFFFFFFu;
constexpr uint32_t REF_EMPTY = 0x7FFFFFFFu; // child slot removed by flat-box pruning

// Material word carried by every primitive: the kind and "samples an image" bit ride along with the
// index so that shading can branch without first fetching the material record.
constexpr uint32_t MAT_INDEX_MASK = 0x00FFFFFFu;
constexpr uint32_t MAT_KIND_SHIFT = 24;           // 2 bits: ZRT_MATERIAL_*
constexpr uint32_t MAT_IMAGE_BIT = 1u << 26;      // lambertian/metal whose texture is an image

// sphere.zig:15-20, 32 B = two 128-bit loads.  r2 and inv_r are the f32 values the reference
// recomputes on every test (`radius * radius` sphere.zig:34, `1.0/radius` sphere.zig:46): computing
// them once on the host with the same IEEE operation is bit-identical.
struct alignas(16) DevSphere {
    float cx, cy, cz, r2;
    float inv_r;
    uint32_t material;   // packed material word
    uint32_t surface_id; // index in the caller's surface list
    uint32_t slot;       // tie-break key: left-first DFS position in the reference tree / list position
};
static_assert(sizeof(DevSphere) == 32, "DevSphere layout");

// Triangles: structure of three float4 planes (16-byte SoA), indexed by slot:
//   triA [i] = (a.x,  a.y,  a.z,  n.x)     triangle.zig:15-30 fields a, e1, e2, face_normal
//   triE1[i] = (e1.x, e1.y, e1.z, n.y)     (face_unit_normal is recomputed for the one final hit)
//   triE2[i] = (e2.x, e2.y, e2.z, n.z)
// plus triMeta[i] = (packed material word, surface_id).  48 B read per test as three LDG.128.
struct alignas(8) TriMeta {
    uint32_t material, surface_id;
};

// Materials with their texture folded in (material.zig:16-29 + texture.zig:7-16), 64 B.
struct alignas(16) DevMaterial {
    uint32_t kind;     // ZRT_MATERIAL_*
    uint32_t tex_kind; // ZRT_TEXTURE_*
    float ior;         // index_of_refraction
    float inv_ior;     // 1.0f / ior (material.zig:111)
    float r, g, b;     // ColorTexture.color
    float u_off;
    float v_off;
    uint32_t w, h, ch;
    const uint8_t *pixels; // device pointer, rows bottom-up
    float r0_front, r0_back; // dielectric: (1-ratio)/(1+ratio) for ratio = 1/ior and ior (material.zig:126)
};
static_assert(sizeof(DevMaterial) == 64, "DevMaterial layout");

// 64-byte BVH node, 64-byte aligned, read as four 128-bit loads.  Left and right child boxes are interleaved
// so that every float2 is a (left, right) pair ready for the packed f32x2 slab test:
//   q0 = (lmin.x, rmin.x, lmin.y, rmin.y)
//   q1 = (lmin.z, rmin.z, lmax.x, rmax.x)
//   q2 = (lmax.y, rmax.y, lmax.z, rmax.z)
//   q3 = (left_ref, right_ref, -, -)
struct alignas(64) DevNode {
    float mn_x[2], mn_y[2], mn_z[2], mx_x[2], mx_y[2], mx_z[2]; // [0] = left child, [1] = right child
    uint32_t left, right, pad0, pad1;
    void set_box(int child, const float mn[3], const float mx[3]) {
        mn_x[child] = mn[0]; mn_y[child] = mn[1]; mn_z[child] = mn[2];
        mx_x[child] = mx[0]; mx_y[child] = mx[1]; mx_z[child] = mx[2];
    }
};
static_assert(sizeof(DevNode) == 64, "DevNode layout");

constexpr int MAX_INLINE_SPHERES = 8;
constexpr int TRAVERSAL_STACK = 64;

// Everything a kernel needs, passed by value as a __grid_constant__ parameter (constant bank).
struct KParams {
    // camera.zig:11-15
    float ox, oy, oz, llx, lly, llz, hx, hy, hz, vx, vy, vz;
    float f_width, f_height;
    float rcp_width, rcp_height; // RN(1/width), RN(1/height) for the exact quotients of raytrace.zig:173-174
    uint32_t width, height, x_end;
    uint32_t s_begin, s_end; // global sample range of this launch
    uint32_t lanes;   // L: slices per pixel (power of two <= 32); a work item is (pixel, slice)
    uint32_t lanes_log2;
    uint32_t x_end_magic; // ceil(2^32 / x_end) (0 if x_end == 1): q / x_end ~ umulhi(q, magic), corrected by one
    uint32_t *work_counter; // global item queue head, zeroed before the launch
    uint32_t max_depth, seed32;
    float color_scale; // 1/spp, or 1 for ZRT_FLAG_RAW_SUM
    uint32_t count_pixels; // 1 if this launch owns sample 0 (pixels_processed is counted once)
    uint32_t jitter;       // primary-hit kernel only
    uint32_t sorted_shading; // 1: k_trace_sorted (block-sorted shading), 0: k_trace (one thread per path)
    uint32_t halton, roulette; // sampler extensions (ZRT_FLAG_SAMPLER_HALTON, ZRT_FLAG_RUSSIAN_ROULETTE): k_trace<EXT>
    uint32_t warp_scheduled; // BVH scenes: 1: k_trace_ws (warp-scheduled node / leaf / shade sections)
    uint32_t ws_node_min, ws_leaf_min, ws_shade_min; // k_trace_ws: lanes that must wait for a section before the warp runs it
    uint32_t ws_burst_num; // k_trace_bpool: a node-step burst ends when fewer than ws_burst_num / 4 of its lanes are left at inner nodes
    uint32_t ws_batch_min; // k_trace_bpool: smallest shade batch worth running when the TRAV ring is dry (ws_shade_min = lanes idle before a refill)
    // scene
    uint32_t n_spheres, n_list;
    uint32_t root; // BVH root ref
    const DevSphere *spheres;
    const float4 *triA, *triE1, *triE2;
    const TriMeta *triMeta;
    const uint32_t *list; // list mode: refs in caller order
    const DevNode *nodes;
    const DevMaterial *mats;
    // outputs
    float *out;                   // [chunks][height][width][3]
    unsigned long long *counters; // 6 x u64, zrt_counters order
    unsigned long long *stats;    // non-NULL selects the instrumented kernel: 4 x u64 event counts
    uint32_t *hit_id;             // primary-hit kernel
    float *hit_t;
    // spheres-only scenes: operands straight from the constant bank, two spheres per packed f32x2 operand.
    // pair p holds spheres 2p and 2p+1: negated centre coordinates and negated r^2 (a - b == a + (-b) exactly);
    // an odd count is padded with a sphere that can never be hit (r^2 = -1e30)
    struct SpherePair {
        float ncx[2], ncy[2], ncz[2], nr2[2];
    } inl[MAX_INLINE_SPHERES / 2];
    // k_trace_x2 (two paths per thread): one sphere per entry, every value twice, so that an operand pair comes
    // straight from the constant bank; neg_zero is the -0.0 pair no compiler pass can see through (see K1x2)
    struct SphereX2 {
        float ncx[2], ncy[2], ncz[2], nr2[2];
    } inl2[MAX_INLINE_SPHERES];
    float neg_zero[2];
    // packed operands of primary_direction_raw: (x, y) pairs of the camera vectors, (width, height) and reciprocals
    float pk_ll[2], pk_h[2], pk_v[2], pk_no[2], pk_nwh[2], pk_rcp[2];
    uint32_t two_paths; // 1: k_trace_x2
    uint32_t pool;      // spheres-only scenes: slots per warp of k_trace_pool (0 = k_trace)
    uint32_t inl_kinds; // k_trace_pool: shade ring (PoolKind) of inline sphere i in bits 3i .. 3i+2
    uint32_t queue_window, queue_taper; // items a warp draws from the global queue per atomic (32, or more while the counter is below queue_taper)
    uint32_t row_order; // scanline order of the item queue: 0 bottom up, 1 top down, 2 middle outwards, 3 edges inwards
    uint32_t pool_split; // 1: image-textured Lambertian / metal spheres have rings of their own (PK_LAMB_IMG, PK_METAL_IMG)
    SpherePair inl_prim[MAX_INLINE_SPHERES / 2]; // rays from the camera origin: (oc.x, oc.y, oc.z, -(|oc|^2 - r^2)) per sphere
};

enum TraceMode { MODE_SPHERES = 0, MODE_LIST = 1, MODE_BVH = 2 };

// A helper thread that cannot fail to start (if the system refuses another thread, the work runs on the caller's) and
// whose exceptions come back to the joining thread: nothing may reach std::terminate below the C ABI.
struct Worker {
    std::thread t;
    std::shared_ptr<std::exception_ptr> err;
    Worker() = default;
    Worker(Worker &&) = default;
    Worker &operator=(Worker &&o) {
        if (this != &o) {
            finish();
            t = std::move(o.t);
            err = std::move(o.err);
        }
        return *this;
    }
    template <class F>
    explicit Worker(F f) : err(std::make_shared<std::exception_ptr>()) {
        auto slot = err;
        auto body = [f, slot]() mutable {
            try {
                f();
            } catch (...) {
                *slot = std::current_exception();
            }
        };
        try {
            t = std::thread(body);
        } catch (const std::system_error &) {
            body();
        }
    }
    ~Worker() { finish(); } // a caller that unwinds past a running worker waits for it instead of terminating
    void finish() noexcept {
        if (t.joinable()) t.join();
    }
    void join() { // rethrows what the worker threw
        finish();
        if (err && *err) {
            std::exception_ptr e = *err;
            *err = nullptr;
            std::rethrow_exception(e);
        }
    }
};

// fn(begin, end) over [0, n) in contiguous chunks on up to 16 host threads
template <class F>
void parallelFor(size_t n, size_t min_chunk, F fn) {
    if (n < 2 * min_chunk) { fn((size_t)0, n); return; }
    static const size_t hw = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    const size_t threads = std::min(hw, n / min_chunk);
    if (threads <= 1) { fn((size_t)0, n); return; }
    std::vector<Worker> th(threads - 1);
    const size_t chunk = (n + threads - 1) / threads;
    for (size_t t = 1; t < threads; t++) th[t - 1] = Worker([=] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
    fn((size_t)0, std::min(n, chunk));
    for (auto &t : th) t.join();
}

// ZRT_TIMING=1 in the environment prints the host-side phases of a BVH scene build to stderr
struct BuildLap {
    bool on = std::getenv("ZRT_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void operator()(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[zrt build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---- host-side flattened scene (zrt_flatten.cpp) ------------------------------------------------
// std::allocator whose value-less construct() default-initialises: resize() of a vector of plain structs then leaves the
// new elements uninitialised instead of zeroing (and page-faulting) tens of MB on one thread
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    template <class U> void construct(U *p) { ::new ((void *)p) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};

struct FlatBvh {
    std::vector<DevNode, DefaultInitAlloc<DevNode>> nodes;
    uint32_t root = REF_EMPTY;
    std::vector<uint32_t> slot_surface;  // slot -> surface id (DFS order of the reference tree)
    std::vector<uint8_t> slot_visible;   // 0 if pruned (under a zero-thickness box)
    uint32_t ref_nodes = 0, ref_max_depth = 0, leaves = 0, pruned = 0, max_depth = 0;
};

struct HostScene {
    zrt_scene_desc desc{}; // pointers into the vectors below
    std::vector<zrt_surface> surfaces;
    std::vector<zrt_sphere> spheres;
    std::vector<zrt_triangle> triangles;
    std::vector<zrt_material> materials;
    std::vector<zrt_texture> textures;
    std::vector<std::vector<uint8_t>> texels;
};

// Faithful rebuild of the reference tree (bvh.zig:62-185) followed by flattening and flat-box
// pruning.  sah = true additionally re-splits the surviving primitives with a binned SAH builder;
// slots (tie-break keys) always come from the reference tree.
void build_flat_bvh(const HostScene &scene, bool sah, FlatBvh *out);

} // namespace zrt
