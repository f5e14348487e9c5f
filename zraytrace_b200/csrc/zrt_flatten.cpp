// zrt_flatten.cpp — host side of the acceleration structure (runs once per scene, like the
// reference's "prepare" phase raytrace.zig:150,200).
//
// 1. Rebuild the reference's object-split tree (bvh.zig:62-185) on index arrays.  Hit results of the
//    reference depend on that tree in exactly two ways, and both are extracted here:
//      * ties in t are won by the surface that comes first in left-first DFS order
//        (bvh.zig:193-204 accepts the right child only if strictly closer)            -> slot numbers
//      * a node whose box has zero thickness on an axis is never entered, because
//        aabb.zig:121 rejects on `tmax <= tmin`                            -> pruning
// 2. Flatten what survives into 64-byte two-child nodes (DevNode) in DFS pre-order.
// 3. By default (unless ZRT_FLAG_BVH_REFERENCE) throw the reference topology away and re-split the surviving
//    primitives with a binned surface-area heuristic; slots keep the reference order, so the hits
//    are the same and only the number of node fetches changes.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

#include "zrt_internal.h"

namespace zrt {
namespace {

struct Box {
    float mn[3], mx[3];
};
inline float zmin(float x, float y) { return (x < y) ? x : y; } // Zig std.math.min
inline float zmax(float x, float y) { return (x > y) ? x : y; }
inline Box boxUnion(const Box &a, const Box &b) { // aabb.zig:68-71
    Box r;
    for (int i = 0; i < 3; i++) {
        r.mn[i] = zmin(a.mn[i], b.mn[i]);
        r.mx[i] = zmax(a.mx[i], b.mx[i]);
    }
    return r;
}
inline Box boxEmpty() {
    const float inf = std::numeric_limits<float>::infinity();
    return Box{{inf, inf, inf}, {-inf, -inf, -inf}};
}
inline bool boxFlat(const Box &b) { return b.mn[0] == b.mx[0] || b.mn[1] == b.mx[1] || b.mn[2] == b.mx[2]; }
inline float boxScore(const Box &b) { // aabb.zig:99-105: 2*(dx^2+dy^2+dz^2), not an area
    const float dx = std::fabs(b.mn[0] - b.mx[0]), dy = std::fabs(b.mn[1] - b.mx[1]), dz = std::fabs(b.mn[2] - b.mx[2]);
    return 2 * (dx * dx + dy * dy + dz * dz);
}

struct RefNode {          // one bvh.zig BVHNode
    Box box;
    int32_t left, right;  // >= 0 node index, < 0: ~surface id (a surface referenced directly)
};

struct RefTree {
    const HostScene &sc;
    std::vector<Box> sbox;            // per surface: aabb min/max
    std::vector<float> smid[3];       // per surface: aabb midpoint (the sort key, bvh.zig:38-48)
    std::vector<RefNode> nodes;             // preallocated (< 2n nodes); slots handed out atomically
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<uint32_t> max_depth{0};
    // The two halves of a split are independent (disjoint sub-ranges of the id array, nodes allocated
    // atomically), so the top levels of the recursion fan out over host threads.  The tree is the same
    // tree; only the node numbering (never observable) depends on timing.
    static constexpr uint32_t kParallelDepth = 4;
    static constexpr size_t kParallelMin = 8192;

    explicit RefTree(const HostScene &s) : sc(s) {
        const size_t n = sc.surfaces.size();
        nodes.resize(2 * n + 2);
        sbox.resize(n);
        for (auto &m : smid) m.resize(n);
        for (size_t i = 0; i < n; i++) {
            Box b;
            if (sc.surfaces[i].kind == ZRT_SURFACE_SPHERE) { // sphere.zig:24-29
                const zrt_sphere &s = sc.spheres[sc.surfaces[i].index];
                const float c[3] = {s.center.x, s.center.y, s.center.z};
                for (int k = 0; k < 3; k++) {
                    const float lo = c[k] - s.radius, hi = c[k] + s.radius;
                    b.mn[k] = zmin(lo, hi);
                    b.mx[k] = zmax(lo, hi);
                    smid[k][i] = (lo + hi) / 2.0f; // aabb.zig:30-34 on the two corners as given
                }
            } else { // triangle.zig:33: initAabb(initMinMax(a,b), initMinMax(a,c))
                const zrt_triangle &t = sc.triangles[sc.surfaces[i].index];
                const float a[3] = {t.a.x, t.a.y, t.a.z}, bb[3] = {t.b.x, t.b.y, t.b.z}, c[3] = {t.c.x, t.c.y, t.c.z};
                for (int k = 0; k < 3; k++) {
                    b.mn[k] = zmin(zmin(a[k], bb[k]), zmin(a[k], c[k]));
                    b.mx[k] = zmax(zmax(a[k], bb[k]), zmax(a[k], c[k]));
                    smid[k][i] = (b.mn[k] + b.mx[k]) / 2.0f;
                }
            }
            sbox[i] = b;
        }
    }

    Box rangeBox(const uint32_t *ids, size_t n) const { // bvh.zig:62-69 surfaces_to_aabb (min/max are exact)
        Box b = boxEmpty();
        for (size_t i = 0; i < n; i++) b = boxUnion(b, sbox[ids[i]]);
        return b;
    }
    // bvh.zig:71-72: std.sort.sort is a stable comparison sort on `a.midpoint < b.midpoint`.  Small ranges use
    // std::stable_sort; large ones an LSD radix sort on the order-preserving integer image of the float key, which
    // is stable too and therefore yields the same permutation (-0.0 is folded onto +0.0 first, because the
    // comparison treats them as equal; midpoints are never NaN).
    // scratch for the radix sort, indexed like the id array: concurrent subtree builds work on disjoint ranges
    const uint32_t *ids_base = nullptr;
    mutable std::vector<uint64_t> scratch_a, scratch_b;
    void sortAxis(int axis, uint32_t *ids, size_t n) const {
        const float *key = smid[axis].data();
        if (n < 4096) {
            std::stable_sort(ids, ids + n, [key](uint32_t a, uint32_t b) { return key[a] < key[b]; });
            return;
        }
        uint64_t *a = scratch_a.data() + (ids - ids_base), *b = scratch_b.data() + (ids - ids_base); // (key << 32) | id
        size_t count[4][257];
        std::memset(count, 0, sizeof(count));
        for (size_t i = 0; i < n; i++) {
            uint32_t u;
            const float k = key[ids[i]] + 0.0f; // -0.0 + 0.0 = +0.0
            std::memcpy(&u, &k, 4);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            a[i] = ((uint64_t)u << 32) | ids[i];
            count[0][(u & 0xff) + 1]++;
            count[1][((u >> 8) & 0xff) + 1]++;
            count[2][((u >> 16) & 0xff) + 1]++;
            count[3][(u >> 24) + 1]++;
        }
        for (int pass = 0; pass < 4; pass++) {
            size_t *c = count[pass];
            bool trivial = false; // every key has the same digit: the pass would be the identity
            for (int k = 1; k <= 256 && !trivial; k++) trivial = c[k] == n;
            if (trivial) continue;
            const int shift = 32 + 8 * pass;
            for (int k = 0; k < 256; k++) c[k + 1] += c[k];
            for (size_t i = 0; i < n; i++) b[c[(a[i] >> shift) & 0xff]++] = a[i];
            std::swap(a, b);
        }
        for (size_t i = 0; i < n; i++) ids[i] = (uint32_t)a[i];
    }
    // bvh.zig:85-120.  The reference re-sorts before each of the three candidate splits of an axis;
    // the 2nd and 3rd sort of an already sorted range by the same key with a stable sort change
    // nothing, so one sort per axis reproduces the same permutation sequence.
    size_t optimalAxisDivide(uint32_t *ids, size_t n) const {
        int best_axis = 0;
        float best_ratio = std::numeric_limits<float>::infinity();
        size_t best_split = n / 2;
        const float total = boxScore(rangeBox(ids, n));
        size_t splits[3] = {n / 2, 0, 0};
        int n_splits = 1;
        if (n >= 4) {
            splits[0] = n / 4; splits[1] = n / 2; splits[2] = n / 4 + n / 2;
            n_splits = 3;
        }
        for (int axis = 0; axis < 3; axis++) {
            sortAxis(axis, ids, n);
            // boxes of the segments between consecutive split points; halves are exact unions of them
            Box seg[4];
            size_t lo = 0;
            for (int k = 0; k <= n_splits; k++) {
                const size_t hi = (k < n_splits) ? splits[k] : n;
                seg[k] = rangeBox(ids + lo, hi - lo);
                lo = hi;
            }
            for (int k = 0; k < n_splits; k++) {
                Box left = boxEmpty(), right = boxEmpty();
                for (int j = 0; j <= k; j++) left = boxUnion(left, seg[j]);
                for (int j = k + 1; j <= n_splits; j++) right = boxUnion(right, seg[j]);
                const float area = boxScore(right) + boxScore(left);
                const float ratio = area / total;
                if (ratio < best_ratio) {
                    best_ratio = ratio;
                    best_axis = axis;
                    best_split = splits[k];
                }
            }
        }
        sortAxis(best_axis, ids, n); // "redo the best split"
        return best_split;
    }
    const Box &childBox(int32_t c) const { return c >= 0 ? nodes[c].box : sbox[~c]; }
    int32_t create(int32_t left, int32_t right) { // bvh.zig:162-169
        const uint32_t i = n_nodes.fetch_add(1);
        nodes[i] = RefNode{boxUnion(childBox(left), childBox(right)), left, right};
        return (int32_t)i;
    }
    int32_t divide(uint32_t *ids, size_t n, uint32_t depth) { // bvh.zig:129-160
        uint32_t seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        if (n == 1) return create(~(int32_t)ids[0], ~(int32_t)ids[0]);
        if (n == 2) return create(~(int32_t)ids[1], ~(int32_t)ids[0]);
        const size_t split = optimalAxisDivide(ids, n);
        int32_t l, r;
        if (depth <= kParallelDepth && n >= kParallelMin) {
            std::thread left([&] { l = divide(ids, split, depth + 1); });
            r = divide(ids + split, n - split, depth + 1);
            left.join();
        } else {
            l = divide(ids, split, depth + 1);
            r = divide(ids + split, n - split, depth + 1);
        }
        return create(l, r);
    }
};

struct Emitted {
    uint32_t ref;
    Box box;
};

struct Flattener {
    const HostScene &sc;
    const RefTree &rt;
    FlatBvh *out;
    std::vector<uint32_t> slot_of; // surface id -> slot (DFS first visit)
    std::vector<uint32_t> sphere_seq; // surface id -> index into the BVH-mode device sphere array

    void assignSlots(int32_t c) {
        if (c < 0) {
            const uint32_t s = (uint32_t)~c;
            if (slot_of[s] == UINT32_MAX) {
                slot_of[s] = (uint32_t)out->slot_surface.size();
                out->slot_surface.push_back(s);
                out->slot_visible.push_back(0);
            }
            return;
        }
        assignSlots(rt.nodes[c].left);
        if (rt.nodes[c].right != rt.nodes[c].left) assignSlots(rt.nodes[c].right);
    }
    uint32_t leafRef(uint32_t surface) {
        out->slot_visible[slot_of[surface]] = 1;
        if (sc.surfaces[surface].kind == ZRT_SURFACE_SPHERE) return REF_LEAF | REF_SPHERE | sphere_seq[surface];
        return REF_LEAF | slot_of[surface];
    }
    std::vector<int8_t> alive; // per reference node: -1 unknown, 0 never entered, 1 reachable
    // Can the reference ever reach a surface through child c?  A surface referenced directly by a
    // two-leaf node is tested without a box of its own (bvh.zig:193-199), so it always survives; a
    // BVHNode survives unless its own box is flat or nothing below it survives.
    bool survives(int32_t c) {
        if (c < 0) return true;
        if (alive[c] >= 0) return alive[c] != 0;
        const RefNode &n = rt.nodes[c];
        bool ok;
        if (boxFlat(n.box)) ok = false;
        else if (n.left == n.right && n.left < 0) ok = true; // node(s,s) with a non-flat box
        else ok = survives(n.left) | survives(n.right);
        alive[c] = ok ? 1 : 0;
        return ok;
    }
    // precondition: survives(c)
    Emitted emit(int32_t c, uint32_t depth) {
        if (c < 0) return Emitted{leafRef((uint32_t)~c), rt.sbox[~c]};
        const RefNode &n = rt.nodes[c];
        if (n.left == n.right && n.left < 0) return Emitted{leafRef((uint32_t)~n.left), n.box}; // node(s,s)
        const bool sl = survives(n.left), sr = survives(n.right);
        if (!sr) return emit(n.left, depth); // a one-sided node collapses into its surviving child
        if (!sl) return emit(n.right, depth);
        const uint32_t my = (uint32_t)out->nodes.size();
        out->nodes.emplace_back(); // pre-order: parent before its subtrees, left subtree contiguous
        if (depth > out->max_depth) out->max_depth = depth;
        const Emitted l = emit(n.left, depth + 1);
        const Emitted r = emit(n.right, depth + 1);
        writeNode(my, l, r);
        return Emitted{my, boxUnion(l.box, r.box)};
    }
    void writeNode(uint32_t idx, const Emitted &l, const Emitted &r) {
        DevNode &d = out->nodes[idx];
        d.set_box(0, l.box.mn, l.box.mx);
        d.set_box(1, r.box.mn, r.box.mx);
        d.left = l.ref; d.right = r.ref; d.pad0 = d.pad1 = 0;
    }
};

// ---- binned SAH rebuild over the surviving primitives (next-row component) ----
struct SahBuilder {
    const std::vector<Box> &pbox; // per primitive
    const std::vector<uint32_t> &pref; // per primitive leaf ref
    std::vector<DevNode> *nodes;  // preallocated; build() hands slots out atomically, renumber() restores pre-order
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<uint32_t> max_depth{0};
    std::atomic<uint32_t> spawned{0};
    std::vector<float> cen[3];

    static float area(const Box &b) {
        const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
    Emitted build(uint32_t *ids, size_t n, uint32_t depth) {
        if (n == 1) return Emitted{pref[ids[0]], pbox[ids[0]]};
        uint32_t seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        Box bounds = boxEmpty(), cb = boxEmpty();
        for (size_t i = 0; i < n; i++) {
            bounds = boxUnion(bounds, pbox[ids[i]]);
            for (int k = 0; k < 3; k++) {
                cb.mn[k] = std::min(cb.mn[k], cen[k][ids[i]]);
                cb.mx[k] = std::max(cb.mx[k], cen[k][ids[i]]);
            }
        }
        constexpr int NB = 16;
        int best_axis = -1, best_bin = 0;
        float best_cost = std::numeric_limits<float>::infinity();
        // one sweep over the primitives fills the bins of all three axes
        Box bb[3][NB];
        size_t cnt[3][NB];
        float lo3[3], scale3[3];
        bool use[3];
        for (int axis = 0; axis < 3; axis++) {
            const float ext = cb.mx[axis] - cb.mn[axis];
            use[axis] = ext > 0.0f;
            lo3[axis] = cb.mn[axis];
            scale3[axis] = use[axis] ? NB / ext : 0.0f;
            for (int k = 0; k < NB; k++) { bb[axis][k] = boxEmpty(); cnt[axis][k] = 0; }
        }
        for (size_t i = 0; i < n; i++) {
            const uint32_t id = ids[i];
            const Box &pb = pbox[id];
            for (int axis = 0; axis < 3; axis++) {
                if (!use[axis]) continue;
                int k = (int)((cen[axis][id] - lo3[axis]) * scale3[axis]);
                k = std::min(std::max(k, 0), NB - 1);
                cnt[axis][k]++;
                bb[axis][k] = boxUnion(bb[axis][k], pb);
            }
        }
        for (int axis = 0; axis < 3; axis++) {
            if (!use[axis]) continue;
            float right_area[NB];
            size_t right_cnt[NB];
            Box acc = boxEmpty();
            size_t c = 0;
            for (int k = NB - 1; k > 0; k--) {
                acc = boxUnion(acc, bb[axis][k]);
                c += cnt[axis][k];
                right_area[k] = area(acc);
                right_cnt[k] = c;
            }
            acc = boxEmpty();
            c = 0;
            for (int k = 0; k < NB - 1; k++) {
                acc = boxUnion(acc, bb[axis][k]);
                c += cnt[axis][k];
                if (c == 0 || right_cnt[k + 1] == 0) continue;
                const float cost = area(acc) * (float)c + right_area[k + 1] * (float)right_cnt[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
            }
        }
        size_t mid;
        if (best_axis < 0) {
            mid = n / 2; // all centroids coincide: split by count
        } else {
            const float lo = cb.mn[best_axis], scale = NB / (cb.mx[best_axis] - cb.mn[best_axis]);
            const float *c = cen[best_axis].data();
            uint32_t *m = std::partition(ids, ids + n, [&](uint32_t id) {
                int b = (int)((c[id] - lo) * scale);
                b = std::min(std::max(b, 0), NB - 1);
                return b <= best_bin;
            });
            mid = (size_t)(m - ids);
            if (mid == 0 || mid == n) mid = n / 2;
        }
        const uint32_t my = n_nodes.fetch_add(1);
        Emitted l, r;
        // SAH splits are uneven (a ground sphere, a small mesh next to a big one), so the fan-out over host threads goes
        // by subtree size, not by depth; `spawned` bounds the number of threads ever created per build
        const size_t small = std::min(mid, n - mid);
        if (small >= RefTree::kParallelMin && spawned.fetch_add(1) < 64) {
            std::thread left([&] { l = build(ids, mid, depth + 1); });
            r = build(ids + mid, n - mid, depth + 1);
            left.join();
        } else {
            l = build(ids, mid, depth + 1);
            r = build(ids + mid, n - mid, depth + 1);
        }
        DevNode &d = (*nodes)[my];
        d.set_box(0, l.box.mn, l.box.mx);
        d.set_box(1, r.box.mn, r.box.mx);
        d.left = l.ref; d.right = r.ref; d.pad0 = d.pad1 = 0;
        return Emitted{my, boxUnion(l.box, r.box)};
    }
    // DFS pre-order renumbering (parent before its subtrees, left subtree contiguous): memory locality of the
    // traversal, and a node order that does not depend on thread timing
    uint32_t renumber(uint32_t ref, const std::vector<DevNode> &src, std::vector<DevNode> *dst) const {
        if (ref & REF_LEAF) return ref;
        const uint32_t my = (uint32_t)dst->size();
        dst->push_back(src[ref]);
        const uint32_t l = renumber(src[ref].left, src, dst);
        const uint32_t r = renumber(src[ref].right, src, dst);
        (*dst)[my].left = l;
        (*dst)[my].right = r;
        return my;
    }
};

} // namespace

void build_flat_bvh(const HostScene &scene, bool sah, FlatBvh *out) {
    *out = FlatBvh{};
    const size_t n = scene.surfaces.size();
    if (n == 0) return;
    const bool timing = std::getenv("ZRT_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[zrt build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    RefTree rt(scene);
    lap("surface boxes");
    std::vector<uint32_t> ids(n);
    for (size_t i = 0; i < n; i++) ids[i] = (uint32_t)i;
    rt.ids_base = ids.data();
    rt.scratch_a.resize(n);
    rt.scratch_b.resize(n);
    const int32_t root = rt.divide(ids.data(), n, 1); // bvh.zig:171-185
    std::vector<uint64_t>().swap(rt.scratch_a);
    std::vector<uint64_t>().swap(rt.scratch_b);
    lap("reference tree");
    out->ref_nodes = rt.n_nodes.load();
    out->ref_max_depth = rt.max_depth.load();

    Flattener fl{scene, rt, out, std::vector<uint32_t>(n, UINT32_MAX), std::vector<uint32_t>(n, 0),
                 std::vector<int8_t>(rt.n_nodes.load(), -1)};
    uint32_t nsph = 0;
    for (size_t i = 0; i < n; i++)
        if (scene.surfaces[i].kind == ZRT_SURFACE_SPHERE) fl.sphere_seq[i] = nsph++;
    fl.assignSlots(root);
    out->root = fl.survives(root) ? fl.emit(root, 1).ref : REF_EMPTY;
    for (uint8_t v : out->slot_visible) {
        out->leaves += v;
        out->pruned += !v;
    }
    lap("slots + flatten");
    if (sah && out->leaves > 1) {
        std::vector<Box> pbox;
        std::vector<uint32_t> pref;
        for (size_t s = 0; s < out->slot_surface.size(); s++) {
            if (!out->slot_visible[s]) continue;
            const uint32_t surf = out->slot_surface[s];
            pbox.push_back(rt.sbox[surf]);
            pref.push_back(scene.surfaces[surf].kind == ZRT_SURFACE_SPHERE ? (REF_LEAF | REF_SPHERE | fl.sphere_seq[surf])
                                                                           : (REF_LEAF | (uint32_t)s));
        }
        std::vector<DevNode> scratch(pbox.size() + 1);
        SahBuilder sb{pbox, pref, &scratch};
        for (int k = 0; k < 3; k++) {
            sb.cen[k].resize(pbox.size());
            for (size_t i = 0; i < pbox.size(); i++) sb.cen[k][i] = 0.5f * (pbox[i].mn[k] + pbox[i].mx[k]);
        }
        std::vector<uint32_t> pid(pbox.size());
        for (size_t i = 0; i < pid.size(); i++) pid[i] = (uint32_t)i;
        lap("sah setup");
        const uint32_t root_ref = sb.build(pid.data(), pid.size(), 1).ref;
        lap("sah build");
        out->nodes.clear();
        out->nodes.reserve(sb.n_nodes.load());
        out->root = sb.renumber(root_ref, scratch, &out->nodes);
        out->max_depth = sb.max_depth.load();
        lap("sah renumber");
    }
}

} // namespace zrt
