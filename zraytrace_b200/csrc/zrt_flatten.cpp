// zrt_flatten.cpp — host side of the acceleration structure (runs once per scene, like the
// reference's "prepare" phase raytrace.zig:150,200).
//
// 1. Rebuild the reference's object-split tree (bvh.zig:62-185) on index arrays.  Hit results of the
//    reference depend on that tree in exactly two ways, and both are extracted here:
//      * ties in t are won by the surface that comes first in left-first DFS order
//        (bvh.zig:193-204 accepts the right child only if strictly closer)            -> slot numbers
//      * a node whose box has zero thickness on an axis is never entered, because
//        aabb.zig:121 rejects on `tmax <= tmin`                            -> pruning
//    Both are known where the build creates its leaves, so no pass over the finished tree is needed.  From 2048
//    surfaces on the tree is built from presorted index lists without a sort inside the recursion (RefTree below).
// 2. ZRT_FLAG_BVH_REFERENCE: flatten what survives into 64-byte two-child nodes (DevNode) in DFS pre-order.
// 3. Default: throw the reference topology away and re-split the surviving primitives with a binned surface-area
//    heuristic; slots keep the reference order, so the hits are the same and only the number of node fetches
//    changes.
// Measured phase by phase in DESIGN.md section 5.2 (ZRT_TIMING=1 prints them).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

#include "zrt_internal.h"

namespace zrt {
namespace {

struct alignas(16) Box {
    float mn[4], mx[4]; // lane 3 is padding (one SSE register per corner)
};
inline float zmin(float x, float y) { return (x < y) ? x : y; } // Zig std.math.min
inline float zmax(float x, float y) { return (x > y) ? x : y; }
inline Box boxUnion(const Box &a, const Box &b) { // aabb.zig:68-71
    Box r;
#if defined(__SSE__)
    // MINPS / MAXPS are exactly the two ternaries above, operand order included (second operand on equality)
    _mm_store_ps(r.mn, _mm_min_ps(_mm_load_ps(a.mn), _mm_load_ps(b.mn)));
    _mm_store_ps(r.mx, _mm_max_ps(_mm_load_ps(a.mx), _mm_load_ps(b.mx)));
#else
    for (int i = 0; i < 4; i++) {
        r.mn[i] = zmin(a.mn[i], b.mn[i]);
        r.mx[i] = zmax(a.mx[i], b.mx[i]);
    }
#endif
    return r;
}
inline Box boxEmpty() {
    const float inf = std::numeric_limits<float>::infinity();
    return Box{{inf, inf, inf, inf}, {-inf, -inf, -inf, -inf}};
}
inline bool boxFlat(const Box &b) { return b.mn[0] == b.mx[0] || b.mn[1] == b.mx[1] || b.mn[2] == b.mx[2]; }
inline float boxScore(const Box &b) { // aabb.zig:99-105: 2*(dx^2+dy^2+dz^2), not an area
    const float dx = std::fabs(b.mn[0] - b.mx[0]), dy = std::fabs(b.mn[1] - b.mx[1]), dz = std::fabs(b.mn[2] - b.mx[2]);
    return 2 * (dx * dx + dy * dy + dz * dz);
}

// Array whose elements start uninitialised: a std::vector would zero tens of MB on one thread (and fault the pages in
// there) before the parallel passes that fill them.
template <class T>
struct Raw {
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void alloc(size_t count) { p.reset(new T[count]); n = count; }
    size_t size() const { return n; }
    T *data() { return p.get(); }
    const T *data() const { return p.get(); }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// f() on its own thread, or in line when a thread would cost more than the work
struct Maybe {
    Worker w;
    template <class F>
    Maybe(bool threaded, F f) {
        if (threaded) w = Worker(f);
        else f();
    }
    void join() { w.join(); }
};

struct RefNode {          // one bvh.zig BVHNode
    Box box;
    int32_t left, right;  // >= 0 node index, < 0: ~surface id (a surface referenced directly)
};

struct RefTree {
    const HostScene &sc;
    Raw<Box> sbox;            // per surface: aabb min/max
    Raw<float> smid[3];       // per surface: aabb midpoint (the sort key, bvh.zig:38-48)
    Raw<RefNode> nodes;       // preallocated (< 2n nodes); slots handed out atomically
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<uint32_t> max_depth{0};
    // The two halves of a split are independent (disjoint sub-ranges of the id array, nodes allocated
    // atomically), so the top levels of the recursion fan out over host threads.  The tree is the same
    // tree; only the node numbering (never observable) depends on timing.
    static constexpr uint32_t kParallelDepth = 4;
    // A half of a split gets its own thread from this many primitives on: 8192 on large meshes (at most 64 threads are created per
    // build and they must go to the large subtrees: 322 k triangles with a fixed 768 cost +55 ms), down to 768 on small ones, where a
    // thread's ~20 us is still a fraction of the half's work (teapot, 6 k triangles: c2 end to end 12.4 -> 11.2 ms)
    static size_t parallelMin(size_t n_total) { return std::max<size_t>(768, std::min<size_t>(8192, n_total / 8)); }
    static constexpr size_t kParallelMin = 8192; // the literal (sort in the recursion) rebuild, by depth

    explicit RefTree(const HostScene &s) : sc(s) {
        const size_t n = sc.surfaces.size();
        nodes.alloc(2 * n + 2);
        sbox.alloc(n);
        for (auto &m : smid) m.alloc(n);
        parallelFor(n, 16384, [this](size_t begin, size_t end) { for (size_t i = begin; i < end; i++) {
            Box b{};
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
        } });
    }

    Box rangeBox(const uint32_t *ids, size_t n) const { // bvh.zig:62-69 surfaces_to_aabb (min/max are exact)
        Box b = boxEmpty();
        for (size_t i = 0; i < n; i++) b = boxUnion(b, sbox[ids[i]]);
        return b;
    }
    // bvh.zig:71-72: std.sort.sort is a stable comparison sort on `a.midpoint < b.midpoint`.  Small ranges use
    // std::stable_sort; large ones an LSD radix sort on the order-preserving integer image of the float key, which
    // is stable too and therefore yields the same permutation (-0.0 is folded onto +0.0 first, because the
    // comparison treats them as equal; midpoints are never NaN).  a, b: scratch of n (key << 32 | id) words.
    void sortAxis(int axis, uint32_t *ids, size_t n, uint64_t *a, uint64_t *b) const {
        const float *key = smid[axis].data();
        if (n < 4096) {
            std::stable_sort(ids, ids + n, [key](uint32_t x, uint32_t y) { return key[x] < key[y]; });
            return;
        }
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
    // scratch for the literal build, indexed like the id array: concurrent subtree builds work on disjoint ranges
    const uint32_t *ids_base = nullptr;
    mutable std::vector<uint64_t> scratch_a, scratch_b;
    void sortAxis(int axis, uint32_t *ids, size_t n) const {
        if (n < 4096) sortAxis(axis, ids, n, nullptr, nullptr);
        else sortAxis(axis, ids, n, scratch_a.data() + (ids - ids_base), scratch_b.data() + (ids - ids_base));
    }
    // bvh.zig:85-120: candidate splits at n/4, n/2, n/4 + n/2 (n/2 alone below four surfaces) on each axis; the
    // first strictly smaller (left score + right score) / total score wins, axes in x, y, z order.
    struct SplitSearch {
        size_t splits[3];
        int n_splits;
        float total = 0.0f;
        int best_axis = 0;
        float best_ratio = std::numeric_limits<float>::infinity();
        size_t best_split;
        explicit SplitSearch(size_t n) : splits{n / 2, 0, 0}, n_splits(1), best_split(n / 2) {
            if (n >= 4) {
                splits[0] = n / 4; splits[1] = n / 2; splits[2] = n / 4 + n / 2;
                n_splits = 3;
            }
        }
    };
    // boxes of the segments between consecutive split points of a range sorted on one axis; the halves of every
    // candidate split (and the box of the whole range) are exact unions of them
    void segmentBoxes(const SplitSearch &ss, const uint32_t *ids, size_t n, Box seg[4]) const {
        size_t lo = 0;
        for (int k = 0; k <= ss.n_splits; k++) {
            const size_t hi = (k < ss.n_splits) ? ss.splits[k] : n;
            seg[k] = rangeBox(ids + lo, hi - lo);
            lo = hi;
        }
    }
    static void scoreAxis(SplitSearch &ss, int axis, const Box seg[4]) {
        for (int k = 0; k < ss.n_splits; k++) {
            Box left = boxEmpty(), right = boxEmpty();
            for (int j = 0; j <= k; j++) left = boxUnion(left, seg[j]);
            for (int j = k + 1; j <= ss.n_splits; j++) right = boxUnion(right, seg[j]);
            const float area = boxScore(right) + boxScore(left);
            const float ratio = area / ss.total;
            if (ratio < ss.best_ratio) {
                ss.best_ratio = ratio;
                ss.best_axis = axis;
                ss.best_split = ss.splits[k];
            }
        }
    }
    // The literal sequence.  The reference re-sorts before each of the three candidate splits of an axis; the 2nd
    // and 3rd sort of an already sorted range by the same key with a stable sort change nothing, so one sort per
    // axis reproduces the same permutation sequence.
    size_t optimalAxisDivide(uint32_t *ids, size_t n) const {
        SplitSearch ss(n);
        ss.total = boxScore(rangeBox(ids, n));
        for (int axis = 0; axis < 3; axis++) {
            sortAxis(axis, ids, n);
            Box seg[4];
            segmentBoxes(ss, ids, n, seg);
            scoreAxis(ss, axis, seg);
        }
        sortAxis(ss.best_axis, ids, n); // "redo the best split"
        return ss.best_split;
    }
    const Box &childBox(int32_t c) const { return c >= 0 ? nodes[c].box : sbox[~c]; }
    int32_t create(int32_t left, int32_t right) { // bvh.zig:162-169
        const uint32_t i = n_nodes.fetch_add(1);
        nodes[i] = RefNode{boxUnion(childBox(left), childBox(right)), left, right};
        return (int32_t)i;
    }
    // Slot numbers (position in left-first DFS order, the tie-break key of bvh.zig:193-204) fall out of the build: a
    // subtree over n surfaces owns the n slots from `base` on, its left half the first `split` of them.
    // So does the pruning (aabb.zig:121): node boxes are exact unions, hence nested, so a box that is flat
    // on an axis makes every box below it flat on that axis; conversely, if the node that references a surface
    // directly has a box with thickness, so do all its ancestors and the reference reaches the surface.
    std::vector<uint32_t> slot_surface, slot_of; // slot -> surface id, surface id -> slot
    std::vector<uint8_t> slot_visible;           // slot -> 0 if the surface sits under a zero-thickness box
    void leafSlot(size_t slot, uint32_t surface, bool visible) {
        slot_surface[slot] = surface;
        slot_of[surface] = (uint32_t)slot;
        slot_visible[slot] = visible ? 1 : 0;
    }
    int32_t divide(uint32_t *ids, size_t n, uint32_t depth, size_t base) { // bvh.zig:129-160
        uint32_t seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        if (n == 1) {
            leafSlot(base, ids[0], !boxFlat(sbox[ids[0]]));
            return create(~(int32_t)ids[0], ~(int32_t)ids[0]);
        }
        if (n == 2) { // left child = surfaces[1]
            const bool visible = !boxFlat(boxUnion(sbox[ids[1]], sbox[ids[0]]));
            leafSlot(base, ids[1], visible);
            leafSlot(base + 1, ids[0], visible);
            return create(~(int32_t)ids[1], ~(int32_t)ids[0]);
        }
        const size_t split = optimalAxisDivide(ids, n);
        int32_t l, r;
        if (depth <= kParallelDepth && n >= kParallelMin) {
            Worker left([&] { l = divide(ids, split, depth + 1, base); });
            r = divide(ids + split, n - split, depth + 1, base + split);
            left.join();
        } else {
            l = divide(ids, split, depth + 1, base);
            r = divide(ids + split, n - split, depth + 1, base + split);
        }
        return create(l, r);
    }

    // ---- the same tree without sorting inside the recursion ------------------------------------------------------
    // Every sort above is stable, so the order of a range after a sort is a fixed lexicographic order on the three
    // midpoint keys, ending in the input index, that does not depend on the range.  With F(p) the order the parent
    // left its range in, a node sees
    //      after the x sort   (x, F(p))          after the y sort   (y, x, F(p)) = (y, x, z, idx)
    //      after the z sort   (z, y, x, idx)     after the redo     (b, z, y, x, idx) = F(node), b the winning axis
    // and (x, F(p)) is (x, y, z, idx) if the parent split on y, (x, z, y, idx) otherwise.  That is five global orders;
    // they are sorted once (ten stable one-key sorts, the ones the reference does at its root among them) and every
    // split then partitions the five index lists stably instead of re-sorting: O(n) per node.  The root is the one
    // node whose incoming order is the input order itself: its x and y sorts give (x, idx) and (y, x, idx).
    // Ranges below kLiteralBelow run the literal code on a copy of the range in the parent's order.
    enum Order { O_XZY = 0, O_XYZ = 1, O_YXZ = 2, O_ZYX = 3, O_YZX = 4, N_ORDERS = 5 };
    static constexpr size_t kPresortMin = 2048;
    static constexpr size_t kLiteralBelow = 3;
    std::vector<uint32_t> ord[N_ORDERS];
    std::vector<uint32_t> part_tmp; // scratch of the partitions, indexed like the lists
    std::vector<uint8_t> side;      // per surface: which half of the current split it went to
    std::atomic<uint32_t> spawned{0};

    // One of the global orders (k0, k1, k2, idx) from the list sorted on (k0, idx): only runs of equal k0 are out of
    // place, and within a run the order is (k1, k2, idx).  Ties between midpoints are rare in a mesh, so this is a
    // scan; a run as long as the list (a flat sheet of triangles) costs one comparison sort.
    void refineRuns(const std::vector<uint32_t> &base, int k0, int k1, int k2, std::vector<uint32_t> *out) const {
        *out = base;
        const float *a = smid[k0].data(), *b = smid[k1].data(), *c = k2 >= 0 ? smid[k2].data() : nullptr;
        uint32_t *ids = out->data();
        const size_t n = out->size();
        for (size_t i = 0; i < n;) {
            size_t j = i + 1;
            while (j < n && a[ids[j]] == a[ids[i]]) j++; // == is the comparator's notion of a tie (-0.0 == +0.0)
            if (j - i > 1)
                std::sort(ids + i, ids + j, [b, c](uint32_t p, uint32_t q) {
                    if (b[p] != b[q]) return b[p] < b[q];
                    if (c && c[p] != c[q]) return c[p] < c[q];
                    return p < q;
                });
            i = j;
        }
    }
    static constexpr size_t kWideNode = 65536; // above this the passes over one node run on separate threads too
    void partitionList(int k, size_t lo, size_t n, uint32_t *tmp) {
        uint32_t *list = ord[k].data() + lo;
        size_t a = 0, b = 0;
        for (size_t i = 0; i < n; i++) {
            const uint32_t id = list[i];
            if (side[id]) tmp[b++] = id;
            else list[a++] = id;
        }
        std::memcpy(list + a, tmp, b * sizeof(uint32_t));
    }
    void partitionLists(size_t lo, size_t n, int final_order, size_t at) {
        const uint32_t *f = ord[final_order].data() + lo;
        for (size_t i = 0; i < at; i++) side[f[i]] = 0;
        for (size_t i = at; i < n; i++) side[f[i]] = 1;
        if (n >= kWideNode) { // four lists, four threads (the final order's own list doubles as the fourth scratch)
            Worker th[N_ORDERS];
            std::vector<uint32_t> extra[N_ORDERS];
            int used = 0;
            for (int k = 0; k < N_ORDERS; k++) {
                if (k == final_order) continue;
                uint32_t *tmp = part_tmp.data() + lo;
                if (used++) { extra[k].resize(n); tmp = extra[k].data(); }
                th[k] = Worker([this, k, lo, n, tmp] { partitionList(k, lo, n, tmp); });
            }
            for (int k = 0; k < N_ORDERS; k++) if (k != final_order) th[k].join();
            return;
        }
        for (int k = 0; k < N_ORDERS; k++)
            if (k != final_order) partitionList(k, lo, n, part_tmp.data() + lo);
    }
    int32_t splitPresorted(size_t lo, size_t n, const uint32_t *const lists[3], uint32_t depth) {
        SplitSearch ss(n);
        Box seg[3][4];
        if (n >= kWideNode) {
            Worker ty([&] { segmentBoxes(ss, lists[1], n, seg[1]); });
            Worker tz([&] { segmentBoxes(ss, lists[2], n, seg[2]); });
            segmentBoxes(ss, lists[0], n, seg[0]);
            ty.join();
            tz.join();
        } else {
            for (int axis = 0; axis < 3; axis++) segmentBoxes(ss, lists[axis], n, seg[axis]);
        }
        Box whole = boxEmpty();
        for (int k = 0; k <= ss.n_splits; k++) whole = boxUnion(whole, seg[0][k]);
        ss.total = boxScore(whole);
        for (int axis = 0; axis < 3; axis++) scoreAxis(ss, axis, seg[axis]);
        const int final_order = ss.best_axis == 0 ? O_XZY : (ss.best_axis == 1 ? O_YZX : O_ZYX);
        const int child_x = ss.best_axis == 1 ? O_XYZ : O_XZY;
        const size_t at = ss.best_split;
        partitionLists(lo, n, final_order, at);
        int32_t l, r;
        if (std::min(at, n - at) >= parallelMin(sc.surfaces.size()) && spawned.fetch_add(1) < 64) {
            Worker left([&] { l = dividePresorted(lo, at, final_order, child_x, depth + 1); });
            r = dividePresorted(lo + at, n - at, final_order, child_x, depth + 1);
            left.join();
        } else {
            l = dividePresorted(lo, at, final_order, child_x, depth + 1);
            r = dividePresorted(lo + at, n - at, final_order, child_x, depth + 1);
        }
        return create(l, r);
    }
    int32_t dividePresorted(size_t lo, size_t n, int parent_final, int x_order, uint32_t depth) {
        if (n < kLiteralBelow) {
            uint32_t ids[kLiteralBelow];
            std::memcpy(ids, ord[parent_final].data() + lo, n * sizeof(uint32_t));
            return divide(ids, n, depth, lo);
        }
        uint32_t seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        const uint32_t *const lists[3] = {ord[x_order].data() + lo, ord[O_YXZ].data() + lo, ord[O_ZYX].data() + lo};
        return splitPresorted(lo, n, lists, depth);
    }
    int32_t buildPresorted() {
        const size_t n = sc.surfaces.size();
        const auto t_start = std::chrono::steady_clock::now();
        std::vector<uint32_t> idx(n), root_x, root_y;
        for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
        {
            std::vector<uint32_t> base[3]; // (x, idx), (y, idx), (z, idx): the one-key sorts of the input order
            auto sortBase = [&](int axis) {
                std::vector<uint64_t> sa(n), sb(n);
                base[axis] = idx;
                sortAxis(axis, base[axis].data(), n, sa.data(), sb.data());
            };
            // a thread costs more than sorting a few thousand keys: small scenes do the ten steps in line
            const bool mt = n >= 32768;
            Maybe ty(mt, [&] {
                sortBase(1);
                Maybe t1(mt, [&] { refineRuns(base[1], 1, 2, 0, &ord[O_YZX]); });
                Maybe t2(mt, [&] { refineRuns(base[1], 1, 0, -1, &root_y); }); // (y, x, idx): the root's y sort
                refineRuns(base[1], 1, 0, 2, &ord[O_YXZ]);
                t1.join();
                t2.join();
            });
            Maybe tz(mt, [&] { sortBase(2); refineRuns(base[2], 2, 1, 0, &ord[O_ZYX]); });
            sortBase(0);
            Maybe tx(mt, [&] { refineRuns(base[0], 0, 1, 2, &ord[O_XYZ]); });
            refineRuns(base[0], 0, 2, 1, &ord[O_XZY]);
            tx.join();
            ty.join();
            tz.join();
            root_x.swap(base[0]); // (x, idx): the root's x sort
        }
        part_tmp.resize(n);
        side.resize(n);
        max_depth.store(1);
        if (std::getenv("ZRT_TIMING")) {
            const auto t1 = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[zrt build]   of which presort     %8.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t_start).count());
        }
        const uint32_t *const lists[3] = {root_x.data(), root_y.data(), ord[O_ZYX].data()};
        return splitPresorted(0, n, lists, 1);
    }
    int32_t build(std::vector<uint32_t> *ids, bool literal) {
        const size_t n = ids->size();
        slot_surface.resize(n);
        slot_of.resize(n);
        slot_visible.resize(n);
        if (!literal && n >= kPresortMin) return buildPresorted();
        ids_base = ids->data();
        scratch_a.resize(n);
        scratch_b.resize(n);
        const int32_t root = divide(ids->data(), n, 1, 0); // bvh.zig:171-185
        std::vector<uint64_t>().swap(scratch_a);
        std::vector<uint64_t>().swap(scratch_b);
        return root;
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
    const std::vector<uint32_t> &slot_of; // surface id -> slot (DFS first visit), from the tree build
    const std::vector<uint32_t> &sphere_seq; // surface id -> index into the BVH-mode device sphere array
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
    const Box *pbox;      // per primitive
    const uint32_t *pref; // per primitive leaf ref
    // preallocated, one node per split.  A subtree over m primitives has exactly m - 1 nodes, so DFS pre-order
    // numbers (parent before its subtrees, left subtree contiguous: memory locality of the traversal, and a node order
    // that does not depend on thread timing) are known on the way down: left child = my + 1, right = my + |left|.
    decltype(FlatBvh::nodes) *nodes;
    std::atomic<uint32_t> max_depth{0};
    std::atomic<uint32_t> spawned{0};
    Raw<float> cen[3];
    size_t par_min = 8192; // RefTree::parallelMin(primitives)

    static float area(const Box &b) {
        const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
    // Binned split (16 bins per axis): partitions ids and returns the size of the left part.
    size_t splitBinned(uint32_t *ids, size_t n) {
        // the passes over a wide node are split over host threads: min / max and counts merge exactly, so the tree does
        // not depend on how many threads there are
        constexpr size_t kWide = 65536, kChunk = 16384;
        Box cb = boxEmpty();
        {
            std::mutex m;
            parallelFor(n >= kWide ? n : 1, kChunk, [&](size_t begin, size_t end) {
                if (n < kWide) end = n;
                Box c = boxEmpty();
                for (size_t i = begin; i < end; i++)
                    for (int k = 0; k < 3; k++) {
                        c.mn[k] = std::min(c.mn[k], cen[k][ids[i]]);
                        c.mx[k] = std::max(c.mx[k], cen[k][ids[i]]);
                    }
                std::lock_guard<std::mutex> g(m);
                for (int k = 0; k < 3; k++) {
                    cb.mn[k] = std::min(cb.mn[k], c.mn[k]);
                    cb.mx[k] = std::max(cb.mx[k], c.mx[k]);
                }
            });
        }
        constexpr int NB = 16;
        int best_axis = -1, best_bin = 0;
        float best_cost = std::numeric_limits<float>::infinity();
        // one sweep over the primitives fills the bins of all three axes
        struct Bins {
            Box bb[3][NB];
            size_t cnt[3][NB];
        };
        Bins bins;
        auto clearBins = [](Bins &b) {
            for (int axis = 0; axis < 3; axis++)
                for (int k = 0; k < NB; k++) { b.bb[axis][k] = boxEmpty(); b.cnt[axis][k] = 0; }
        };
        clearBins(bins);
        float lo3[3], scale3[3];
        bool use[3];
        for (int axis = 0; axis < 3; axis++) {
            const float ext = cb.mx[axis] - cb.mn[axis];
            use[axis] = ext > 0.0f;
            lo3[axis] = cb.mn[axis];
            scale3[axis] = use[axis] ? NB / ext : 0.0f;
        }
        auto fill = [&](Bins &b, size_t begin, size_t end) {
            for (size_t i = begin; i < end; i++) {
                const uint32_t id = ids[i];
                const Box &pb = pbox[id];
                for (int axis = 0; axis < 3; axis++) {
                    if (!use[axis]) continue;
                    int k = (int)((cen[axis][id] - lo3[axis]) * scale3[axis]);
                    k = std::min(std::max(k, 0), NB - 1);
                    b.cnt[axis][k]++;
                    b.bb[axis][k] = boxUnion(b.bb[axis][k], pb);
                }
            }
        };
        if (n >= kWide) {
            std::mutex m;
            parallelFor(n, kChunk, [&](size_t begin, size_t end) {
                Bins local;
                clearBins(local);
                fill(local, begin, end);
                std::lock_guard<std::mutex> g(m);
                for (int axis = 0; axis < 3; axis++)
                    for (int k = 0; k < NB; k++) {
                        bins.bb[axis][k] = boxUnion(bins.bb[axis][k], local.bb[axis][k]);
                        bins.cnt[axis][k] += local.cnt[axis][k];
                    }
            });
        } else {
            fill(bins, 0, n);
        }
        auto &bb = bins.bb;
        auto &cnt = bins.cnt;
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
        return mid;
    }
    Emitted build(uint32_t *ids, size_t n, uint32_t depth, uint32_t my) {
        if (n == 1) return Emitted{pref[ids[0]], pbox[ids[0]]};
        uint32_t seen = max_depth.load();
        while (depth > seen && !max_depth.compare_exchange_weak(seen, depth)) {}
        if (n == 2) { // half of all nodes: nothing to bin
            DevNode &d = (*nodes)[my];
            const Box &a = pbox[ids[0]], &b = pbox[ids[1]];
            d.set_box(0, a.mn, a.mx);
            d.set_box(1, b.mn, b.mx);
            d.left = pref[ids[0]]; d.right = pref[ids[1]]; d.pad0 = d.pad1 = 0;
            return Emitted{my, boxUnion(a, b)};
        }
        const size_t mid = splitBinned(ids, n);
        Emitted l, r;
        // SAH splits are uneven (a ground sphere, a small mesh next to a big one), so the fan-out over host threads goes
        // by subtree size, not by depth; `spawned` bounds the number of threads ever created per build
        const size_t small = std::min(mid, n - mid);
        if (small >= par_min && spawned.fetch_add(1) < 64) {
            Worker left([&] { l = build(ids, mid, depth + 1, my + 1); });
            r = build(ids + mid, n - mid, depth + 1, my + (uint32_t)mid);
            left.join();
        } else {
            l = build(ids, mid, depth + 1, my + 1);
            r = build(ids + mid, n - mid, depth + 1, my + (uint32_t)mid);
        }
        DevNode &d = (*nodes)[my];
        d.set_box(0, l.box.mn, l.box.mx);
        d.set_box(1, r.box.mn, r.box.mx);
        d.left = l.ref; d.right = r.ref; d.pad0 = d.pad1 = 0;
        return Emitted{my, boxUnion(l.box, r.box)};
    }
};

} // namespace

void build_flat_bvh(const HostScene &scene, bool sah, FlatBvh *out) {
    *out = FlatBvh{};
    const size_t n = scene.surfaces.size();
    if (n == 0) return;
    BuildLap lap_timer;
    auto lap = [&](const char *what) { lap_timer(what); };
    RefTree rt(scene);
    lap("surface boxes");
    std::vector<uint32_t> ids(n);
    for (size_t i = 0; i < n; i++) ids[i] = (uint32_t)i;
    // ZRT_BVH_BUILD=literal: the sort-in-the-recursion rebuild at every size (the presorted build's cross-check)
    const char *how = std::getenv("ZRT_BVH_BUILD");
    const int32_t root = rt.build(&ids, how && std::strcmp(how, "literal") == 0);
    lap("reference tree");
    out->ref_nodes = rt.n_nodes.load();
    out->ref_max_depth = rt.max_depth.load();

    out->slot_surface.swap(rt.slot_surface);
    std::vector<uint32_t> sphere_seq(n, 0); // surface id -> index into the BVH-mode device sphere array
    uint32_t nsph = 0;
    for (size_t i = 0; i < n; i++)
        if (scene.surfaces[i].kind == ZRT_SURFACE_SPHERE) sphere_seq[i] = nsph++;
    auto countLeaves = [&] {
        out->leaves = out->pruned = 0;
        for (uint8_t v : out->slot_visible) {
            out->leaves += v;
            out->pruned += !v;
        }
    };
    if (sah) { // the SAH rebuild needs only the set of reachable surfaces from the reference topology
        out->slot_visible.swap(rt.slot_visible);
        countLeaves();
    }
    if (!sah || out->leaves <= 1) { // the reference topology itself (also when there is nothing to re-split)
        Flattener fl{scene, rt, out, rt.slot_of, sphere_seq, std::vector<int8_t>(rt.n_nodes.load(), -1)};
        out->slot_visible.assign(n, 0); // emit() marks what it reaches
        out->nodes.reserve(n);
        out->root = fl.survives(root) ? fl.emit(root, 1).ref : REF_EMPTY;
        countLeaves();
    }
    lap("slots + flatten");
    if (sah && out->leaves > 1) {
        std::vector<uint32_t> vis; // the visible slots, in slot order
        vis.reserve(out->leaves);
        for (size_t s = 0; s < out->slot_surface.size(); s++)
            if (out->slot_visible[s]) vis.push_back((uint32_t)s);
        const size_t m = vis.size();
        Raw<Box> pbox;
        Raw<uint32_t> pref, pid;
        pbox.alloc(m); pref.alloc(m); pid.alloc(m);
        out->nodes.clear();
        out->nodes.resize(m - 1);
        SahBuilder sb{pbox.data(), pref.data(), &out->nodes, {}, {}, {}};
        sb.par_min = RefTree::parallelMin(m);
        for (int k = 0; k < 3; k++) sb.cen[k].alloc(m);
        parallelFor(m, 16384, [&](size_t begin, size_t end) {
            for (size_t i = begin; i < end; i++) {
                const uint32_t s = vis[i], surf = out->slot_surface[s];
                const Box &b = rt.sbox[surf];
                pbox[i] = b;
                pref[i] = scene.surfaces[surf].kind == ZRT_SURFACE_SPHERE ? (REF_LEAF | REF_SPHERE | sphere_seq[surf]) : (REF_LEAF | s);
                for (int k = 0; k < 3; k++) sb.cen[k][i] = 0.5f * (b.mn[k] + b.mx[k]);
                pid[i] = (uint32_t)i;
            }
        });
        lap("sah setup");
        out->root = sb.build(pid.data(), m, 1, 0).ref;
        out->max_depth = sb.max_depth.load();
        lap("sah build");
    }
}

} // namespace zrt
