// zrt_pool_bvh.cuh — K1p: the BVH path tracer over a slot pool (included by zrt_kernels.cu, inside namespace zrt).
//
// K1 and K1w keep ONE path per lane.  On the BVH configurations ncu shows what that costs (profiles/r2_*_k_trace_ws*):
// after a closest-hit query the lanes of a warp want up to six different things (sky + next sample, sphere / triangle
// hit record, Lambertian, metal, glass, texture lookup) and the warp runs all of them one after the other at a few lanes
// each; and a lane that has finished its traversal idles until enough neighbours have finished too.  K1p separates the
// two roles:
//   * a SLOT (shared memory, N per warp, N > 32) is a work item with its current path: accumulator, throughput, the ray
//     to trace or the hit to shade.  A slot is always in exactly one place: a ring (TRAV: "ray waits for a traversal
//     lane"; REGEN / LAMB / METAL / GLASS: "hit waits to be shaded with 31 others of its kind"), a lane, or nowhere
//     (the item queue is exhausted).
//   * a LANE is a traversal engine: it owns the query of one slot in registers (origin, direction, reciprocals, closest
//     hit, stack) and is re-armed from the TRAV ring the moment it finishes (section X), so the node / leaf sections of
//     K1w run at the occupancy of "lanes that are mid-traversal" instead of "lanes whose whole path is mid-traversal";
//   * shading runs in batches of up to 32 slots of ONE kind (section S): convergent by construction.
// Sections per warp iteration, chosen by the warp's scheduler from ballots and ring counts (warp-uniform):
//   X  hand a finished query over to its slot, classify it, push it on the ring of its kind; re-arm free lanes
//   S  pop a batch of one kind, shade (K1q's code generalised to spheres + triangles), Ray.init, depth bookkeeping,
//      push the new ray on TRAV (or the ended path on REGEN)
//   N  one BVH node step for the lanes at an inner node          (K1w's section, same arithmetic)
//   L  one leaf test for the lanes that postponed one            (K1w's section)
// Every path sees the arithmetic, the RNG keys and the order of its item's f32 sum of K1, so images and counters are
// bit-identical to K1 / K1w (tests/test_gpu_parity.py, tests/test_gpu_full_size.py).
#pragma once

enum BpState : uint32_t { BS_NODE = 0, BS_LEAF = 1, BS_DONE = 2, BS_FREE = 3 };
constexpr uint32_t BP_TRAV_RING = 4; // ring index of TRAV; 0..3 are the PoolKind rings (REGEN, LAMB, METAL, GLASS)

template <int N, int RING>
struct alignas(16) BPoolSlots {
    float ox[N], oy[N], oz[N]; // TRAV: ray origin.  Shade rings: the hit location o + d t (ray.zig:14-16)
    float dx[N], dy[N], dz[N]; // unit direction of the ray (Ray.init, ray.zig:11-13)
    float tr[N], tg[N], tb[N]; // throughput of the path
    float ar[N], ag[N], ab[N]; // the item's f32 sum (raytrace.zig:156,177)
    float hu[N], hv[N];        // triangle barycentrics of the pending hit (triangle.zig:66)
    uint32_t href[N];          // leaf ref of the pending hit
    uint32_t pxy[N], meta[N];  // px | py << 16; K1q's meta word (PM_*), hit bits unused
    uint8_t ring[5][RING];
};

// hit_record.zig:28-41 for a hit whose location is already known (hit_record<MODE> computes it first)
DI void hit_record_at(const KParams &P, V3 loc, V3 d, uint32_t href, float hu, float hv, Surf &s) {
    const uint32_t idx = href & REF_INDEX_MASK;
    s.loc = loc;
    V3 on;
    if (href & REF_SPHERE) {
        const float4 a = ldg4(reinterpret_cast<const float4 *>(P.spheres + idx));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(P.spheres + idx) + 1);
        on = (loc - mk(a.x, a.y, a.z)) * __uint_as_float(b.x); // sphere.zig:46
        s.material = b.y;
        s.surface_id = b.z;
        s.tu = s.tv = 0.0f;
        if (b.y & MAT_IMAGE_BIT) sphere_uv(P, on, s.tu, s.tv);
    } else {
        const float nx = ldg4(P.triA + idx).w, ny = ldg4(P.triE1 + idx).w, nz = ldg4(P.triE2 + idx).w;
        on = unit(mk(nx, ny, nz)); // triangle.zig:36 face_unit_normal
        const TriMeta m = P.triMeta[idx];
        s.material = m.material;
        s.surface_id = m.surface_id;
        s.tu = hu;
        s.tv = hv;
    }
    s.front = !(dot(d, on) > 0.0f); // hit_record.zig:29
    s.normal = s.front ? on : neg(on);
}

template <int N, int RING, int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_trace_bpool(const __grid_constant__ KParams P) {
    static_assert((RING & (RING - 1)) == 0 && RING >= N && N >= 32 && RING <= 128, "ring: power of two >= N, counts live in bytes");
    __shared__ BPoolSlots<N, RING> pools[4];
    BPoolSlots<N, RING> &S = pools[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t L = P.lanes;
    const uint32_t total_items = P.x_end * P.height * L;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const float F_INF = __int_as_float(0x7f800000);
    constexpr uint32_t RM = RING - 1;
    ItemQueue iq;
    uint32_t n_refl = 0, n_bg = 0, n_depth = 0; // pixels, samples and rays: k_finish_counters (see K1)

    // ring state, warp-uniform: one byte per shade ring; TRAV on its own
    uint32_t heads = 0, counts = (uint32_t)N << (8 * PK_REGEN);
    uint32_t trav_head = 0, trav_count = 0;
    for (uint32_t s = lane; s < (uint32_t)N; s += 32u) { // every slot starts without an item, waiting for one
        S.ring[PK_REGEN][s] = (uint8_t)s;
        S.meta[s] = 0;
        S.ar[s] = S.ag[s] = S.ab[s] = 0.0f;
    }
    __syncwarp();

    // the lane's traversal engine
    uint32_t slot = 0, st = BS_FREE, cur = REF_EMPTY;
    int sp = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1), inv = mk(0, 0, 0);
    Hit h;
    h.t = F_INF; h.ref = REF_EMPTY; h.slot = 0xFFFFFFFFu; h.u = h.v = 0.0f;
    uint2 stack[TRAVERSAL_STACK]; // (ref, entry distance of the subtree)

    uint32_t burst_floor = 33u; // node steps run back to back while at least this many lanes are at an inner node (33: decide anew)
    for (;;) {
        // ---- scheduler (warp-uniform).  A decision for N holds for a burst: the warp keeps stepping nodes on ONE ballot per
        //      step until a quarter of the lanes that started the burst have left the inner nodes; every other section
        //      is followed by a full decision ----
        ZRT_PROF_TICK();
        ZRT_PROF(0, true);
        const uint32_t n_node = __popc(__ballot_sync(0xffffffffu, st == BS_NODE));
        int section = 0; // 0 = N, 1 = L, 2 = X, 3 = S
        uint32_t k = 0;  // S: the kind to shade
        if (n_node < burst_floor) {
            ZRT_PROF(1, true);
            burst_floor = 33u;
            const uint32_t n_leaf = __popc(__ballot_sync(0xffffffffu, st == BS_LEAF));
            const uint32_t n_done = __popc(__ballot_sync(0xffffffffu, st == BS_DONE));
            const uint32_t n_idle = 32u - n_node - n_leaf; // done + free
            const bool x_useful = n_done > 0u || (trav_count > 0u && n_idle > 0u);
            const bool hungry = n_idle >= P.ws_shade_min; // enough lanes without a query to make a refill worth a section
            const bool full_batch = (counts & 0xE0E0E0E0u) != 0u; // some shade ring holds >= 32 slots
            if (hungry && x_useful) section = 2;
            else if (full_batch || (hungry && counts != 0u)) section = 3; // a full batch, or TRAV ran dry (batch size checked below)
            else if (n_node >= P.ws_node_min) section = 0;
            else if (n_leaf >= P.ws_leaf_min) section = 1;
            else section = 4; // nothing runs well: take what occupies the most lanes
            if (section >= 3) {
                const uint32_t c0 = counts & 0xFFu, c1 = (counts >> 8) & 0xFFu, c2 = (counts >> 16) & 0xFFu, c3 = counts >> 24;
                const uint32_t m01 = max(c0, c1), m23 = max(c2, c3), best = max(m01, m23);
                k = (m01 >= m23) ? ((c0 >= c1) ? 0u : 1u) : ((c2 >= c3) ? 2u : 3u);
                if (section == 3 && best < 32u && best < P.ws_batch_min) // dry, but the fullest batch is too small to be worth it
                    section = (n_node >= P.ws_node_min) ? 0 : ((n_leaf >= P.ws_leaf_min) ? 1 : 4);
                if (section == 4) {
                    const uint32_t x_use = x_useful ? max(n_done, min(n_idle, trav_count)) : 0u;
                    const uint32_t top = max(max(n_node, n_leaf), max(x_use, best));
                    if (top == 0u) break; // all rings empty, every lane free: the queue is exhausted and all paths ended
                    section = (top == best) ? 3 : ((top == n_node) ? 0 : ((top == n_leaf) ? 1 : 2));
                }
            }
            if (section == 0) burst_floor = max(min(P.ws_node_min, n_node), (n_node * P.ws_burst_num) >> 2);
        }

        bool need_pop = false;
        if (section == 0) {
            // ================= N: one node step (bvh.zig:187-205 as an ordered stack traversal, see closest_bvh) =================
            ZRT_PROF(2, st == BS_NODE);
            if (st == BS_NODE) {
                const float4 *q = reinterpret_cast<const float4 *>(P.nodes + cur);
                const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2);
                const uint2 q3 = __ldg(reinterpret_cast<const uint2 *>(q + 3));
                const SlabHit sh = slab2(q0, q1, q2, o, inv, h.t * 1.00001f);
                if (sh.hl && sh.hr) {
                    const bool left_first = sh.tl <= sh.tr;
                    stack[sp] = make_uint2(left_first ? q3.y : q3.x, __float_as_uint(left_first ? sh.tr : sh.tl));
                    sp++;
                    cur = left_first ? q3.x : q3.y;
                } else if (sh.hl) {
                    cur = q3.x;
                } else if (sh.hr) {
                    cur = q3.y;
                } else {
                    need_pop = true;
                }
                if (!need_pop && (cur & REF_LEAF)) st = BS_LEAF;
            }
        } else if (section == 1) {
            // ================= L: one postponed leaf test (sphere.zig:31-71 / triangle.zig:48-70) =================
            ZRT_PROF(3, st == BS_LEAF);
            if (st == BS_LEAF) {
                leaf_test<false>(P, cur, o, d, h);
                need_pop = true;
            }
        } else if (section == 2) {
            // ================= X: hand finished queries over, re-arm free lanes =================
            uint32_t nk = PK_IDLE;
            ZRT_PROF(13, true);
            ZRT_PROF(4, st == BS_DONE);
            if (st == BS_DONE) {
                if (h.ref == REF_EMPTY) { // raytrace.zig:82-86: the path ends on the background
                    S.meta[slot] |= PM_BG;
                    nk = PK_REGEN;
                } else {
                    const V3 loc = o + d * h.t; // ray.zig:14-16
                    S.ox[slot] = loc.x; S.oy[slot] = loc.y; S.oz[slot] = loc.z;
                    S.href[slot] = h.ref;
                    S.hu[slot] = h.u; S.hv[slot] = h.v;
                    const uint32_t idx = h.ref & REF_INDEX_MASK;
                    const uint32_t mat = (h.ref & REF_SPHERE) ? __ldg(&P.spheres[idx].material) : __ldg(&P.triMeta[idx].material);
                    nk = PK_LAMB + ((mat >> MAT_KIND_SHIFT) & 3u);
                }
                st = BS_FREE;
            }
            { // push onto the shade rings: lanes of a kind find each other with one MATCH (as K1q)
                const uint32_t grp = __match_any_sync(0xffffffffu, nk);
                const uint32_t rank = __popc(grp & lane_lt);
                const uint32_t tails = heads + counts; // bytewise, no carries: head < RING <= 128, count <= N <= 128
                uint32_t add = 0;
                if (nk != PK_IDLE) {
                    S.ring[nk][(((tails >> (8 * nk)) & 0xFFu) + rank) & RM] = (uint8_t)slot;
                    if (rank == 0) add = (uint32_t)__popc(grp) << (8 * nk);
                }
                counts += __reduce_add_sync(0xffffffffu, add);
            }
            { // re-arm: free lanes take the oldest rays of the TRAV ring, lowest lane first
                const uint32_t want = __ballot_sync(0xffffffffu, st == BS_FREE);
                const uint32_t cnt = min((uint32_t)__popc(want), trav_count);
                const uint32_t rank = __popc(want & lane_lt);
                ZRT_PROF(5, st == BS_FREE && rank < cnt);
                if (st == BS_FREE && rank < cnt) {
                    slot = S.ring[BP_TRAV_RING][(trav_head + rank) & RM];
                    o = mk(S.ox[slot], S.oy[slot], S.oz[slot]);
                    d = mk(S.dx[slot], S.dy[slot], S.dz[slot]);
                    inv.x = rcp_approx(d.x);
                    inv.y = rcp_approx(d.y);
                    inv.z = rcp_approx(d.z);
                    h.t = F_INF; h.ref = REF_EMPTY; h.slot = 0xFFFFFFFFu; h.u = h.v = 0.0f;
                    sp = 0;
                    cur = P.root;
                    st = (cur == REF_EMPTY) ? BS_DONE : ((cur & REF_LEAF) ? BS_LEAF : BS_NODE);
                }
                trav_head = (trav_head + cnt) & RM;
                trav_count -= cnt;
            }
            __syncwarp();
        } else {
            // ================= S: shade a batch of kind k =================
            const uint32_t cnt_k = (counts >> (8 * k)) & 0xFFu;
            const uint32_t m = min(cnt_k, 32u);
            const bool active = lane < m;
            ZRT_PROF(6 + (int)k, active);
            const uint32_t head = (heads >> (8 * k)) & 0xFFu;
            const uint32_t ss = S.ring[k][(head + lane) & RM];
            heads = (heads & ~(0xFFu << (8 * k))) | (((head + m) & RM) << (8 * k));
            counts -= m << (8 * k);

            uint32_t meta = 0;
            V3 x = mk(0, 0, 1), nrm = mk(0, 0, 0);
            bool alive = false, to_trav = false, to_regen = false;
            if (k == PK_REGEN) { // warp-uniform
                // ---- the path ended (background: raytrace.zig:82-86, or absorbed / depth limit: black); next sample ----
                uint32_t pxy = 0;
                if (active) {
                    meta = S.meta[ss];
                    pxy = S.pxy[ss];
                    if (meta & PM_BG) { // backgroundColor raytrace.zig:53-58 on the re-normalised direction (:54)
                        const float udy = unit_y(mk(S.dx[ss], S.dy[ss], S.dz[ss]));
                        n_bg++;
                        const float t = 0.5f * (udy + 1.0f);
                        const float it = 1.0f - t;
                        S.ar[ss] += S.tr[ss] * (it + 0.5f * t);
                        S.ag[ss] += S.tg[ss] * (it + 0.7f * t);
                        S.ab[ss] += S.tb[ss] * (it + 1.0f * t);
                    }
                    const uint32_t nsamp = meta & PM_NSAMP_MASK;
                    if ((meta & PM_ITEM) && nsamp >= P.s_end) { // the item hands its sum over (raytrace.zig:180-182)
                        const uint32_t l = (nsamp - P.s_begin) & (L - 1u);
                        const uint32_t pixel = (pxy >> 16) * P.width + (pxy & 0xFFFFu);
                        float *out = P.out + ((size_t)l * P.width * P.height + pixel) * 3;
                        const float sc = (L == 1u) ? P.color_scale : 1.0f;
                        out[0] = S.ar[ss] * sc; out[1] = S.ag[ss] * sc; out[2] = S.ab[ss] * sc;
                        S.ar[ss] = S.ag[ss] = S.ab[ss] = 0.0f;
                        meta &= ~PM_ITEM;
                    }
                }
                const uint32_t g = iq.take(P, total_items, __ballot_sync(0xffffffffu, active && !(meta & PM_ITEM)), lane, lane_lt);
                if (g != ITEM_NONE) {
                    uint32_t l, px, py;
                    item_decode(P, g, l, px, py);
                    pxy = px | (py << 16);
                    S.pxy[ss] = pxy;
                    meta = PM_ITEM | (P.s_begin + l);
                }
                if (active) {
                    if (meta & PM_ITEM) { // raytrace.zig:170-176
                        const uint32_t nsamp = meta & PM_NSAMP_MASK;
                        const uint32_t px = pxy & 0xFFFFu, py = pxy >> 16;
                        const U4 r = rng_ctr(py * P.width + px, nsamp, 0u, P.seed32);
                        x = primary_direction_raw(P, px, py, u01(r.x), u01(r.y));
                        S.tr[ss] = S.tg[ss] = S.tb[ss] = 1.0f;
                        S.ox[ss] = P.ox; S.oy[ss] = P.oy; S.oz[ss] = P.oz;
                        meta = PM_ITEM | (nsamp + L); // bounce 0: the bookkeeping below counts no reflection for this ray
                        alive = true;
                    } else {
                        S.meta[ss] = 0; // the queue is exhausted: this slot leaves the rings for good
                    }
                }
            } else if (active) {
                // ---- a hit: hit record + scatter of kind k (material.zig:43-51) ----
                meta = S.meta[ss];
                const uint32_t pxy = S.pxy[ss];
                const uint32_t pixel = (pxy >> 16) * P.width + (pxy & 0xFFFFu);
                const V3 dd = mk(S.dx[ss], S.dy[ss], S.dz[ss]);
                const uint32_t bounce = (meta >> PM_BOUNCE_SHIFT) & PM_BOUNCE_MASK;
                const uint32_t cur_sample = (meta & PM_NSAMP_MASK) - L;
                Surf s;
                hit_record_at(P, mk(S.ox[ss], S.oy[ss], S.oz[ss]), dd, S.href[ss], S.hu[ss], S.hv[ss], s);
                const DevMaterial *mp = P.mats + (s.material & MAT_INDEX_MASK);
                if (k == PK_LAMB) {
                    x = scatter_lambertian(s.normal, rng_ctr(pixel, cur_sample, bounce, P.seed32));
                } else if (k == PK_METAL) {
                    x = scatter_mirror(unit(dd), s.normal); // material.zig:88
                    nrm = s.normal;
                } else {
                    const U4 r = rng_ctr(pixel, cur_sample, bounce, P.seed32);
                    x = scatter_dielectric(mp, s.front, unit(dd), s.normal, r.x);
                }
                if (k != PK_GLASS) { // attenuation = texture albedo; white for glass
                    const V3 a = albedo(mp, (s.material & MAT_IMAGE_BIT) != 0, s.tu, s.tv);
                    S.tr[ss] *= a.x; S.tg[ss] *= a.y; S.tb[ss] *= a.z;
                }
                meta += 1u << PM_BOUNCE_SHIFT; // provisional: the scatter counts unless the metal absorbs it (below)
                alive = true;
            }
            if (alive) {
                // ---- Ray.init normalises (ray.zig:11-13); bookkeeping of the scatter that produced this ray ----
                const V3 dn = unit(x);
                const bool absorbed = k == PK_METAL && !(dot(dn, nrm) > 0.0f); // material.zig:90-95: black, no reflection counted
                const uint32_t bounce = (meta >> PM_BOUNCE_SHIFT) & PM_BOUNCE_MASK; // index of the ray about to be cast (K1's bounce)
                const uint32_t ok = (k != PK_REGEN && !absorbed) ? 1u : 0u;
                n_refl += ok; // raytrace.zig:95
                const bool exhausted = ok && bounce == P.max_depth + 1u; // the next rayColor call returns black (:64-68)
                n_depth += exhausted ? 1u : 0u;
                meta &= ~PM_BG;
                if (k == PK_REGEN) meta += 1u << PM_BOUNCE_SHIFT; // the primary ray is ray 1
                S.meta[ss] = meta;
                if (absorbed || exhausted) {
                    to_regen = true; // black: nothing to add, the item's next sample starts in a REGEN batch
                } else {
                    S.dx[ss] = dn.x; S.dy[ss] = dn.y; S.dz[ss] = dn.z; // the origin is in place: camera (REGEN) or hit location
                    to_trav = true;
                }
            }
            { // push: new rays onto TRAV, ended paths onto REGEN
                const uint32_t mt = __ballot_sync(0xffffffffu, to_trav), mr = __ballot_sync(0xffffffffu, to_regen);
                if (to_trav) S.ring[BP_TRAV_RING][(trav_head + trav_count + __popc(mt & lane_lt)) & RM] = (uint8_t)ss;
                trav_count += __popc(mt);
                if (to_regen) S.ring[PK_REGEN][((heads & 0xFFu) + (counts & 0xFFu) + __popc(mr & lane_lt)) & RM] = (uint8_t)ss;
                counts += (uint32_t)__popc(mr); // PK_REGEN is byte 0
            }
            __syncwarp(); // slot state and ring entries written by one lane are read by another in a later section
        }
        ZRT_PROF(12, need_pop);
        if (need_pop) { // skip subtrees that fell behind the closest hit found since they were pushed
            st = BS_DONE;
            while (sp > 0) {
                sp--;
                const uint2 e = stack[sp];
                if (__uint_as_float(e.y) <= h.t * 1.00001f) {
                    cur = e.x;
                    st = (cur & REF_LEAF) ? BS_LEAF : BS_NODE;
                    break;
                }
            }
        }
    }

    n_depth = __reduce_add_sync(0xffffffffu, n_depth);
    n_refl = __reduce_add_sync(0xffffffffu, n_refl);
    n_bg = __reduce_add_sync(0xffffffffu, n_bg);
    if (lane == 0) {
        if (n_depth) atomicAdd(P.counters + 0, (unsigned long long)n_depth);
        if (n_refl) atomicAdd(P.counters + 1, (unsigned long long)n_refl);
        if (n_bg) atomicAdd(P.counters + 2, (unsigned long long)n_bg);
    }
}
