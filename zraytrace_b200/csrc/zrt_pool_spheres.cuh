// zrt_pool_spheres.cuh — K1q, third cut (included by zrt_kernels.cu inside namespace zrt).
//
// Same machine as k_trace_pool above (a pool of N work items per warp in shared memory, one ring of slot ids per shading
// kind, batches of up to 32 slots of ONE kind per iteration), rebuilt around what its ncu profile showed
// (profiles/r2_a_pool128_*): 27.3 active lanes per instruction, but 465 thread instructions per ray against 349 of K1 -
// the pool's own bookkeeping cost what the convergence saved.  Changes:
//   * slot state in 16-byte groups (A = direction + meta, B = hit location + pixel, C = throughput + acc.r, D = acc.g, acc.b):
//     one LDS.128 / STS.128 per group instead of a 32-bit access per word;
//   * the kind-specific front half of an iteration is compiled once per kind (switch on the warp-uniform kind): ring
//     bytes at constant positions, no run-time kind tests, dead state of other kinds removed;
//   * ring bytes read with PRMT (__byte_perm), the scheduler takes the first ring that can fill a warp and looks for
//     the fullest one only when none can.
//   * an item that has run out of samples does not hand its sum over inside the REGEN batch that notices (a ~90-instruction
//     path a whole warp would run for typically one lane, 31 samples per item at 32 slices): the slot goes onto a ring of its
//     own (PK_HAND) and 32 of them hand over, take the next 32 items of the global queue - the 32 slices of one pixel - and
//     start their first sample together.
// Arithmetic, RNG keys and per-item summation order are those of K1: images and counters are bit-identical.
#pragma once

template <int N>
struct alignas(16) PoolSlots3 {
    float4 A[N]; // unit direction of the ray that was cast (x, y, z), meta word
    float4 B[N]; // pending hit: location (ray.zig:14-16), px | py << 16
    float4 C[N]; // throughput (r, g, b), the item's f32 sum .r (raytrace.zig:156,177)
    float2 D[N]; // the item's f32 sum .g, .b
    uint8_t ring[PK_COUNT][128];
};

DI uint32_t ring_byte(uint32_t w0, uint32_t w1, uint32_t k) { return __byte_perm(w0, w1, k) & 0xFFu; } // byte k of (w1:w0)

// K1q's state shared by the halves of an iteration (all warp-uniform except the per-lane members)
struct Pool3Rings {
    uint32_t h0 = 0, c0 = 0, h1 = 0, c1 = 0; // heads / counts, one byte per ring: rings 0-3 in word 0, 4-6 in word 1
    template <uint32_t K>
    DI uint32_t head() const { return ((K < 4u ? h0 : h1) >> (8u * (K & 3u))) & 0xFFu; }
    template <uint32_t K>
    DI uint32_t count() const { return ((K < 4u ? c0 : c1) >> (8u * (K & 3u))) & 0xFFu; }
    template <uint32_t K>
    DI void pop(uint32_t m) { // m <= count<K>(); heads stay below 128: 0x7F per byte
        if (K < 4u) { h0 = (h0 + (m << (8u * (K & 3u)))) & 0x7F7F7F7Fu; c0 -= m << (8u * (K & 3u)); }
        else { h1 = (h1 + (m << (8u * (K & 3u)))) & 0x7F7F7F7Fu; c1 -= m << (8u * (K & 3u)); }
    }
    // the first ring that fills a warp; the fullest one when none does.  best = its count (0: every ring is empty)
    DI uint32_t choose(uint32_t &best) const {
        const uint32_t f0 = c0 & 0xE0E0E0E0u, f1 = c1 & 0x00E0E0E0u; // bytes >= 32
        if (f0 | f1) {
            best = 32u;
            return f0 ? (uint32_t)(__ffs((int)f0) - 1) >> 3 : 4u + ((uint32_t)(__ffs((int)f1) - 1) >> 3);
        }
        const uint32_t a0 = c0 & 0xFFu, a1 = (c0 >> 8) & 0xFFu, a2 = (c0 >> 16) & 0xFFu, a3 = c0 >> 24;
        const uint32_t a4 = c1 & 0xFFu, a5 = (c1 >> 8) & 0xFFu, a6 = (c1 >> 16) & 0xFFu;
        const uint32_t m01 = max(a0, a1), m23 = max(a2, a3), m45 = max(a4, a5);
        best = max(max(m01, m23), max(m45, a6));
        if (m01 == best) return (a0 >= a1) ? 0u : 1u;
        if (m23 == best) return (a2 >= a3) ? 2u : 3u;
        if (m45 == best) return (a4 >= a5) ? 4u : 5u;
        return 6u;
    }
};

// what the front half of an iteration hands to the common back half
struct Pool3Lane {
    uint32_t slot, meta, pxy;
    V3 o, x, nrm;
    bool alive;
    uint32_t park; // a lane that is not alive: PK_IDLE (its slot leaves the rings)
};

// backgroundColor (raytrace.zig:53-58) on the re-normalised direction (:54) of a path that left the scene, added to the item's
// sum (raytrace.zig:177): acc (r in c.w, g and b in S.D) += throughput * colour
template <int N>
DI void pool3_add_background(PoolSlots3<N> &S, uint32_t slot, const float4 &a, float4 &c, uint32_t &n_bg) {
    const float2 dd = S.D[slot];
    const float udy = unit_y(mk(a.x, a.y, a.z));
    n_bg++;
    const float t = 0.5f * (udy + 1.0f);
    const float it = 1.0f - t;
    c.w += c.x * (it + 0.5f * t);
    S.D[slot] = make_float2(dd.x + c.y * (it + 0.7f * t), dd.y + c.z * (it + 1.0f * t));
}

// ---- front half, kind REGEN: the path ended (background: raytrace.zig:82-86, or absorbed / depth limit: black) and the
//      item has a sample left (the back half sends the others straight to PK_HAND): the next sample starts (:170-176) ----
template <int N>
DI void pool3_front_regen(const KParams &P, PoolSlots3<N> &S, Pool3Rings &R, uint32_t best, uint32_t lane, uint32_t &n_bg, Pool3Lane &ln) {
    const uint32_t L = P.lanes;
    const uint32_t m = min(best, 32u);
    const bool active = lane < m;
    ZRT_PROF(41, active);
    const uint32_t slot = S.ring[PK_REGEN][(R.head<PK_REGEN>() + lane) & 127u];
    R.pop<PK_REGEN>(m);
    ln.slot = slot;
    ln.alive = false;
    ln.park = PK_IDLE;
    ln.nrm = mk(0, 0, 0);
    ln.o = mk(P.ox, P.oy, P.oz);
    ln.x = mk(0, 0, 1);
    ln.meta = 0;
    ln.pxy = 0;
    if (!active) return;
    const float4 a = S.A[slot];
    float4 c = S.C[slot];
    const uint32_t meta = __float_as_uint(a.w);
    const uint32_t pxy = __float_as_uint(S.B[slot].w);
    if (meta & PM_BG) pool3_add_background<N>(S, slot, a, c, n_bg);
    const uint32_t nsamp = meta & PM_NSAMP_MASK;
    ln.pxy = pxy;
    // raytrace.zig:170-176
    const uint32_t px = pxy & 0xFFFFu, py = pxy >> 16;
    const U4 r = rng_ctr(py * P.width + px, nsamp, 0u, P.seed32);
    ln.x = primary_direction_raw(P, px, py, u01(r.x), u01(r.y));
    S.C[slot] = make_float4(1.0f, 1.0f, 1.0f, c.w);
    ln.meta = PM_ITEM | (nsamp + L); // bounce 0: the bookkeeping of the back half counts no reflection for this ray
    ln.alive = true;
}

// ---- front half, kind HAND: the item hands its sum over (raytrace.zig:180-182), the slot takes a new item from the global
//      queue and starts its first sample.  A full batch draws ONE window of 32 consecutive items: the 32 slices of a pixel ----
template <int N>
DI void pool3_front_hand(const KParams &P, PoolSlots3<N> &S, Pool3Rings &R, ItemQueue &iq, uint32_t total_items, uint32_t best,
                         uint32_t lane, uint32_t lane_lt, uint32_t &n_bg, Pool3Lane &ln) {
    const uint32_t L = P.lanes;
    const uint32_t m = min(best, 32u);
    const bool active = lane < m;
    ZRT_PROF(46, active);
    const uint32_t slot = S.ring[PK_HAND][(R.head<PK_HAND>() + lane) & 127u];
    R.pop<PK_HAND>(m);
    ln.slot = slot;
    ln.alive = false;
    ln.park = PK_IDLE;
    ln.nrm = mk(0, 0, 0);
    ln.o = mk(P.ox, P.oy, P.oz);
    ln.x = mk(0, 0, 1);
    ln.meta = 0;
    ln.pxy = 0;
    if (active && (__float_as_uint(S.A[slot].w) & PM_ITEM)) { // every slot starts here without an item
        const float4 a = S.A[slot];
        float4 c = S.C[slot];
        const uint32_t meta = __float_as_uint(a.w), pxy = __float_as_uint(S.B[slot].w);
        if (meta & PM_BG) pool3_add_background<N>(S, slot, a, c, n_bg); // the item's last path left the scene
        const float2 dd = S.D[slot];
        const uint32_t l = ((meta & PM_NSAMP_MASK) - P.s_begin) & (L - 1u);
        const uint32_t pixel = (pxy >> 16) * P.width + (pxy & 0xFFFFu);
        float *out = P.out + ((size_t)l * P.width * P.height + pixel) * 3;
        const float sc = (L == 1u) ? P.color_scale : 1.0f;
        out[0] = c.w * sc; out[1] = dd.x * sc; out[2] = dd.y * sc;
    }
    const uint32_t g = iq.take(P, total_items, __ballot_sync(0xffffffffu, active), lane, lane_lt);
    if (g != ITEM_NONE) {
        uint32_t l, px, py;
        item_decode(P, g, l, px, py);
        const uint32_t pxy = px | (py << 16), nsamp = P.s_begin + l;
        const U4 r = rng_ctr(py * P.width + px, nsamp, 0u, P.seed32); // raytrace.zig:170-176
        ln.x = primary_direction_raw(P, px, py, u01(r.x), u01(r.y));
        S.B[slot].w = __uint_as_float(pxy);
        S.C[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        S.D[slot] = make_float2(0.0f, 0.0f);
        ln.pxy = pxy;
        ln.meta = PM_ITEM | (nsamp + L);
        ln.alive = true;
    } else if (active) {
        S.A[slot].w = __uint_as_float(0u); // the queue is exhausted: this slot leaves the rings for good
    }
}

// ---- front half, a hit of kind K: hit record + scatter (material.zig:43-51, hit_record.zig:28-41, sphere.zig:45-51) ----
template <uint32_t K, int N>
DI void pool3_front_hit(const KParams &P, PoolSlots3<N> &S, Pool3Rings &R, uint32_t best, uint32_t lane, Pool3Lane &ln) {
    constexpr bool LAMB = K == PK_LAMB || K == PK_LAMB_IMG, METAL = K == PK_METAL || K == PK_METAL_IMG, GLASS = K == PK_GLASS;
    constexpr bool IMG = K == PK_LAMB_IMG || K == PK_METAL_IMG;
    const uint32_t L = P.lanes;
    const uint32_t m = min(best, 32u);
    const bool active = lane < m;
    ZRT_PROF(41 + (int)K, active);
    const uint32_t slot = S.ring[K][(R.head<K>() + lane) & 127u];
    R.pop<K>(m);
    ln.slot = slot;
    ln.alive = false;
    ln.park = PK_IDLE;
    ln.nrm = mk(0, 0, 0);
    ln.x = mk(0, 0, 1);
    ln.o = mk(0, 0, 0);
    ln.meta = 0;
    ln.pxy = 0;
    if (!active) return;
    const float4 a = S.A[slot], b = S.B[slot];
    uint32_t meta = __float_as_uint(a.w);
    const uint32_t pxy = __float_as_uint(b.w);
    const uint32_t pixel = (pxy >> 16) * P.width + (pxy & 0xFFFFu);
    const V3 d = mk(a.x, a.y, a.z);
    const uint32_t hi = (meta >> PM_HIT_SHIFT) & 7u;
    const uint32_t bounce = (meta >> PM_BOUNCE_SHIFT) & PM_BOUNCE_MASK;
    const uint32_t cur_sample = (meta & PM_NSAMP_MASK) - L;
    const V3 o = mk(b.x, b.y, b.z);
    const float4 ca = ldg4(reinterpret_cast<const float4 *>(P.spheres + hi));
    const uint4 cb = __ldg(reinterpret_cast<const uint4 *>(P.spheres + hi) + 1);
    const V3 on = (o - mk(ca.x, ca.y, ca.z)) * __uint_as_float(cb.x); // sphere.zig:46 (1.0 / radius precomputed)
    const bool front = !(dot(d, on) > 0.0f);                          // hit_record.zig:29
    const V3 normal = front ? on : neg(on);
    const DevMaterial *mp = P.mats + (cb.y & MAT_INDEX_MASK);
    if (LAMB) {
        ln.x = scatter_lambertian(normal, rng_ctr(pixel, cur_sample, bounce, P.seed32));
    } else if (METAL) {
        ln.x = scatter_mirror(unit(d), normal); // material.zig:88
        ln.nrm = normal;
    } else {
        const U4 r = rng_ctr(pixel, cur_sample, bounce, P.seed32);
        ln.x = scatter_dielectric(mp, front, unit(d), normal, r.x);
    }
    if (!GLASS) { // attenuation = texture albedo; white for glass
        // a sphere whose material samples an image sits in an IMG ring when the host split the rings (P.pool_split), in
        // the plain ring of its kind otherwise: the plain rings still test the bit, the IMG rings know it
        const bool is_image = IMG || (!P.pool_split && (cb.y & MAT_IMAGE_BIT) != 0);
        float tu = 0.0f, tv = 0.0f;
        if (is_image) sphere_uv(P, on, tu, tv); // only image textures ever read (u, v)
        const V3 al = albedo(mp, is_image, tu, tv);
        const float4 c = S.C[slot];
        S.C[slot] = make_float4(c.x * al.x, c.y * al.y, c.z * al.z, c.w);
    }
    ln.o = o;
    ln.meta = meta + (1u << PM_BOUNCE_SHIFT); // provisional: the scatter counts unless the metal absorbs it (back half)
    ln.pxy = pxy;
    ln.alive = true;
}

template <int NS, int N, int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS) k_trace_pool3(const __grid_constant__ KParams P) {
    static_assert(N >= 32 && N <= 128 && (N % 16) == 0, "slots per warp; ring bytes hold counts up to 128");
    __shared__ PoolSlots3<N> pools[4];
    PoolSlots3<N> &S = pools[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t total_items = P.x_end * P.height * P.lanes;
    const uint32_t lane_lt = (1u << lane) - 1u;
    ItemQueue iq;
    uint32_t n_refl = 0, n_bg = 0, n_depth = 0; // pixels, samples and rays: k_finish_counters (see K1)

    Pool3Rings R;
    R.c1 = (uint32_t)N << (8 * (PK_HAND & 3u));
    for (uint32_t s = lane; s < (uint32_t)N; s += 32u) { // every slot starts without an item, waiting for one
        S.ring[PK_HAND][s] = (uint8_t)s;
        S.A[s] = make_float4(0.0f, 0.0f, 1.0f, __uint_as_float(0u));
        S.C[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        S.D[s] = make_float2(0.0f, 0.0f);
    }
    __syncwarp();

    for (;;) {
        ZRT_PROF_TICK();
        ZRT_PROF(40, true);
        uint32_t best;
        const uint32_t k = R.choose(best);
        if (best == 0) break; // every slot is idle: the global queue is exhausted and all paths have ended
        Pool3Lane ln;
        switch (k) { // warp-uniform
        case PK_REGEN: pool3_front_regen<N>(P, S, R, best, lane, n_bg, ln); break;
        case PK_HAND: pool3_front_hand<N>(P, S, R, iq, total_items, best, lane, lane_lt, n_bg, ln); break;
        case PK_LAMB: pool3_front_hit<PK_LAMB, N>(P, S, R, best, lane, ln); break;
        case PK_METAL: pool3_front_hit<PK_METAL, N>(P, S, R, best, lane, ln); break;
        case PK_GLASS: pool3_front_hit<PK_GLASS, N>(P, S, R, best, lane, ln); break;
        case PK_LAMB_IMG: pool3_front_hit<PK_LAMB_IMG, N>(P, S, R, best, lane, ln); break;
        default: pool3_front_hit<PK_METAL_IMG, N>(P, S, R, best, lane, ln); break;
        }
        // ---- back half, common: Ray.init normalises (ray.zig:11-13); bookkeeping of the scatter that produced this ray;
        //      the closest-hit query (raytrace.zig:71-81); classification ----
        uint32_t next_kind = ln.park;
        if (ln.alive) {
            ZRT_PROF(47, true);
            const bool metal = k == PK_METAL || k == PK_METAL_IMG, primary = k == PK_REGEN || k == PK_HAND; // warp-uniform
            uint32_t meta = ln.meta;
            const V3 dn = unit(ln.x);
            const bool absorbed = metal && !(dot(dn, ln.nrm) > 0.0f); // material.zig:90-95: black, no reflection counted
            const uint32_t bounce = (meta >> PM_BOUNCE_SHIFT) & PM_BOUNCE_MASK; // index of the ray about to be cast (K1's bounce)
            const uint32_t ok = (!primary && !absorbed) ? 1u : 0u;
            n_refl += ok; // raytrace.zig:95
            const bool exhausted = ok && bounce == P.max_depth + 1u; // the next rayColor call returns black (:64-68)
            n_depth += exhausted ? 1u : 0u;
            meta &= ~((7u << PM_HIT_SHIFT) | PM_BG);
            if (primary) meta += 1u << PM_BOUNCE_SHIFT; // the primary ray is ray 1
            // a path that ends now: next sample in a REGEN batch, or - no sample left (meta holds the NEXT sample index) - the
            // hand-over ring directly, so that no REGEN lane is spent on noticing it
            next_kind = ((meta & PM_NSAMP_MASK) >= P.s_end) ? PK_HAND : PK_REGEN;
            if (!(absorbed || exhausted)) {
                Hit h;
                ZRT_PROF(primary ? 48 : 49, true);
                if (primary) closest_spheres_primary<NS>(P, dn, h);
                else closest_hit<MODE_SPHERES, NS, false>(P, ln.o, dn, h);
                if (h.ref == REF_EMPTY) {
                    meta |= PM_BG;
                } else {
                    const uint32_t hi = h.ref & 7u;
                    const V3 loc = ln.o + dn * h.t; // ray.zig:14-16
                    S.B[ln.slot] = make_float4(loc.x, loc.y, loc.z, __uint_as_float(ln.pxy));
                    meta |= hi << PM_HIT_SHIFT;
                    next_kind = (P.inl_kinds >> (3u * hi)) & 7u; // the ring of this sphere's material
                }
            }
            S.A[ln.slot] = make_float4(dn.x, dn.y, dn.z, __uint_as_float(meta));
        }
        // ---- push every slot of the batch onto the ring of its next kind: lanes of a kind find each other with one
        //      MATCH, the group's first lane reports its size ----
        {
            const uint32_t grp = __match_any_sync(0xffffffffu, next_kind);
            const uint32_t rank = __popc(grp & lane_lt);
            uint32_t add0 = 0, add1 = 0;
            if (next_kind != PK_IDLE) {
                const uint32_t tail = ring_byte(R.h0 + R.c0, R.h1 + R.c1, next_kind); // bytewise sums: no carries (head < 128, count <= 128)
                S.ring[next_kind][(tail + rank) & 127u] = (uint8_t)ln.slot;
                if (rank == 0) {
                    const uint32_t a = (uint32_t)__popc(grp) << (8u * (next_kind & 3u));
                    if (next_kind < 4u) add0 = a; else add1 = a;
                }
            }
            R.c0 += __reduce_add_sync(0xffffffffu, add0);
            R.c1 += __reduce_add_sync(0xffffffffu, add1);
        }
        __syncwarp(); // slot state and ring entries written by one lane are read by another in the next iteration
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
