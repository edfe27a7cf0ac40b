// mesh.cu -- the triangle-scene path: primary pass (ray-gen + traversal + G-buffer) and the
// wavefront path tracer (north_star rows n3-n7):
//     per sample:  shade(primary hits) -> [ trace(queue) -> shade(queue) ] x bounces
// trace = persistent-warp traversal (trace.cuh).  shade = material + blue-noise-rotated PCG sampling +
// sky on miss + accumulation; it emits the next bounce's rays into a compacted queue with one
// warp-aggregated atomic per warp (row n5).
// Option "fused_shade": the shade stage of a bounce wave runs inside the traversal kernel (k_trace_shade): a
// persistent warp that has traced its last ray and found the queue empty turns into a shading worker for the
// wave, so that shading fills the SMs which the kernel's drain phase (the longest rays finishing, ~25 % of the
// launch) leaves idle.  The 8-byte hit record itself is the ready flag (sentinel until the ray is traced).
// Measured: warps only become free late in the drain, so the overlap is small (+1.5 % at 1080p 1 spp, -2 % at
// 4K 8 spp where a full-GPU shade kernel is more efficient) -- off by default.
// Per-pixel semantics are those of secondaryRays.comp:64-135 with the primary hit carried in fp32:
// the PCG state threads through all samples of a pixel, so samples run sequentially and the
// wavefront is over pixels.  Every pixel owns at most one live path => accumulation needs no atomics
// and the image is independent of queue order (deterministic).
#include "shading.cuh"
#include "trace.cuh"

// The two halves of the visit counters (primary / trace_rays: [0..3], bounce waves: [4..7]) are reset by their own
// passes, and mrt_stats adds them up: a fresh allocation must start at zero or the half that has not run yet
// reports garbage (seen as 4 phantom stack overflows in a context that had only called mrt_trace_rays).
static int reserve_visit_counters(mrt_context* ctx) {
    if (ctx->visit_counters.p) return MRT_OK;
    MRT_TRY(dev_reserve(ctx, ctx->visit_counters, 8));
    return mrt_check_cuda(ctx, cudaMemsetAsync(ctx->visit_counters.p, 0, 8 * sizeof(unsigned long long), ctx->stream), "visit counters");
}

namespace {

struct MeshFrame {
    RayGen gen;
    Mat4 PV, PVprev;
    Partition part;
    uint32_t local_rows;
};

MRT_D void flush_counters(const TraceCounters& c, unsigned long long* counters, bool count_visits) {
    if (c.overflow) atomicAdd(&counters[2], (unsigned long long)c.overflow);
    if (count_visits) {
        unsigned n = c.nodes, t = c.tris;
        for (int off = 16; off > 0; off >>= 1) {
            n += __shfl_down_sync(0xFFFFFFFFu, n, off);
            t += __shfl_down_sync(0xFFFFFFFFu, t, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&counters[0], (unsigned long long)n);
            atomicAdd(&counters[1], (unsigned long long)t);
        }
    }
}

// ---- primary pass: primaryRay.comp:38-76 with the 5-sphere loop replaced by BVH traversal ----
// Ray index r -> 8x4-pixel tile r/32 (row-major tile grid), pixel r%32 inside it, so the 32 rays a fresh
// warp takes are one compact tile (coherent) and padding pixels are skipped by load().
#ifndef PRIMARY_TILE_W
#define PRIMARY_TILE_W 8u   // pixels per tile row of a warp's 32-pixel tile (8 x 4; A/B: 4 x 8, 16 x 2, 32 x 1)
#endif
struct PrimaryJob {
    static constexpr bool ANY_HIT = false;
    MeshFrame F;
    BvhDev bvh;
    uint32_t tiles_x, tiles;
    uint32_t* vis;
    uint16_t *depth, *normal, *motion;
    float* hit_t;
    float4 *hit0_pos, *hit0_n;

    MRT_D uint32_t count() const { return tiles * 32u; }
    MRT_D bool pixel(uint32_t i, uint32_t& x, uint32_t& lr) const {
        uint32_t tile = i >> 5, in = i & 31u;
        x = (tile % tiles_x) * PRIMARY_TILE_W + (in % PRIMARY_TILE_W);
        lr = (tile / tiles_x) * (32u / PRIMARY_TILE_W) + (in / PRIMARY_TILE_W);
        return x < F.gen.W && lr < F.local_rows;
    }
    MRT_D bool load(uint32_t i, float3& o, float3& d) const {
        uint32_t x, lr;
        if (!pixel(i, x, lr)) return false;
        ray_gen(F.gen, x, partition_local_to_y(F.part, lr), o, d);
        return true;
    }
    MRT_D bool load_prepared(uint32_t, float3&, RayPre&) const { return false; }  // rays are set up by the traversal kernel
    MRT_D void store(uint32_t i, const TraceHit& h) const {
        uint32_t x, lr;
        pixel(i, x, lr);
        float3 o, d;
        ray_gen(F.gen, x, partition_local_to_y(F.part, lr), o, d);  // cheaper than carrying d through traversal
        store_with_ray(i, h, o, d);
    }
    MRT_D void store_with_ray(uint32_t i, const TraceHit& h, float3 o, float3 d) const {
        uint32_t x, lr;
        pixel(i, x, lr);
        store_xy(x, lr, h, o, d);
    }
    MRT_D void store_xy(uint32_t x, uint32_t lr, const TraceHit& h, float3 o, float3 d) const {
        const size_t p = (size_t)lr * F.gen.W + x;
        float dep = 0.0f;
        float2 mo = make_float2(0.0f, 0.0f);
        float3 n = d, pos = f3s(0.0f);
        if (h.prim != MRT_MISS_ID) {
            pos = o + d * h.t;
            n = tri_facing_normal(bvh, h.tri, d, nullptr);
            project_hit(F.PV, F.PVprev, pos, F.gen.W, F.gen.H, dep, mo);
        }
        store_gbuffer(vis, depth, normal, motion, p, h.prim, dep, n, mo);
        hit_t[p] = h.prim != MRT_MISS_ID ? h.t : 0.0f;
        hit0_pos[p] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(h.prim));
        hit0_n[p] = make_float4(n.x, n.y, n.z, 0.0f);
    }
};

// ---- tile-frustum entry (primary rays) ----
// The rays of a screen tile share their origin and differ by a fraction of a degree, yet each of them used to repeat
// the child tests of the upper BVH levels for itself (ncu r1: 268 M warp instructions for 2 M rays).  Now a CTA owns a
// 32x16-pixel tile and
//   A. walks the top of the tree ONCE against the tile's frustum, all 128 threads together (8 threads test the 8
//      children of a node, 16 nodes per round), down to the nodes that are no larger than the footprint of an 8x4 SUB-tile
//      at their distance; the result is an ENTRY LIST in shared memory -- node entries (node index + fp32 box relative to
//      the camera) and leaf entries (the triangles of a leaf slot of a descended node + its box) -- sorted near to far;
//   B. each warp then takes four of the tile's sixteen 8x4 sub-tiles in turn: it culls the list against the sub-tile's own
//      frustum (one entry per lane, a ballot per 32 entries), and every ray tests itself against the surviving boxes
//      (12 instructions each, no divergence), tests the triangles of the leaf entries it touches, pushes the node
//      entries it touches onto its own stack, far to near, and continues with the ordinary per-ray walk.
// A first version that built the list per warp (8x4 tile) removed 70 % of the primary node steps and gained nothing:
// the descent and the sort cost every lane as much as the steps they saved.  Amortised over 512 rays they cost ~3 %.
// Exactness: the frustum tests and the entry-box test only ever REJECT boxes that no ray of the tile / the ray cannot
// touch (planes through the corner rays of the tile, slack for rounding; boxes decoded from the quantised planes and
// widened by 2 ulp), so the set of triangles a ray tests still contains its closest hit, and closest hit = lexicographic
// minimum of (t, primitive id) does not depend on the order: ids and t stay bit-identical to brute force (tested).
// Tiles whose corner rays do not share one direction octant (they straddle an axis plane through the camera), trees of
// a single node and tiles whose list would overflow fall back to the walk from the root.
#ifndef PRIMARY_ENTRY
#define PRIMARY_ENTRY 1
#endif
#ifndef ENTRY_MAX
#define ENTRY_MAX 256      // entries per 32x16 tile (a power of two: bitonic sort)
#endif
#ifndef ENTRY_WORK
#define ENTRY_WORK 256     // nodes per level of the frustum walk
#endif
#ifndef ENTRY_SUB
#define ENTRY_SUB 96       // entries a sub-tile may keep after culling
#endif
#ifndef ENTRY_K
#define ENTRY_K 1.0f       // descend while a node's diagonal exceeds ENTRY_K x the sub-tile's footprint at its distance
#endif
#define ENTRY_NODE 0xFFFFFFFFu
#define BIG_W 32
#define BIG_H 16

struct EntryTile {                   // per CTA
    float rn[3][ENTRY_MAX];          // entry planes relative to the camera: the ones the tile's octant enters through ...
    float rf[3][ENTRY_MAX];          // ... and leaves through
    uint32_t a[ENTRY_MAX];           // node entry: node index; leaf entry: tri_base of the node
    uint32_t b[ENTRY_MAX];           // node entry: ENTRY_NODE; leaf entry: the slot's triangle bits
    uint32_t c[ENTRY_MAX];           // leaf entry: leafmask24 of the node
    unsigned long long key[ENTRY_MAX];  // sort key: ordered distance bits << 32 | entry index
    uint32_t work[2][ENTRY_WORK];
    unsigned char sub[TRACE_BLOCK / 32][ENTRY_SUB];  // per warp: the entries that survive its sub-tile's frustum
    int nwork, nnext, nent, overflow;
};

struct Frustum { float3 pn[4]; float3 axis; float angle; };

// frustum through four corner directions (any order around the tile)
MRT_D Frustum make_frustum(float3 c00, float3 c10, float3 c01, float3 c11) {
    Frustum F;
    F.axis = normalize3(c00 + c10 + c01 + c11);
    F.pn[0] = cross3(c00, c10); F.pn[1] = cross3(c10, c11); F.pn[2] = cross3(c11, c01); F.pn[3] = cross3(c01, c00);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (dot3(F.pn[k], F.axis) < 0.0f) F.pn[k] = -F.pn[k];
        F.pn[k] = normalize3(F.pn[k]);
    }
    F.angle = length3(c11 - c00);
    return F;
}

// false: no point of the camera-relative box [lo, hi] lies inside the frustum (conservative: slack for rounding)
MRT_D bool frustum_touches(const Frustum& F, float3 lo, float3 hi, float& dmin) {
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
    const float slack = 1e-5f * mag;
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float3 n = F.pn[k];
        const float s = n.x * (n.x > 0.0f ? hi.x : lo.x) + n.y * (n.y > 0.0f ? hi.y : lo.y) + n.z * (n.z > 0.0f ? hi.z : lo.z);
        inside = inside && s >= -slack;
    }
    const float3 a = F.axis;
    const float dmax = a.x * (a.x > 0.0f ? hi.x : lo.x) + a.y * (a.y > 0.0f ? hi.y : lo.y) + a.z * (a.z > 0.0f ? hi.z : lo.z);
    dmin = a.x * (a.x > 0.0f ? lo.x : hi.x) + a.y * (a.y > 0.0f ? lo.y : hi.y) + a.z * (a.z > 0.0f ? lo.z : hi.z);
    return inside && dmax >= -slack;  // (not entirely behind the camera)
}

MRT_D unsigned long long entry_key(float dist, int e) {
    unsigned u = __float_as_uint(dist);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of the float
    return ((unsigned long long)u << 32) | (unsigned)e;
}

// Stage A, all threads of the CTA: entry list of the tile with frustum F (camera at o).  Returns the number of entries
// (sorted near to far), or -1 for "walk from the root".  sub_angle: angular size of an 8x4 sub-tile.
MRT_D int build_entry_list(const BvhDev& bvh, float3 o, const Frustum& F, float sub_angle, float t_far, unsigned oct_inv, EntryTile& E,
                           TraceCounters& cnt) {
    const bool px = oct_inv & 1u, py = oct_inv & 2u, pz = oct_inv & 4u;
    if (threadIdx.x == 0) { E.work[0][0] = 0u; E.nwork = 1; E.nnext = 0; E.nent = 0; E.overflow = 0; }
    __syncthreads();
    int cur = 0;
    for (;;) {
        const int nwork = E.nwork;
        if (nwork == 0 || E.overflow) break;
        for (int base = 0; base < nwork; base += TRACE_BLOCK / 8) {
            const int wi = base + (int)(threadIdx.x >> 3);
            const unsigned j = threadIdx.x & 7u;
            if (wi >= nwork) continue;
            const uint32_t node = E.work[cur][wi];
            const uint4* np = reinterpret_cast<const uint4*>(bvh.nodes + node);
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (j == 0) cnt.nodes++;
            const float stx = __uint_as_float((n0.w & 0xFFu) << 23), sty = __uint_as_float((n0.w & 0xFF00u) << 15),
                        stz = __uint_as_float((n0.w & 0xFF0000u) << 7);
            const unsigned sh = 8u * (j & 3u);
            const bool hi4 = j >= 4u;
            const float qlx = (float)(((hi4 ? n2.y : n2.x) >> sh) & 0xFFu), qly = (float)(((hi4 ? n2.w : n2.z) >> sh) & 0xFFu),
                        qlz = (float)(((hi4 ? n3.y : n3.x) >> sh) & 0xFFu);
            const float qhx = (float)(((hi4 ? n3.w : n3.z) >> sh) & 0xFFu), qhy = (float)(((hi4 ? n4.y : n4.x) >> sh) & 0xFFu),
                        qhz = (float)(((hi4 ? n4.w : n4.z) >> sh) & 0xFFu);
            if (!(qlx <= qhx && qly <= qhy && qlz <= qhz)) continue;  // (empty slots carry the inverted box)
            const float ox = __uint_as_float(n0.x), oy = __uint_as_float(n0.y), oz = __uint_as_float(n0.z);
            float3 lo = f3(fmaf(qlx, stx, ox), fmaf(qly, sty, oy), fmaf(qlz, stz, oz));
            float3 hi = f3(fmaf(qhx, stx, ox), fmaf(qhy, sty, oy), fmaf(qhz, stz, oz));
            // the planes the traversal sees are origin + q * step evaluated exactly: widen the fp32 values by 2 ulp of the
            // largest magnitude involved, then move to camera-relative coordinates
            const float ex = 2.4e-7f * fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fabsf(o.x)),
                        ey = 2.4e-7f * fmaxf(fmaxf(fabsf(lo.y), fabsf(hi.y)), fabsf(o.y)),
                        ez = 2.4e-7f * fmaxf(fmaxf(fabsf(lo.z), fabsf(hi.z)), fabsf(o.z));
            lo = f3(lo.x - ex, lo.y - ey, lo.z - ez) - o;
            hi = f3(hi.x + ex, hi.y + ey, hi.z + ez) - o;
            float dmin;
            if (!frustum_touches(F, lo, hi, dmin)) continue;
            const unsigned imask = n0.w >> 24;
            const bool is_node = (imask >> j) & 1u;
            const uint32_t leafbits = n1.z & (7u << (3u * j));
            uint32_t ea = n1.y, eb = leafbits, ec = n1.z;
            bool as_entry = !is_node && leafbits != 0u;
            if (is_node) {
                const uint32_t child = n1.x + __popc(imask & ((1u << j) - 1u));
                as_entry = true;
                ea = child; eb = ENTRY_NODE; ec = 0u;
                // still large, and not beyond everything the probe rays hit (occluded parts of the frustum stay coarse:
                // a ray that does get there walks them the ordinary way)
                if (dmin <= t_far && (!(dmin > 0.0f) || length3(hi - lo) > ENTRY_K * sub_angle * dmin)) {
                    const int p = atomicAdd(&E.nnext, 1);
                    if (p < ENTRY_WORK) { E.work[cur ^ 1][p] = child; as_entry = false; }
                }
            }
            if (as_entry) {
                const int e = atomicAdd(&E.nent, 1);
                if (e < ENTRY_MAX) {
                    E.rn[0][e] = px ? lo.x : hi.x; E.rf[0][e] = px ? hi.x : lo.x;
                    E.rn[1][e] = py ? lo.y : hi.y; E.rf[1][e] = py ? hi.y : lo.y;
                    E.rn[2][e] = pz ? lo.z : hi.z; E.rf[2][e] = pz ? hi.z : lo.z;
                    E.a[e] = ea; E.b[e] = eb; E.c[e] = ec;
                    E.key[e] = entry_key(dmin, e);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            E.nwork = min(E.nnext, ENTRY_WORK);
            E.nnext = 0;
            if (E.nent > ENTRY_MAX) E.overflow = 1;
        }
        cur ^= 1;
        __syncthreads();
    }
    const int n = E.nent;
    if (E.overflow || n > ENTRY_MAX) return -1;
    // ---- sort near to far: bitonic sort of (distance, index) keys, then the permutation through registers
    int m = 1;
    while (m < n) m <<= 1;
    for (int e = n + (int)threadIdx.x; e < m; e += TRACE_BLOCK) E.key[e] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int t = threadIdx.x; t < m; t += TRACE_BLOCK) {
                const int p = t ^ jj;
                if (p > t) {
                    const unsigned long long x = E.key[t], y = E.key[p];
                    const bool up = (t & k) == 0;
                    if ((x > y) == up) { E.key[t] = y; E.key[p] = x; }
                }
            }
            __syncthreads();
        }
    float v[2][6];
    uint32_t w[2][3];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = (int)threadIdx.x + TRACE_BLOCK * h;
        if (r < n) {
            const int e = (int)(unsigned)E.key[r];
            v[h][0] = E.rn[0][e]; v[h][1] = E.rn[1][e]; v[h][2] = E.rn[2][e];
            v[h][3] = E.rf[0][e]; v[h][4] = E.rf[1][e]; v[h][5] = E.rf[2][e];
            w[h][0] = E.a[e]; w[h][1] = E.b[e]; w[h][2] = E.c[e];
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = (int)threadIdx.x + TRACE_BLOCK * h;
        if (r < n) {
            E.rn[0][r] = v[h][0]; E.rn[1][r] = v[h][1]; E.rn[2][r] = v[h][2];
            E.rf[0][r] = v[h][3]; E.rf[1][r] = v[h][4]; E.rf[2][r] = v[h][5];
            E.a[r] = w[h][0]; E.b[r] = w[h][1]; E.c[r] = w[h][2];
        }
    }
    __syncthreads();
    return n;
}

// Stage B, one warp: the entries of the tile's list that the sub-tile's frustum touches -> E.sub[warp][0..count).
// Returns the count, or -1 when more than ENTRY_SUB survive (the sub-tile then walks from the root).
MRT_D int cull_entries(const Frustum& F, unsigned oct_inv, EntryTile& E, int n) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool px = oct_inv & 1u, py = oct_inv & 2u, pz = oct_inv & 4u;
    int count = 0;
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        const int e = base + (int)lane;
        bool keep = false;
        if (e < n) {
            const float nx = E.rn[0][e], fx = E.rf[0][e], ny = E.rn[1][e], fy = E.rf[1][e], nz = E.rn[2][e], fz = E.rf[2][e];
            float dmin;
            keep = frustum_touches(F, f3(px ? nx : fx, py ? ny : fy, pz ? nz : fz), f3(px ? fx : nx, py ? fy : ny, pz ? fz : nz), dmin);
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        if (count + __popc(m) > ENTRY_SUB) return -1;
        if (keep) E.sub[warp][count + __popc(m & ((1u << lane) - 1u))] = (unsigned char)e;
        count += __popc(m);
    }
    __syncwarp();
    return count;
}

// closest hit of one ray of a sub-tile, starting from the sub-tile's entries
MRT_D TraceHit trace_from_entries(const BvhDev& bvh, float3 o, float3 d, const EntryTile& E, int count, TraceShared& S,
                                  TraceCounters& cnt) {
    uint2* const sm = &S.stack[0][threadIdx.x];
    const unsigned char* const sub = E.sub[threadIdx.x >> 5];
    uint2 spill[TRACE_LOCAL_STACK];
    LaneState L;
    lane_begin(L, o, d);
    L.ng.y = 0u;
    // node entries the ray touches, as a bit list over the sub-tile's entries (ENTRY_SUB <= 96 bits)
    unsigned h0 = 0u, h1 = 0u, h2 = 0u;
    for (int k = 0; k < count; k++) {
        const int e = sub[k];
        const float t0x = E.rn[0][e] * L.idn.x, t0y = E.rn[1][e] * L.idn.y, t0z = E.rn[2][e] * L.idn.z;
        const float t1x = E.rf[0][e] * L.idf.x, t1y = E.rf[1][e] * L.idf.y, t1z = E.rf[2][e] * L.idf.z;
        const float tmin = fmaxf(fmaxf(t0x, t0y), t0z), tmax = fminf(fminf(t1x, t1y), t1z);
        const bool hit = tmin <= tmax && tmin <= L.tlimit && tmax >= 0.0f;
        const uint32_t kind = E.b[e];
        if (kind == ENTRY_NODE) {
            const unsigned bit = hit ? 1u << (k & 31) : 0u;
            if (k < 32) h0 |= bit; else if (k < 64) h1 |= bit; else h2 |= bit;
        } else if (hit) {
            L.tg = make_uint2(E.a[e], kind);
            L.tgmask = E.c[e];
            while (L.tg.y) lane_tri_step<false>(L, bvh, cnt);
        }
    }
    // onto the stack far to near; the nearest becomes the pending group
    const unsigned one = (1u << (24u + L.oct_inv)) | 1u;  // "group" of the single node a: slot 0, see lane_node_step
    while (h0 | h1 | h2) {
        int k;
        if (h2) { k = 31 - __clz(h2); h2 &= ~(1u << k); k += 64; }
        else if (h1) { k = 31 - __clz(h1); h1 &= ~(1u << k); k += 32; }
        else { k = 31 - __clz(h0); h0 &= ~(1u << k); }
        const uint2 g = make_uint2(E.a[sub[k]], one);
        if (!(h0 | h1 | h2)) { L.ng = g; break; }
        if (L.sp < TRACE_SM_STACK) sm[L.sp * TRACE_BLOCK] = g;
        else if (L.sp < TRACE_SM_STACK + TRACE_LOCAL_STACK) spill[L.sp - TRACE_SM_STACK] = g;
        else cnt.overflow++;
        L.sp = min(L.sp + 1, TRACE_SM_STACK + TRACE_LOCAL_STACK);
    }
    for (;;) {
        if (L.ng.y & 0xFF000000u) lane_node_step<false>(L, bvh, S, spill, cnt);
        while (L.tg.y) lane_tri_step<false>(L, bvh, cnt);
        if (!(L.ng.y & 0xFF000000u)) {
            if (L.sp == 0) break;
            L.sp--;
            L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
        }
    }
    return L.hit;
}

// Primary rays are coherent: one thread per pixel, 8x4-pixel tile per warp, per-lane traversal loop
// (trace_coherent).  Measured 2x faster than running them through the persistent state machine.
#ifndef PRIMARY_MIN_BLOCKS
#define PRIMARY_MIN_BLOCKS 6  // register cap 85 (ptxas picks 80 instead of 72): measured best of 1, 4, 5, 6, 8, 9
#endif
template <bool BATCHED>
__global__ void __launch_bounds__(TRACE_BLOCK, PRIMARY_MIN_BLOCKS)
k_mesh_primary(PrimaryJob J, unsigned long long* counters, int count_visits) {
    __shared__ TraceShared S;
    trace_shared_init(S);
    // CTA = 4 warps = four 8x4 tiles side by side; ray index as in PrimaryJob::pixel
    const uint32_t i = blockIdx.x * TRACE_BLOCK + threadIdx.x;
    TraceCounters cnt{0, 0, 0};
    float3 o = f3s(0.0f), d = f3(1.0f, 0.0f, 0.0f);
    if (BATCHED) {  // option "primary_batched": warp-voted triangle steps (trace_coherent_batched)
        const bool active = i < J.count() && J.load(i, o, d);
        TraceHit h = trace_coherent_batched(J.bvh, active, o, d, S, cnt);
        if (active) J.store_with_ray(i, h, o, d);
    } else if (i < J.count() && J.load(i, o, d)) {
        TraceHit h = trace_coherent(J.bvh, o, d, S, cnt);
        J.store_with_ray(i, h, o, d);
    }
    flush_counters(cnt, counters, count_visits != 0);
}

#if PRIMARY_ENTRY
// Primary pass with the tile-frustum entry list: CTA = one 32x16-pixel tile (grid = tiles_x32 x tiles_y16).
__global__ void __launch_bounds__(TRACE_BLOCK, PRIMARY_MIN_BLOCKS)
k_mesh_primary_entry(PrimaryJob J, uint32_t big_x, unsigned long long* counters, int count_visits) {
    __shared__ TraceShared S;
    __shared__ EntryTile E;
    trace_shared_init(S);
    TraceCounters cnt{0, 0, 0};
    const uint32_t bx = blockIdx.x % big_x, by = blockIdx.x / big_x;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, full = 0xFFFFFFFFu;
    const uint32_t x0 = bx * BIG_W, r0 = by * BIG_H;
    // the tile's corner rays (pixels outside the image still have rays); local rows map monotonically to image rows
    float3 o, c00, c10, c01, c11;
    ray_gen(J.F.gen, x0, partition_local_to_y(J.F.part, r0), o, c00);
    ray_gen(J.F.gen, x0 + BIG_W - 1, partition_local_to_y(J.F.part, r0), o, c10);
    ray_gen(J.F.gen, x0, partition_local_to_y(J.F.part, r0 + BIG_H - 1), o, c01);
    ray_gen(J.F.gen, x0 + BIG_W - 1, partition_local_to_y(J.F.part, r0 + BIG_H - 1), o, c11);
    auto octant = [](float3 d) { return 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u)); };
    const unsigned oct_inv = octant(c00);
    int n = -1;
    const Frustum FB = make_frustum(c00, c10, c01, c11);
    __shared__ float probe_far[TRACE_BLOCK / 32];
    // one octant for the whole tile (the interior rays are convex combinations of the corners), a sane frustum, a real tree
    const bool usable = J.bvh.num_nodes >= 2 && octant(c10) == oct_inv && octant(c01) == oct_inv && octant(c11) == oct_inv &&
                        FB.angle > 0.0f && FB.angle < 0.5f;
    for (int it = 0; it < (BIG_W / 8) * (BIG_H / 4) / (TRACE_BLOCK / 32); it++) {
        // sub-tile of this warp: 4 x 4 grid of 8x4 tiles; the first round takes the diagonal (spread over the tile) and walks
        // from the root: its hit distances bound how deep into the frustum the entry list needs to be fine
        const unsigned sidx = ((warp + (unsigned)it) & 3u) * 4u + warp;  // column = warp, row = (warp + it) mod 4
        const uint32_t x = x0 + (sidx & 3u) * 8u + (lane & 7u), lr = r0 + (sidx >> 2) * 4u + (lane >> 3);
        const bool valid = x < J.F.gen.W && lr < J.F.local_rows;
        float3 d;
        ray_gen(J.F.gen, x, partition_local_to_y(J.F.part, lr), o, d);
        int count = -1;
        if (it == 1) {  // (uniform across the CTA)
            float t_far = fmaxf(fmaxf(probe_far[0], probe_far[1]), fmaxf(probe_far[2], probe_far[3]));
            if (usable) n = build_entry_list(J.bvh, o, FB, FB.angle * 0.25f, t_far, oct_inv, E, cnt);
        }
        if (n >= 0) {
            const Frustum FS = make_frustum(f3(__shfl_sync(full, d.x, 0), __shfl_sync(full, d.y, 0), __shfl_sync(full, d.z, 0)),
                                            f3(__shfl_sync(full, d.x, 7), __shfl_sync(full, d.y, 7), __shfl_sync(full, d.z, 7)),
                                            f3(__shfl_sync(full, d.x, 24), __shfl_sync(full, d.y, 24), __shfl_sync(full, d.z, 24)),
                                            f3(__shfl_sync(full, d.x, 31), __shfl_sync(full, d.y, 31), __shfl_sync(full, d.z, 31)));
            count = cull_entries(FS, oct_inv, E, n);
        }
        TraceHit h;
        h.t = 0.0f; h.prim = h.tri = MRT_MISS_ID;
        if (valid) {
            h = count >= 0 ? trace_from_entries(J.bvh, o, d, E, count, S, cnt) : trace_coherent(J.bvh, o, d, S, cnt);
            J.store_xy(x, lr, h, o, d);
        }
        if (it == 0) {  // farthest hit of the warp's probe sub-tile; a miss (or a pixel off the image) means "no bound"
            float far = (valid && h.prim != MRT_MISS_ID) ? h.t * 1.001f : 3.0e38f;
            for (int off = 16; off > 0; off >>= 1) far = fmaxf(far, __shfl_xor_sync(full, far, off));
            if (lane == 0) probe_far[warp] = far;
            __syncthreads();
        }
#ifdef ENTRY_DEBUG
        if (it > 0 && lane == 0 && blockIdx.x % 397 == 0) printf("tile %u sub %u: n %d count %d\n", blockIdx.x, sidx, n, count);
#endif
        __syncwarp();
    }
    flush_counters(cnt, counters, count_visits != 0);
}
#endif

// ---- one wave of the wavefront: queue entry k -> hit record k ----
struct QueueJob {
    static constexpr bool ANY_HIT = false;
    const float4* ray_o;
    const float4* ray_d;
    const uint32_t* count_ptr;
    unsigned long long* hits;  // t bits | tri << 32; MRT_HIT_PENDING until the ray has been traced
    const uint32_t* order;  // optional: ray k of the sorted order is queue entry order[k] (row n5 sort stage)
    // option "ray_split": the shade stage fills the queue from both ends -- rays it expects to be LONG (leaving their
    // surface at a grazing angle) from the front, the others from the back (slot cap - 1 - j) -- and the persistent launch
    // hands the front out first, so that the long rays start early and the launch ends with short ones
    const uint32_t* back_ptr;  // number of back entries, or nullptr
    uint32_t cap;
    MRT_D uint32_t count() const { return *count_ptr + (back_ptr ? *back_ptr : 0u); }
    MRT_D uint32_t slot(uint32_t i) const {
        if (!back_ptr) return i;
        const uint32_t nf = *count_ptr;
        return i < nf ? i : cap - 1u - (i - nf);
    }
    MRT_D bool load(uint32_t i0, float3& o, float3& d) const {
        const uint32_t i = slot(i0);
        const uint32_t k = order ? __ldg(&order[i]) : i;
        float4 o4 = __ldg(&ray_o[k]), d4 = __ldg(&ray_d[k]);
        o = f3(o4.x, o4.y, o4.z);
        d = f3(d4.x, d4.y, d4.z);
        return true;
    }
    const float4* ray_p;  // option "prepared_rays": (1/d, Sx) and (Sy, Sz, axes | octant) written by the shade stage, or nullptr
    const float4* ray_s;
    MRT_D bool load_prepared(uint32_t i0, float3& o, RayPre& pre) const {
        if (!ray_p) return false;
        const uint32_t i = slot(i0);
        const uint32_t k = order ? __ldg(&order[i]) : i;
        const float4 o4 = __ldg(&ray_o[k]);
        o = f3(o4.x, o4.y, o4.z);
        pre = ray_pre_unpack(__ldg(&ray_p[k]), __ldg(&ray_s[k]));
        return true;
    }
    MRT_D void store(uint32_t i0, const TraceHit& h) const {
        const uint32_t i = slot(i0);
        const uint32_t k = order ? __ldg(&order[i]) : i;  // hit records stay in queue order for the shade kernel
        // one 64-bit relaxed store: the record doubles as the "ray k is done" flag of the fused shade stage
        const unsigned long long rec = (unsigned long long)__float_as_uint(h.t) | ((unsigned long long)h.tri << 32);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(hits + k), "l"(rec) : "memory");
    }
};

// ---- ray sort between bounces (row n5, optional): stable counting sort of the queue by direction octant ----
// Compaction already leaves the queue in pixel order (neighbouring origins); binning by octant on top of that
// gives every warp rays that share the near/far plane selection and the child visiting order.  One histogram
// pass, one scan over 8 x tiles counters, one scatter that writes only the 4-byte permutation: the 32-byte ray
// records are not moved, the trace kernel gathers them through the permutation.
constexpr int OB_THREADS = 256, OB_ITEMS = 8, OB_TILE = OB_THREADS * OB_ITEMS, OB_WARPS = OB_THREADS / 32;

MRT_D unsigned ray_octant(float4 d) { return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u); }

__global__ void __launch_bounds__(OB_THREADS) k_oct_hist(const float4* __restrict__ ray_d, const uint32_t* __restrict__ count_ptr,
                                                         uint32_t* __restrict__ hist, uint32_t tiles) {
    __shared__ uint32_t h[8];
    if (threadIdx.x < 8) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t count = *count_ptr;
    const uint32_t base = blockIdx.x * OB_TILE;
    if (base < count) {
#pragma unroll
        for (int j = 0; j < OB_ITEMS; j++) {
            uint32_t i = base + j * OB_THREADS + threadIdx.x;
            bool valid = i < count;
            unsigned oct = valid ? ray_octant(__ldg(&ray_d[i])) : 8u + (threadIdx.x & 31);
            unsigned peers = __match_any_sync(0xFFFFFFFFu, oct);
            if (valid && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[oct], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) hist[threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(OB_THREADS) k_oct_scatter(const float4* __restrict__ ray_d, const uint32_t* __restrict__ count_ptr,
                                                            const uint32_t* __restrict__ offsets, uint32_t tiles,
                                                            uint32_t* __restrict__ order) {
    __shared__ uint32_t wc[OB_WARPS][8];
    __shared__ uint32_t gbase[8];
    const uint32_t count = *count_ptr;
    const uint32_t tile_base = blockIdx.x * OB_TILE;
    if (tile_base >= count) return;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < OB_WARPS * 8) (&wc[0][0])[threadIdx.x] = 0;
    if (threadIdx.x < 8) gbase[threadIdx.x] = offsets[threadIdx.x * tiles + blockIdx.x];
    __syncthreads();
    unsigned oct[OB_ITEMS];
    uint32_t loc[OB_ITEMS];
    const uint32_t wbase = tile_base + warp * (32 * OB_ITEMS);
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < OB_ITEMS; j++) {
        uint32_t i = wbase + j * 32 + lane;
        bool valid = i < count;
        oct[j] = valid ? ray_octant(__ldg(&ray_d[i])) : 8u + lane;
        unsigned peers = __match_any_sync(0xFFFFFFFFu, oct[j]);
        unsigned leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) {
            base = wc[warp][oct[j]];
            wc[warp][oct[j]] = base + __popc(peers);
        }
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        loc[j] = base + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 8) {  // exclusive prefix over the tile's warps, per octant
        uint32_t run = 0;
        for (int w = 0; w < OB_WARPS; w++) {
            uint32_t c = wc[w][threadIdx.x];
            wc[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OB_ITEMS; j++) {
        uint32_t i = wbase + j * 32 + lane;
        if (i < count) order[gbase[oct[j]] + wc[warp][oct[j]] + loc[j]] = i;
    }
}

// ---- ray sort by (origin cell, direction octant) (row n5, sort mode 2) ----
// key = 21-bit Morton code of the ray origin in a 128^3 grid over the root node's box, then the 3-bit direction octant:
// rays of one cell stay together (the locality the octant-only binning destroys) and, inside a cell, rays that share the
// near/far plane selection and the child visiting order.  Sorted with the 8-bit LSD radix sort of sort.cu (3 passes over
// 24 bits); entries beyond the queue's size get the key ~0 and end up last.  Only the 4-byte permutation moves.
MRT_D uint32_t spread7(uint32_t v) {  // 7 bits -> every third bit
    v &= 0x7Fu;
    v = (v | (v << 8)) & 0x0000700Fu;
    v = (v | (v << 4)) & 0x000430C3u;
    v = (v | (v << 2)) & 0x00049249u;
    return v;
}
__global__ void __launch_bounds__(256) k_ray_keys(const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                  const uint32_t* __restrict__ count_ptr, BvhDev bvh, uint32_t n,
                                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t key = ~0ull;
    if (k < *count_ptr && bvh.num_nodes) {
        const uint4 n0 = __ldg(reinterpret_cast<const uint4*>(bvh.nodes));  // root: grid origin + per-axis step exponents
        const float3 lo = f3(__uint_as_float(n0.x), __uint_as_float(n0.y), __uint_as_float(n0.z));
        const float3 ext = f3(255.0f * __uint_as_float((n0.w & 0xFFu) << 23), 255.0f * __uint_as_float((n0.w & 0xFF00u) << 15),
                              255.0f * __uint_as_float((n0.w & 0xFF0000u) << 7));
        const float4 o = __ldg(&ray_o[k]);
        const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - lo.x) / ext.x * 128.0f, 0.0f), 127.0f),
                       cy = (uint32_t)fminf(fmaxf((o.y - lo.y) / ext.y * 128.0f, 0.0f), 127.0f),
                       cz = (uint32_t)fminf(fmaxf((o.z - lo.z) / ext.z * 128.0f, 0.0f), 127.0f);
        const uint32_t cell = spread7(cx) | (spread7(cy) << 1) | (spread7(cz) << 2);
        key = ((uint64_t)cell << 3) | ray_octant(__ldg(&ray_d[k]));
    }
    keys[k] = key;
    vals[k] = k;
}

// ---- closest-hit query for mrt_trace_rays ----
struct QueryJob {
    static constexpr bool ANY_HIT = false;
    const float* ro;
    const float* rd;
    uint32_t n;
    uint32_t* ids;
    float* ts;
    MRT_D uint32_t count() const { return n; }
    MRT_D bool load(uint32_t i, float3& o, float3& d) const {
        o = f3(ro[3 * (size_t)i], ro[3 * (size_t)i + 1], ro[3 * (size_t)i + 2]);
        d = f3(rd[3 * (size_t)i], rd[3 * (size_t)i + 1], rd[3 * (size_t)i + 2]);
        return true;
    }
    MRT_D bool load_prepared(uint32_t, float3&, RayPre&) const { return false; }  // rays are set up by the traversal kernel
    MRT_D void store(uint32_t i, const TraceHit& h) const {
        ids[i] = h.prim;
        ts[i] = h.prim != MRT_MISS_ID ? h.t : 0.0f;
    }
};

// Shadow rays of MRT_SECONDARY_NEE_SUN: occlusion queries towards the sun's disc.  A free ray adds the contribution the
// shade stage computed for it to its pixel; a pixel has at most one shadow ray per wave and no other kernel runs on the
// context's stream meanwhile, so the add needs no atomic and the image does not depend on the queue order.
struct ShadowJob {
    static constexpr bool ANY_HIT = true;
    const float4* ray_o;   // origin, pixel
    const float4* ray_d;   // direction
    const float4* contrib; // throughput * E_sun * n.l * limb * Omega / pi
    const uint32_t* count_ptr;
    float4* accum;
    MRT_D uint32_t count() const { return *count_ptr; }
    MRT_D bool load(uint32_t i, float3& o, float3& d) const {
        float4 o4 = __ldg(&ray_o[i]), d4 = __ldg(&ray_d[i]);
        o = f3(o4.x, o4.y, o4.z);
        d = f3(d4.x, d4.y, d4.z);
        return true;
    }
    MRT_D bool load_prepared(uint32_t, float3&, RayPre&) const { return false; }  // rays are set up by the traversal kernel
    MRT_D void store(uint32_t i, const TraceHit& h) const {
        if (h.prim != MRT_MISS_ID) return;
        const uint32_t pixel = __float_as_uint(__ldg(&ray_o[i]).w);
        const float4 c = __ldg(&contrib[i]);
        float4 a = accum[pixel];
        accum[pixel] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w);
    }
};

// Shadow rays with the per-lane loop of the primary pass (option "shadow_coherent"): all of them point into the sun's
// 0.5-degree disc -- one direction octant, parallel within half a degree -- and the queue of the first vertex is in pixel
// order, so neighbouring lanes walk the same nodes.
__global__ void __launch_bounds__(TRACE_BLOCK, PRIMARY_MIN_BLOCKS)
k_shadow_coherent(ShadowJob job, BvhDev bvh, unsigned long long* counters, int count_visits, unsigned long long* total_rays) {
    __shared__ TraceShared S;
    trace_shared_init(S);
    const uint32_t n = job.count();
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)n);
    TraceCounters cnt{0, 0, 0};
    for (uint32_t i = blockIdx.x * TRACE_BLOCK + threadIdx.x; i < n; i += gridDim.x * TRACE_BLOCK) {
        float3 o, d;
        job.load(i, o, d);
        const TraceHit h = trace_coherent<true>(bvh, o, d, S, cnt);
        job.store(i, h);
    }
    flush_counters(cnt, counters, count_visits != 0);
}

template <class Job>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
k_trace(Job job, BvhDev bvh, uint32_t* work_counter, unsigned long long* counters, int count_visits,
        unsigned long long* total_rays = nullptr, unsigned long long extra_rays = 0) {
    __shared__ TraceShared S;
    // running ray total (mrt_stats.total_rays): this wave's rays (+ the frame's primary rays with its first wave)
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)job.count() + extra_rays);
    trace_shared_init(S);
    TraceCounters cnt{0, 0, 0};
    trace_persistent(bvh, job, work_counter, S, cnt);
    flush_counters(cnt, counters, count_visits != 0);
}

struct ShadeParams {
    float3 cameraPos;
    uint32_t seed;       // (frameCounter << 1) | 1, secondaryRays.comp:124
    uint32_t W;
    uint32_t spp;
    uint32_t bnW, bnH;
    uint32_t vertex;     // path vertex being shaded: 0 = primary hit
    uint32_t bounces;
    int first_sample;    // sample 0 of this frame: seed the PCG state, start/continue the accumulator
    int accumulate;
    Partition part;
    uint32_t pixel_base; // FIRST: vertex k of this launch is pixel pixel_base + k (bands of the image, see mesh_secondary)
    uint32_t ext;        // MRT_SECONDARY_NEE_SUN | MRT_SECONDARY_SKY_AT_HIT | MRT_SECONDARY_AERIAL
    uint32_t H;          // image height (aerial-perspective lookup)
    uint32_t first_tiles_x, first_count, first_rows;  // shade_tiles: tile columns, threads (tiles * 32) and rows of vertex 0; 0: row order
};

// Everything the shade stage reads and writes (passed by value to the kernels).
struct ShadeArgs {
    ShadeParams P;
    BvhDev bvh;
    mrt_atmosphere_params A;
    SkyLuts luts;
    const uchar4* bn;
    const float4* albedo;
    const float4* hit0_pos;       // FIRST: primary hits
    const float4* hit0_n;
    const float4* in_o;           // otherwise: the traced queue ...
    const float4* in_d;
    unsigned long long* hits;     // ... and its hit records
    uint32_t npix;
    float4* path_state;
    float4* accum;
    float4* out_o;                // next bounce's queue
    float4* out_d;
    float4* out_p;                // option "prepared_rays": ray_prepare(direction) of the next queue's rays (nullptr: off)
    float4* out_s;
    uint32_t* out_count;
    uint32_t* out_back;           // option "ray_split": entries taken from the back of the next queue (nullptr: off) ...
    uint32_t out_cap;             // ... whose capacity this is
    const uint32_t* in_back;      // ... and the same for the traced queue this stage reads
    uint32_t in_cap;
    float split_cos;              // a bounce ray with n.d below this is expected to be long
    unsigned long long* overflow; // error counter (mrt_stats.stack_overflows) for a fused worker that gave up waiting
    const float4* aerial;         // MRT_SECONDARY_AERIAL: decoded camera volume; hit_t = distance of the primary hit
    const float* hit_t;
    const float4* sun_e;          // MRT_SECONDARY_NEE_SUN: the sun centre's radiance seen from the camera (k_sun_centre)
    float4* sh_o;                 // MRT_SECONDARY_NEE_SUN: shadow-ray queue of this vertex (origin|pixel, direction, contribution)
    float4* sh_d;
    float4* sh_c;
    uint32_t* sh_count;
};

// hit record of a ray that has not been traced yet (byte pattern of cudaMemset(0xFE)): tri index 0xFEFEFEFE
#define MRT_HIT_PENDING_TRI 0xFEFEFEFEu

// Shade path vertex k.  Must be called by all 32 lanes of a warp (k >= count: lane idles through the
// compaction).  FIRST: vertex 0, k = pixel, inputs from the primary pass.  Otherwise k = queue entry of the
// traced wave.  FUSED: called from the traversal kernel while other warps still trace -- the lane waits for
// its hit record and hands the sentinel back for the next wave.
// Option "shade_tiles" (default off: measured neutral, profiles/r2_results.md): vertex 0 is shaded in the order of the primary pass' 8x4-pixel tiles, not row by row,
// so that 32 consecutive entries of the first bounce queue -- one warp of the traversal kernel -- leave from a compact
// 8x4 patch of the image instead of a 32x1 strip; later queues inherit the order through the compaction.  The image does not
// depend on the order (one live path per pixel).  k -> tile k / 32 (row-major tile grid), pixel k % 32 inside it.
MRT_D bool first_tile_valid(const ShadeParams& P, uint32_t k, uint32_t rows) {
    const uint32_t tile = k >> 5, in = k & 31u;
    return (tile % P.first_tiles_x) * 8u + (in & 7u) < P.W && (tile / P.first_tiles_x) * 4u + (in >> 3) < rows;
}
MRT_D uint32_t first_tile_pixel(const ShadeParams& P, uint32_t k) {
    const uint32_t tile = k >> 5, in = k & 31u;
    return ((tile / P.first_tiles_x) * 4u + (in >> 3)) * P.W + (tile % P.first_tiles_x) * 8u + (in & 7u);
}

template <bool FIRST, bool FUSED>
MRT_D void shade_vertex(uint32_t k, bool valid, const ShadeArgs& a) {
    const ShadeParams& P = a.P;
    bool emit = false, emit_shadow = false;
    float3 ro = f3s(0.0f), rd = f3s(0.0f), sl = f3s(0.0f), sc = f3s(0.0f);
    uint32_t pixel = 0;
    bool long_ray = true;
    if (valid) {
        float3 pos, n, thr;
        float3 sky_pos = P.cameraPos;  // MRT_SECONDARY_SKY_AT_HIT: the origin of the ray that escaped
        uint32_t prim, rng;
        if (FIRST) {
            pixel = P.first_tiles_x ? first_tile_pixel(P, k) : P.pixel_base + k;
            float4 hp = __ldg(&a.hit0_pos[pixel]), hn = __ldg(&a.hit0_n[pixel]);
            pos = f3(hp.x, hp.y, hp.z);
            n = f3(hn.x, hn.y, hn.z);
            prim = __float_as_uint(hp.w);
            thr = f3s(1.0f);
            rng = P.first_sample ? P.seed : __float_as_uint(a.path_state[pixel].w);
            const bool ap_on = (P.ext & MRT_SECONDARY_AERIAL) && prim != MRT_MISS_ID;
            if (P.first_sample || ap_on) {
                float4 acc = (!P.first_sample || P.accumulate) ? a.accum[pixel] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (P.first_sample) acc.w += (float)P.spp;
                if (ap_on) {  // contract: oracle ORC_EXT_AERIAL -- the sample starts behind the air between camera and hit
                    const uint32_t lr = pixel / P.W, x = pixel - lr * P.W;
                    const float4 ap = sky_aerial_lookup(a.aerial, ((float)x + 0.5f) / (float)P.W,
                                                        ((float)partition_local_to_y(P.part, lr) + 0.5f) / (float)P.H, __ldg(&a.hit_t[pixel]));
                    thr = f3s(1.0f - ap.w);
                    acc = make_float4(acc.x + ap.x, acc.y + ap.y, acc.z + ap.z, acc.w);
                }
                a.accum[pixel] = acc;
            }
        } else {
            unsigned long long rec;
            if (FUSED) {
                unsigned spins = 0;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(rec) : "l"(a.hits + k) : "memory");
                    if ((uint32_t)(rec >> 32) != MRT_HIT_PENDING_TRI) break;
                    if (++spins > (1u << 22)) {  // ~1 s: never expected; counted instead of hanging the GPU
                        atomicAdd(a.overflow, 1ull);
                        rec = (unsigned long long)MRT_MISS_ID << 32;
                        break;
                    }
                    __nanosleep(200);
                }
                a.hits[k] = ((unsigned long long)MRT_HIT_PENDING_TRI << 32) | MRT_HIT_PENDING_TRI;
            } else {
                rec = a.hits[k];
            }
            float4 o4 = __ldg(&a.in_o[k]), d4 = __ldg(&a.in_d[k]);
            pixel = __float_as_uint(o4.w);
            float3 o = f3(o4.x, o4.y, o4.z), d = f3(d4.x, d4.y, d4.z);
            const uint32_t tri = (uint32_t)(rec >> 32);
            float4 st = a.path_state[pixel];
            thr = f3(st.x, st.y, st.z);
            rng = __float_as_uint(st.w);
            if (tri != MRT_MISS_ID) {
                pos = o + d * __uint_as_float((uint32_t)rec);
                n = tri_facing_normal(a.bvh, tri, d, &prim);
            } else {
                prim = MRT_MISS_ID;
                pos = f3s(0.0f);
                n = d;  // secondaryRays.comp:88
                if (P.ext & MRT_SECONDARY_SKY_AT_HIT) sky_pos = o;
            }
        }
        if (prim == MRT_MISS_ID) {
            // secondaryRays.comp:96: the path ends in the sky (a bounce ray of an NEE path adds the sky-view term only)
            float3 c = thr * sky_color(a.A, a.luts, sky_pos, n, FIRST || !(P.ext & MRT_SECONDARY_NEE_SUN));
            float4 acc = a.accum[pixel];
            a.accum[pixel] = make_float4(acc.x + c.x, acc.y + c.y, acc.z + c.z, acc.w);
            if (FIRST) a.path_state[pixel] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng));
        } else {
            float4 al = __ldg(&a.albedo[prim]);
            thr = thr * f3(al.x, al.y, al.z);  // secondaryRays.comp:94
            if (P.vertex < P.bounces) {
                uint32_t lr = pixel / P.W, x = pixel - lr * P.W;
                float2 rot = blue_noise_rotation(a.bn, P.bnW, P.bnH, x, partition_local_to_y(P.part, lr));
                if (P.ext & MRT_SECONDARY_NEE_SUN) {  // contract: oracle orc_render_tris_ext; drawn BEFORE the bounce's numbers
                    const float u0 = rotated_random(rng, rot.x);
                    const float u1 = rotated_random(rng, rot.y);
                    float wgt;
                    sky_nee_sun_sample(u0, u1, sl, wgt);
                    const float ndotl = dot3(n, sl);
                    if (ndotl > 0.0f) {
                        const float3 so = pos + n * RAY_OFFSET;
                        // at the camera position E is one value per frame: k_sun_centre evaluated it once (same function, same bits)
                        float3 E;
                        if (P.ext & MRT_SECONDARY_SKY_AT_HIT) {
                            E = sky_sun_centre_radiance(a.A, a.luts, so);
                        } else {
                            const float4 e4 = __ldg(a.sun_e);
                            E = f3(e4.x, e4.y, e4.z);
                        }
                        if (E.x > 0.0f || E.y > 0.0f || E.z > 0.0f) {
                            sc = (thr * E) * (ndotl * wgt);
                            emit_shadow = true;
                        }
                    }
                }
                lambert_bounce(pos, n, rng, rot.x, rot.y, ro, rd);
                emit = true;
                long_ray = dot3(rd, n) < a.split_cos;
            }
            a.path_state[pixel] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng));
        }
    }
    // compaction: one atomic per warp, lanes take consecutive queue slots (ray_split: a second one for the back end)
    const bool to_front = emit && (long_ray || !a.out_back);
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, to_front);
    if (ballot) {
        const unsigned lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == (unsigned)(__ffs(ballot) - 1)) base = atomicAdd(a.out_count, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(ballot) - 1);
        if (to_front) {
            uint32_t slot = base + __popc(ballot & ((1u << lane) - 1u));
            a.out_o[slot] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(pixel));
            a.out_d[slot] = make_float4(rd.x, rd.y, rd.z, 0.0f);
            if (a.out_p) {
                float4 pa, pb;
                ray_pre_pack(ray_prepare(rd), pa, pb);
                a.out_p[slot] = pa;
                a.out_s[slot] = pb;
            }
        }
    }
    if (a.out_back) {
        const bool to_back = emit && !to_front;
        const unsigned bb = __ballot_sync(0xFFFFFFFFu, to_back);
        if (bb) {
            const unsigned lane = threadIdx.x & 31;
            uint32_t base = 0;
            if (lane == (unsigned)(__ffs(bb) - 1)) base = atomicAdd(a.out_back, (uint32_t)__popc(bb));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(bb) - 1);
            if (to_back) {
                uint32_t slot = a.out_cap - 1u - (base + __popc(bb & ((1u << lane) - 1u)));
                a.out_o[slot] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(pixel));
                a.out_d[slot] = make_float4(rd.x, rd.y, rd.z, 0.0f);
                if (a.out_p) {
                    float4 pa, pb;
                    ray_pre_pack(ray_prepare(rd), pa, pb);
                    a.out_p[slot] = pa;
                    a.out_s[slot] = pb;
                }
            }
        }
    }
    if (P.ext & MRT_SECONDARY_NEE_SUN) {  // shadow rays: same compaction into their own queue (origin = the bounce's)
        const unsigned sb = __ballot_sync(0xFFFFFFFFu, emit_shadow);
        if (sb) {
            const unsigned lane = threadIdx.x & 31;
            uint32_t base = 0;
            if (lane == (unsigned)(__ffs(sb) - 1)) base = atomicAdd(a.sh_count, (uint32_t)__popc(sb));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(sb) - 1);
            if (emit_shadow) {
                uint32_t slot = base + __popc(sb & ((1u << lane) - 1u));
                a.sh_o[slot] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(pixel));
                a.sh_d[slot] = make_float4(sl.x, sl.y, sl.z, 0.0f);
                a.sh_c[slot] = make_float4(sc.x, sc.y, sc.z, 0.0f);
            }
        }
    }
}

#ifndef SHADE_REGROUP
#define SHADE_REGROUP 1  // bounce waves: order the CTA's 256 vertices hits-first so that warps do not run both branches
#endif

__global__ void k_sun_centre(mrt_atmosphere_params A, SkyLuts luts, float3 cameraPos, float4* out) {
    const float3 e = sky_sun_centre_radiance(A, luts, cameraPos);
    out[0] = make_float4(e.x, e.y, e.z, 0.0f);
}

// Shade stage as its own kernel: one thread per path vertex.
template <bool FIRST>
__global__ void __launch_bounds__(256, 6) k_shade(ShadeArgs a, const uint32_t* __restrict__ in_count_ptr) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nfront = FIRST ? (a.P.first_tiles_x ? a.P.first_count : a.npix) : *in_count_ptr;
    const uint32_t count = nfront + ((!FIRST && a.in_back) ? *a.in_back : 0u);
    if (blockIdx.x * blockDim.x >= count) return;
    bool valid = k < count;
    if (FIRST && a.P.first_tiles_x) valid = valid && first_tile_valid(a.P, k, a.P.first_rows);
    if (!FIRST && a.in_back && valid && k >= nfront) k = a.in_cap - 1u - (k - nfront);  // ray_split: entry k sits at the back
#if SHADE_REGROUP
    if (!FIRST) {
        // A hit runs the bounce (triangle fetch, sincos), a miss the sky evaluation; mixed warps execute both, one
        // after the other (ncu: 17 of 32 lanes active).  Stable partition of the CTA's window: hits, then misses.
        __shared__ uint32_t perm[256];
        __shared__ uint32_t warp_hits[8];
        const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const bool is_hit = valid && (uint32_t)(a.hits[k] >> 32) != MRT_MISS_ID;
        const unsigned hb = __ballot_sync(0xFFFFFFFFu, is_hit);
        if (lane == 0) warp_hits[wid] = __popc(hb);
        __syncthreads();
        uint32_t hits_before = 0, hits_total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t h = warp_hits[w];
            hits_before += w < (int)wid ? h : 0u;
            hits_total += h;
        }
        const uint32_t hit_rank = hits_before + __popc(hb & ((1u << lane) - 1u));
        const uint32_t miss_rank = (wid * 32u - hits_before) + __popc(~hb & ((1u << lane) - 1u));
        perm[is_hit ? hit_rank : hits_total + miss_rank] = valid ? k : 0xFFFFFFFFu;  // invalid lanes count as misses and stay last
        __syncthreads();
        k = perm[threadIdx.x];
        valid = k != 0xFFFFFFFFu;
    }
#endif
    shade_vertex<FIRST, false>(k, valid, a);
}

// One bounce wave, traversal + shading in one persistent launch.  A warp leaves trace_persistent() when the
// queue is exhausted and its own rays are done; it then takes 32-entry chunks of the wave (shade_counter)
// and shades them, waiting per lane for hit records that slower warps are still producing.  All CTAs of
// the grid are resident (trace_grid) and waiting warps hold no rays, so the producers always progress.
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
k_trace_shade(QueueJob job, BvhDev bvh, uint32_t* work_counter, unsigned long long* counters, int count_visits,
              unsigned long long* total_rays, unsigned long long extra_rays, ShadeArgs sa, uint32_t* shade_counter) {
    __shared__ TraceShared S;
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)job.count() + extra_rays);
    trace_shared_init(S);
    TraceCounters cnt{0, 0, 0};
    trace_persistent(bvh, job, work_counter, S, cnt);
    flush_counters(cnt, counters, count_visits != 0);
    const unsigned lane = threadIdx.x & 31;
    const uint32_t count = job.count();
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(shade_counter, 32u);
        c = __shfl_sync(0xFFFFFFFFu, c, 0);
        if (c >= count) break;
        shade_vertex<false, true>(c + lane, c + lane < count, sa);
    }
}

// ---- the whole secondary pass in ONE persistent launch (option "path_kernel") ----
// The wavefront above pays, per bounce wave, a traversal launch whose last quarter is the drain of its longest rays
// (~100 us, bounded by ray latency, not by work), a shade launch, and ~130 B/ray of queue / hit / path-state traffic.
// With 4-8 spp x 2-3 bounces that is 12-24 drains per frame, and on a 1/8-image slab (tile mode on 8 GPUs) the drains
// ARE the frame (tools/bench_slab.py: 62 % render efficiency at 1/8).  Here a lane owns PIXELS for all their samples and
// bounces -- the reference's RNG chain forces the samples of a pixel to run in sequence anyway
// (secondaryRays.comp:124-132) -- nothing is queued in global memory and the only drain is the one at the end of the frame.
// A first version gave each lane one pixel: a lane whose ray had finished waited for a shade step, shade steps ran with
// 8-16 of 32 lanes and both branches (sky / bounce), and the kernel was 25 % slower than the wavefront.  Now each lane
// owns TWO path slots in shared memory; while one path waits to be shaded the lane traverses the other's ray, so
// shade steps can wait until 3/4 of the warp wants one.  The warp-synchronous state machine of trace_persistent() has two
// more step kinds:
//     refill : idle lanes start the ray a shaded slot holds ready; lanes with an empty slot take the next pixel;
//     shade  : lanes with a slot whose ray has finished (or that just took a pixel) resolve the path vertex -- sky on a
//              miss, albedo + blue-noise-rotated Lambert bounce on a hit -- and leave the next ray in the slot, start the
//              next sample, or finish the pixel.
// Per pixel the arithmetic and its order are exactly those of the wavefront (k_shade / shade_vertex), so the two
// produce bit-identical images (tested).
#ifndef PATH_SHADE_MIN
#define PATH_SHADE_MIN 24
#endif
#ifndef PATH_MIN_BLOCKS
#define PATH_MIN_BLOCKS TRACE_MIN_BLOCKS
#endif
enum { SL_PX = 0, SL_PY, SL_PZ, SL_DX, SL_DY, SL_DZ, SL_TX, SL_TY, SL_TZ, SL_TRI, SL_PIXEL, SL_RNG, SL_SV, SL_STATE, SL_COUNT };
enum { SLOT_EMPTY = 0, SLOT_SHADE = 1, SLOT_READY = 2, SLOT_ACTIVE = 3 };

__global__ void __launch_bounds__(TRACE_BLOCK, PATH_MIN_BLOCKS)
k_path(ShadeArgs a, uint32_t* work_counter, unsigned long long* counters, int count_visits, unsigned long long* total_rays) {
    __shared__ TraceShared S;
    __shared__ uint32_t slots[2][SL_COUNT][TRACE_BLOCK];  // two path slots per lane, column layout (conflict-free)
    trace_shared_init(S);
    const ShadeParams& P = a.P;
    const BvhDev& bvh = a.bvh;
    uint2* const sm = &S.stack[0][threadIdx.x];
#define SLU(s, k) slots[s][k][threadIdx.x]
#define SLF(s, k) (reinterpret_cast<float(*)[SL_COUNT][TRACE_BLOCK]>(slots))[s][k][threadIdx.x]
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t total = a.npix;
    uint2 spill[TRACE_LOCAL_STACK];
    TraceCounters cnt{0, 0, 0};
    LaneState L;
    L.ng = L.tg = L.tg2 = make_uint2(0u, 0u);
    L.tgmask = L.tg2mask = 0u;
    L.sp = 0;
    L.o = f3s(0.0f);
    L.hit.t = 0.0f; L.hit.tri = L.hit.prim = MRT_MISS_ID;
    SLU(0, SL_STATE) = SLOT_EMPTY;
    SLU(1, SL_STATE) = SLOT_EMPTY;
    int cur = -1;  // slot whose ray this lane is traversing
    unsigned nrays = 0;
    uint32_t pool_next = 0, pool_end = 0;
    bool exhausted = false;

    for (;;) {
        // ---- traversal bookkeeping: pop, or park the finished ray in its slot for the shade step
        if (cur >= 0 && !(L.ng.y & 0xFF000000u) && !(L.tg.y && L.tg2.y)) {
            if (L.sp == 0) {
                if (!(L.tg.y | L.tg2.y)) {
                    if (L.hit.tri != MRT_MISS_ID) {
                        const float3 d = f3(SLF(cur, SL_DX), SLF(cur, SL_DY), SLF(cur, SL_DZ));
                        const float3 pos = L.o + d * L.hit.t;
                        SLF(cur, SL_PX) = pos.x; SLF(cur, SL_PY) = pos.y; SLF(cur, SL_PZ) = pos.z;
                    }
                    SLU(cur, SL_TRI) = L.hit.tri;
                    SLU(cur, SL_STATE) = SLOT_SHADE;
                    cur = -1;
                }
            } else {
                L.sp--;
                L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
            }
        }
        const uint32_t st0 = SLU(0, SL_STATE), st1 = SLU(1, SL_STATE);
        const bool has_ready = st0 == SLOT_READY || st1 == SLOT_READY;
        const bool has_empty = st0 == SLOT_EMPTY || st1 == SLOT_EMPTY;
        const bool has_shade = st0 == SLOT_SHADE || st1 == SLOT_SHADE;
        const bool tracing = cur >= 0;
        const bool want_tri = tracing && (L.tg.y | L.tg2.y) != 0u;
        const bool want_node = tracing && !(L.tg.y && L.tg2.y) && (L.ng.y & 0xFF000000u);
        const unsigned tmask = __ballot_sync(0xFFFFFFFFu, want_tri);
        const unsigned nmask = __ballot_sync(0xFFFFFFFFu, want_node);
        const unsigned smask = __ballot_sync(0xFFFFFFFFu, has_shade);
        const unsigned amask = __ballot_sync(0xFFFFFFFFu, !tracing && has_ready);                 // lanes that could start a ray
        const unsigned pmask = __ballot_sync(0xFFFFFFFFu, has_empty && !exhausted);               // lanes that could take a pixel
        const unsigned busy = __ballot_sync(0xFFFFFFFFu, tracing || st0 != SLOT_EMPTY || st1 != SLOT_EMPTY);
        if (exhausted && busy == 0u) break;
        const bool no_trace_work = (tmask | nmask) == 0u;
        const unsigned idle_fillable = amask | (pmask & __ballot_sync(0xFFFFFFFFu, !tracing));
        const int tri_min = exhausted ? min(TRACE_TRI_MIN, max(1, (__popc(tmask | nmask) + TRACE_DRAIN_TRI - 1) / TRACE_DRAIN_TRI)) : TRACE_TRI_MIN;
        const int shade_min = exhausted ? min(PATH_SHADE_MIN, max(1, (__popc(busy) * 3 + 3) / 4)) : PATH_SHADE_MIN;

        if (idle_fillable && (__popc(idle_fillable) >= TRACE_REFILL_MIN || no_trace_work)) {
            // ---- refill step: new pixels into empty slots (every lane that has one), then idle lanes start a ready ray
            if (pmask) {
                if (pool_next >= pool_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(work_counter, (uint32_t)TRACE_CHUNK);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    pool_next = base;
                    pool_end = min(base + (uint32_t)TRACE_CHUNK, total);
                    if (base >= total) { exhausted = true; pool_next = pool_end = 0; }
                }
                if (!exhausted) {
                    const uint32_t mine = pool_next + __popc(pmask & lt_mask);
                    if ((pmask >> lane) & 1u) {
                        if (mine < pool_end) {
                            const int s = st0 == SLOT_EMPTY ? 0 : 1;
                            SLU(s, SL_PIXEL) = mine;
                            SLU(s, SL_RNG) = P.seed;
                            SLU(s, SL_SV) = 0u;
                            SLU(s, SL_STATE) = SLOT_SHADE;
                            float4 acc = P.accumulate ? a.accum[mine] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            acc.w += (float)P.spp;
                            a.accum[mine] = acc;
                        }
                    }
                    pool_next = min(pool_next + (uint32_t)__popc(pmask), pool_end);
                }
            }
            if (!tracing && has_ready) {
                const int s = st0 == SLOT_READY ? 0 : 1;
                lane_begin(L, f3(SLF(s, SL_PX), SLF(s, SL_PY), SLF(s, SL_PZ)), f3(SLF(s, SL_DX), SLF(s, SL_DY), SLF(s, SL_DZ)));
                if (bvh.num_nodes == 0) L.ng.y = 0u;
                SLU(s, SL_STATE) = SLOT_ACTIVE;
                cur = s;
                nrays++;
            }
        } else if (smask && (__popc(smask) >= shade_min || no_trace_work)) {
            // ---- shade step: one slot per lane
            if (has_shade) {
                const int s = st0 == SLOT_SHADE ? 0 : 1;
                const uint32_t pixel = SLU(s, SL_PIXEL);
                uint32_t rng = SLU(s, SL_RNG), sv = SLU(s, SL_SV);
                float3 thr = f3(SLF(s, SL_TX), SLF(s, SL_TY), SLF(s, SL_TZ));
                for (;;) {
                    const uint32_t vertex = sv & 0xFFFFu;
                    float3 pos, n;
                    uint32_t prim;
                    if (vertex == 0) {  // the primary hit, carried in fp32 by the primary pass (shade_vertex<FIRST>)
                        const float4 hp = __ldg(&a.hit0_pos[pixel]), hn = __ldg(&a.hit0_n[pixel]);
                        pos = f3(hp.x, hp.y, hp.z);
                        n = f3(hn.x, hn.y, hn.z);
                        prim = __float_as_uint(hp.w);
                        thr = f3s(1.0f);
                    } else {
                        const float3 d = f3(SLF(s, SL_DX), SLF(s, SL_DY), SLF(s, SL_DZ));
                        const uint32_t tri = SLU(s, SL_TRI);
                        if (tri != MRT_MISS_ID) {
                            pos = f3(SLF(s, SL_PX), SLF(s, SL_PY), SLF(s, SL_PZ));
                            n = tri_facing_normal(bvh, tri, d, &prim);
                        } else {
                            prim = MRT_MISS_ID;
                            pos = f3s(0.0f);
                            n = d;  // secondaryRays.comp:88
                        }
                    }
                    bool sample_done;
                    if (prim == MRT_MISS_ID) {  // secondaryRays.comp:96: the path ends in the sky
                        const float3 c = thr * sky_color(a.A, a.luts, P.cameraPos, n);
                        const float4 acc = a.accum[pixel];
                        a.accum[pixel] = make_float4(acc.x + c.x, acc.y + c.y, acc.z + c.z, acc.w);
                        sample_done = true;
                    } else {
                        const float4 al = __ldg(&a.albedo[prim]);
                        thr = thr * f3(al.x, al.y, al.z);  // secondaryRays.comp:94
                        sample_done = vertex >= P.bounces;  // the energy of a path that is still on a surface is dropped (:99)
                        if (!sample_done) {
                            const uint32_t lr = pixel / P.W, x = pixel - lr * P.W;
                            const float2 rot = blue_noise_rotation(a.bn, P.bnW, P.bnH, x, partition_local_to_y(P.part, lr));
                            float3 ro, rd;
                            lambert_bounce(pos, n, rng, rot.x, rot.y, ro, rd);
                            SLF(s, SL_PX) = ro.x; SLF(s, SL_PY) = ro.y; SLF(s, SL_PZ) = ro.z;
                            SLF(s, SL_DX) = rd.x; SLF(s, SL_DY) = rd.y; SLF(s, SL_DZ) = rd.z;
                            sv++;
                            SLU(s, SL_STATE) = SLOT_READY;
                            break;
                        }
                    }
                    const uint32_t sample = (sv >> 16) + 1u;
                    if (sample < P.spp) { sv = sample << 16; continue; }  // next sample: vertex 0 again, same RNG stream
                    SLU(s, SL_STATE) = SLOT_EMPTY;
                    break;
                }
                SLU(s, SL_RNG) = rng;
                SLU(s, SL_SV) = sv;
                SLF(s, SL_TX) = thr.x; SLF(s, SL_TY) = thr.y; SLF(s, SL_TZ) = thr.z;
            }
        } else if (tmask && (nmask == 0u || __popc(tmask) >= tri_min)) {
            if (want_tri) {
                lane_tri_step<true>(L, bvh, cnt);
#pragma unroll 1
                for (int k = 1; k < TRACE_TRI_PER_STEP && (L.tg.y | L.tg2.y); k++) lane_tri_step<true>(L, bvh, cnt);
            }
        } else if (nmask) {
            if (want_node) lane_node_step<true>(L, bvh, S, spill, cnt);
        }
    }
#undef SLU
#undef SLF
    flush_counters(cnt, counters, count_visits != 0);
    // rays traced by this launch: counters[3]; running total (mrt_stats.total_rays) += primary pixels + rays
    unsigned long long r = nrays;
    for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xFFFFFFFFu, r, off);
    if (lane == 0 && r) {
        atomicAdd(&counters[3], r);
        if (total_rays) atomicAdd(total_rays, r);
    }
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)a.npix);
}

// brute-force closest hit over the uploaded mesh (validation of the BVH path; mrt_trace_rays)
__global__ void __launch_bounds__(128) k_trace_brute(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t ntris,
                                                     const float* __restrict__ ro, const float* __restrict__ rdir, uint32_t n,
                                                     uint32_t* __restrict__ ids, float* __restrict__ ts) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float3 o = f3(ro[3 * k], ro[3 * k + 1], ro[3 * k + 2]), d = f3(rdir[3 * k], rdir[3 * k + 1], rdir[3 * k + 2]);
    RayShear rs = make_shear(d);
    TraceHit h;
    h.t = 0.0f; h.tri = h.prim = MRT_MISS_ID;
    for (uint32_t i = 0; i < ntris; i++) {
        uint32_t i0 = idx[3 * (size_t)i], i1 = idx[3 * (size_t)i + 1], i2 = idx[3 * (size_t)i + 2];
        float t;
        if (tri_test(o, rs, f3(pos[3 * (size_t)i0], pos[3 * (size_t)i0 + 1], pos[3 * (size_t)i0 + 2]),
                     f3(pos[3 * (size_t)i1], pos[3 * (size_t)i1 + 1], pos[3 * (size_t)i1 + 2]),
                     f3(pos[3 * (size_t)i2], pos[3 * (size_t)i2 + 1], pos[3 * (size_t)i2 + 2]), t))
            hit_consider(h, t, i, i);
    }
    ids[k] = h.prim;
    ts[k] = h.prim != MRT_MISS_ID ? h.t : 0.0f;
}

// running ray total for frames without bounce waves (otherwise k_trace adds its wave): primary pixels only
__global__ void k_sum_rays(const uint32_t* __restrict__ counts, uint32_t n, unsigned long long extra, unsigned long long* total) {
    unsigned long long s = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 32) s += counts[i];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, off);
    if (threadIdx.x == 0) *total += s + extra;
}

BvhDev make_bvh(mrt_context* ctx) {
    BvhDev b;
    b.nodes = ctx->nodes.p;
    b.tris = ctx->tris.p;
    b.num_nodes = ctx->num_nodes;
    b.num_tris = ctx->num_leaf_tris;
    b.prmt_k = 0x47000000u;
    return b;
}

// persistent grid: every SM gets its full complement of resident trace CTAs
unsigned trace_grid(mrt_context* ctx, size_t max_rays) {
    static int per_sm[64] = {0};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int& occ = per_sm[ctx->device & 63];
    if (occ == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<QueueJob>, TRACE_BLOCK, 0);
        if (occ <= 0) occ = 4;
    }
    size_t want = (size_t)sms * (ctx->opt_trace_ctas_per_sm > 0 && ctx->opt_trace_ctas_per_sm < occ ? ctx->opt_trace_ctas_per_sm : occ);
    size_t need = (max_rays + 31) / 32 / (TRACE_BLOCK / 32) + 1;  // no more CTAs than there are warps of work
    return (unsigned)(need < want ? need : want);
}

unsigned path_grid(mrt_context* ctx, size_t pixels) {
    static int per_sm[64] = {0};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int& occ = per_sm[ctx->device & 63];
    if (occ == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_path, TRACE_BLOCK, 0);
        if (occ <= 0) occ = 4;
    }
    size_t want = (size_t)sms * (ctx->opt_trace_ctas_per_sm > 0 && ctx->opt_trace_ctas_per_sm < occ ? ctx->opt_trace_ctas_per_sm : occ);
    size_t need = (pixels + 31) / 32 / (TRACE_BLOCK / 32) + 1;
    return (unsigned)(need < want ? need : want);
}

}  // namespace

int mesh_primary(mrt_context* ctx) {
    PrimaryJob J;
    memcpy(&J.F.gen.invView, &ctx->pc.invView, sizeof(Mat4));
    memcpy(&J.F.gen.invProj, &ctx->pc.invProjection, sizeof(Mat4));
    J.F.gen.W = ctx->W;
    J.F.gen.H = ctx->H;
    Mat4 P, V, Vp;
    memcpy(&P, &ctx->pc.projection, sizeof(Mat4));
    memcpy(&V, &ctx->pc.view, sizeof(Mat4));
    memcpy(&Vp, &ctx->pc.prevView, sizeof(Mat4));
    J.F.PV = mat_mul(P, V);
    J.F.PVprev = mat_mul(P, Vp);
    J.F.part = ctx->part;
    J.F.local_rows = ctx->local_rows;
    MRT_TRY(dev_reserve(ctx, ctx->hit_t, ctx->npix));
    MRT_TRY(dev_reserve(ctx, ctx->hit0_pos, ctx->npix));
    MRT_TRY(dev_reserve(ctx, ctx->hit0_n, ctx->npix));
    MRT_TRY(reserve_visit_counters(ctx));
    MRT_TRY(dev_reserve(ctx, ctx->counters, 16));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->visit_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->counters.p + 8, 0, sizeof(uint32_t), ctx->stream));
    J.bvh = make_bvh(ctx);
    J.tiles_x = div_up(ctx->W, PRIMARY_TILE_W);
    J.tiles = J.tiles_x * div_up(ctx->local_rows, 32u / PRIMARY_TILE_W);
    J.vis = ctx->visibility.p; J.depth = ctx->depth.p; J.normal = ctx->normal.p; J.motion = ctx->motion.p;
    J.hit_t = ctx->hit_t.p; J.hit0_pos = ctx->hit0_pos.p; J.hit0_n = ctx->hit0_n.p;
    if (ctx->opt_persistent_primary)
        k_trace<PrimaryJob><<<trace_grid(ctx, (size_t)J.tiles * 32), TRACE_BLOCK, 0, ctx->stream>>>(
            J, J.bvh, ctx->counters.p + 8, ctx->visit_counters.p, ctx->opt_count_visits);
#if PRIMARY_ENTRY
    else if (ctx->opt_primary_entry) {
        const uint32_t big_x = div_up(ctx->W, BIG_W), big_y = div_up(ctx->local_rows, BIG_H);
        k_mesh_primary_entry<<<big_x * big_y, TRACE_BLOCK, 0, ctx->stream>>>(J, big_x, ctx->visit_counters.p, ctx->opt_count_visits);
    }
#endif
    else if (ctx->opt_primary_batched)
        k_mesh_primary<true><<<div_up((size_t)J.tiles * 32, TRACE_BLOCK), TRACE_BLOCK, 0, ctx->stream>>>(J, ctx->visit_counters.p,
                                                                                                        ctx->opt_count_visits);
    else
        k_mesh_primary<false><<<div_up((size_t)J.tiles * 32, TRACE_BLOCK), TRACE_BLOCK, 0, ctx->stream>>>(J, ctx->visit_counters.p,
                                                                                                         ctx->opt_count_visits);
    MRT_LAUNCHED(ctx);
    ctx->stats.primary_rays = ctx->npix;
    return mrt_check_cuda(ctx, cudaGetLastError(), "mesh_primary");
}

int mesh_secondary(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp, uint32_t bounces, uint32_t flags) {
    const uint32_t npix = (uint32_t)ctx->npix;
    const bool frame_sum = (flags & MRT_SECONDARY_FRAME_SUM) != 0;  // samples go to the per-frame buffer (mrt_accum_commit)
    if (frame_sum) MRT_TRY(dev_reserve(ctx, ctx->frame_sum, npix));
    const uint32_t ext = flags & (MRT_SECONDARY_NEE_SUN | MRT_SECONDARY_SKY_AT_HIT | MRT_SECONDARY_AERIAL);
    const bool nee = (ext & MRT_SECONDARY_NEE_SUN) != 0;
    // the sky extensions live in the wavefront's shade stage only
    const bool path_kernel = ctx->opt_path_kernel != 0 && spp < 65536u && bounces < 65535u && ext == 0;
    ctx->secondary_was_path_kernel = path_kernel;
    if (path_kernel) {
        // one persistent launch for all samples and bounces: no queues, no hit records, no path-state buffer
        MRT_TRY(dev_reserve(ctx, ctx->queue_counts, 4));
        MRT_TRY(reserve_visit_counters(ctx));
        MRT_CUDA(ctx, cudaMemsetAsync(ctx->queue_counts.p, 0, sizeof(uint32_t) * 4, ctx->stream));
        MRT_CUDA(ctx, cudaMemsetAsync(ctx->visit_counters.p + 4, 0, 4 * sizeof(unsigned long long), ctx->stream));
        ctx->num_queue_counts = 0;
        ctx->num_shadow_counts = 0;
        ctx->num_back_counts = 0;
        ShadeArgs sa{};
        ShadeParams& P = sa.P;
        P.cameraPos = f3(c->cameraPos[0], c->cameraPos[1], c->cameraPos[2]);
        P.seed = (c->frameCounter << 1u) | 1u;
        P.W = ctx->W;
        P.spp = spp;
        P.bnW = ctx->bnW;
        P.bnH = ctx->bnH;
        P.bounces = bounces;
        P.accumulate = ((flags & MRT_SECONDARY_ACCUMULATE) && ctx->have_accum && !frame_sum) ? 1 : 0;
        P.part = ctx->part;
        sa.bvh = make_bvh(ctx);
        sa.A = ctx->atmo;
        sa.luts = SkyLuts{ctx->trans_f.p, nullptr, ctx->view_f.p};
        sa.bn = ctx->bn;
        sa.albedo = ctx->albedo.p;
        sa.hit0_pos = ctx->hit0_pos.p;
        sa.hit0_n = ctx->hit0_n.p;
        sa.npix = npix;
        sa.accum = frame_sum ? ctx->frame_sum.p : ctx->accum.p;
        const bool timed = ctx->opt_trace_timing && ctx->trace_ev_used < 4096;
        while (timed && ctx->trace_ev.size() < 2 * (size_t)(ctx->trace_ev_used + 1)) {
            cudaEvent_t e;
            MRT_CUDA(ctx, cudaEventCreate(&e));
            ctx->trace_ev.push_back(e);
        }
        if (timed) cudaEventRecord(ctx->trace_ev[2 * ctx->trace_ev_used], ctx->stream);
        k_path<<<path_grid(ctx, npix), TRACE_BLOCK, 0, ctx->stream>>>(sa, ctx->queue_counts.p, ctx->visit_counters.p + 4, ctx->opt_count_visits,
                                                                     ctx->total_rays.p);
        MRT_LAUNCHED(ctx);
        if (timed) {
            cudaEventRecord(ctx->trace_ev[2 * ctx->trace_ev_used + 1], ctx->stream);
            ctx->trace_ev_used++;
        }
        return mrt_check_cuda(ctx, cudaGetLastError(), "mesh_secondary (path kernel)");
    }
    const uint32_t waves = spp * bounces;
    const bool fused = ctx->opt_fused_shade != 0 && !nee;
    {   // L1 / shared-memory split of the traversal kernels (option "trace_carveout": percent of the 228 KB that is shared
        // memory; -1 = the driver's choice).  6 CTAs x 13.3 KB need 86 KB; what is not shared memory is L1 for the BVH.
        static int applied[64] = {0};
        int& a = applied[ctx->device & 63];
        if (a != ctx->opt_trace_carveout + 2) {
            cudaFuncSetAttribute(k_trace<QueueJob>, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->opt_trace_carveout);
            cudaFuncSetAttribute(k_trace<ShadowJob>, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->opt_trace_carveout);
            cudaFuncSetAttribute(k_mesh_primary<false>, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->opt_trace_carveout);
            a = ctx->opt_trace_carveout + 2;
        }
    }
    // sort mode: 0 none, 1 direction octant (binning), 2 (origin cell, octant) (radix sort); the flag asks for the configured
    // mode, or the octant binning when none is configured
    const int sort_mode = ctx->opt_sort_rays ? ctx->opt_sort_rays : ((flags & MRT_SECONDARY_SORT_RAYS) ? 1 : 0);
    const bool sort = sort_mode != 0;
    // Bands: the image's pixels are cut into contiguous ranges, each with its own queues, counters and STREAM.  The last
    // quarter of every persistent traversal launch is the drain of its longest rays (latency-bound, ~100 us whatever the
    // wave's size); with several bands in flight the drain of one band's launch is filled by another band's kernels --
    // the frames-in-flight overlap brought inside a frame.  Pixels are independent, so the image does
    // not depend on the number of bands (tested).  fused_shade and the sort stage keep one band (shared scratch).
    uint32_t bands = (fused || sort || ctx->opt_bands < 1) ? 1u : (uint32_t)ctx->opt_bands;
    if (bands > MRT_MAX_BANDS) bands = MRT_MAX_BANDS;
    while (bands > 1 && npix / bands < 32768u) bands--;  // a band must still fill the GPU
    MRT_TRY(dev_reserve(ctx, ctx->path_state, npix));
    const bool prepared = ctx->opt_prepared_rays != 0;
    for (int q = 0; q < 2; q++) {
        MRT_TRY(dev_reserve(ctx, ctx->ray_o[q], npix));
        MRT_TRY(dev_reserve(ctx, ctx->ray_d[q], npix));
        if (prepared) {
            MRT_TRY(dev_reserve(ctx, ctx->ray_p[q], npix));
            MRT_TRY(dev_reserve(ctx, ctx->ray_s[q], npix));
        }
    }
    if (npix > ctx->hits.cap || !ctx->hits.p) ctx->hits_dirty = true;
    MRT_TRY(dev_reserve(ctx, ctx->hits, npix));
    if (fused && ctx->hits_dirty) {  // every record "pending" (MRT_HIT_PENDING_TRI); the fused workers keep it that way
        MRT_CUDA(ctx, cudaMemsetAsync(ctx->hits.p, 0xFE, sizeof(unsigned long long) * ctx->hits.cap, ctx->stream));
        ctx->hits_dirty = false;
    }
    // counters: [band][0..waves] queue sizes; then [band][wave] work counters of the persistent launches; then
    // [band][wave] chunk counters of their (fused) shade stages
    // ... then, with MRT_SECONDARY_NEE_SUN, [band][wave] shadow-queue sizes and [band][wave] their work counters
    // ... then, with option ray_split, [band][0..waves] sizes of the queues' back ends
    const bool split = ctx->opt_ray_split != 0 && !fused && !sort;
    const size_t split_at = (size_t)bands * (3 * (size_t)waves + 3) + (nee ? 2 * (size_t)bands * waves : 0);
    const size_t ncount = split_at + (split ? (size_t)bands * (waves + 1) : 0);
    if (nee)
        for (int q = 0; q < 3; q++) MRT_TRY(dev_reserve(ctx, ctx->shadow_q[q], npix));
    MRT_TRY(dev_reserve(ctx, ctx->queue_counts, ncount));
    MRT_TRY(reserve_visit_counters(ctx));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->queue_counts.p, 0, sizeof(uint32_t) * ncount, ctx->stream));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->visit_counters.p + 4, 0, 4 * sizeof(unsigned long long), ctx->stream));
    ctx->num_queue_counts = bands * (waves + 1);
    ctx->shadow_counts_at = nee ? (uint32_t)(bands * (3 * (size_t)waves + 3)) : 0u;  // mrt_stats: shadow rays are secondary rays
    ctx->num_shadow_counts = nee ? bands * waves : 0u;
    ctx->back_counts_at = split ? (uint32_t)split_at : 0u;
    ctx->num_back_counts = split ? bands * (waves + 1) : 0u;
    if (bands > 1) {
        for (uint32_t b = 0; b < bands; b++) {
            if (!ctx->band_stream[b]) {
                MRT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream[b], cudaStreamNonBlocking));
                MRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_done[b], cudaEventDisableTiming));
            }
        }
        if (!ctx->band_fork) MRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_fork, cudaEventDisableTiming));
        MRT_CUDA(ctx, cudaEventRecord(ctx->band_fork, ctx->stream));
    }

    ShadeArgs sa{};
    ShadeParams& P = sa.P;
    P.cameraPos = f3(c->cameraPos[0], c->cameraPos[1], c->cameraPos[2]);
    P.seed = (c->frameCounter << 1u) | 1u;
    P.W = ctx->W;
    P.spp = spp;
    P.bnW = ctx->bnW;
    P.bnH = ctx->bnH;
    P.bounces = bounces;
    P.accumulate = ((flags & MRT_SECONDARY_ACCUMULATE) && ctx->have_accum && !frame_sum) ? 1 : 0;
    P.part = ctx->part;
    P.ext = ext;
    P.H = ctx->H;
    sa.aerial = ctx->aerial_f.p;
    sa.hit_t = ctx->hit_t.p;
    sa.bvh = make_bvh(ctx);
    sa.A = ctx->atmo;
    sa.luts = SkyLuts{ctx->trans_f.p, nullptr, ctx->view_f.p};
    if (nee) {
        MRT_TRY(dev_reserve(ctx, ctx->sun_e, 1));
        k_sun_centre<<<1, 1, 0, ctx->stream>>>(sa.A, sa.luts, P.cameraPos, ctx->sun_e.p);
        MRT_LAUNCHED(ctx);
        sa.sun_e = ctx->sun_e.p;
    }
    sa.bn = ctx->bn;
    sa.albedo = ctx->albedo.p;
    sa.hit0_pos = ctx->hit0_pos.p;
    sa.hit0_n = ctx->hit0_n.p;
    sa.path_state = ctx->path_state.p;
    sa.accum = frame_sum ? ctx->frame_sum.p : ctx->accum.p;
    sa.overflow = ctx->visit_counters.p + 4 + 2;
    constexpr uint32_t kMaxTimedLaunches = 4096;  // event pairs kept since mrt_stats_reset

    for (uint32_t band = 0; band < bands; band++) {
        // pixel range of the band: multiples of 256 so that the shade CTAs of different bands never share a pixel
        const uint32_t p0 = (uint32_t)(((uint64_t)npix * band / bands) & ~255ull);
        const uint32_t p1 = band + 1 == bands ? npix : (uint32_t)(((uint64_t)npix * (band + 1) / bands) & ~255ull);
        const uint32_t bpix = p1 - p0;
        if (bpix == 0) continue;
        cudaStream_t st = bands > 1 ? ctx->band_stream[band] : ctx->stream;
        if (bands > 1) MRT_CUDA(ctx, cudaStreamWaitEvent(st, ctx->band_fork, 0));
        uint32_t* const qcounts = ctx->queue_counts.p + (size_t)band * (waves + 1);
        uint32_t* const work_counters = ctx->queue_counts.p + (size_t)bands * (waves + 1) + (size_t)band * waves;
        uint32_t* const shade_counters = ctx->queue_counts.p + (size_t)bands * (2 * (size_t)waves + 1) + (size_t)band * waves;
        float4* const ray_o[2] = {ctx->ray_o[0].p + p0, ctx->ray_o[1].p + p0};
        float4* const ray_d[2] = {ctx->ray_d[0].p + p0, ctx->ray_d[1].p + p0};
        float4* const ray_p[2] = {prepared ? ctx->ray_p[0].p + p0 : nullptr, prepared ? ctx->ray_p[1].p + p0 : nullptr};
        float4* const ray_s[2] = {prepared ? ctx->ray_s[0].p + p0 : nullptr, prepared ? ctx->ray_s[1].p + p0 : nullptr};
        sa.hits = ctx->hits.p + p0;
        sa.npix = bpix;
        P.pixel_base = p0;
        const unsigned shade_grid = div_up(bpix, 256), tgrid = trace_grid(ctx, bpix);
        // vertex 0 in tile order (one band only: bands are pixel ranges in row order)
        const bool tiles = ctx->opt_shade_tiles != 0 && bands == 1;
        P.first_tiles_x = tiles ? div_up(ctx->W, 8u) : 0u;
        P.first_rows = ctx->local_rows;
        P.first_count = tiles ? P.first_tiles_x * div_up(ctx->local_rows, 4u) * 32u : bpix;
        const unsigned first_grid = div_up(P.first_count, 256);
        uint32_t* const sh_counts = nee ? ctx->queue_counts.p + ctx->shadow_counts_at + (size_t)band * waves : nullptr;
        uint32_t* const sh_work = nee ? ctx->queue_counts.p + ctx->shadow_counts_at + (size_t)bands * waves + (size_t)band * waves : nullptr;
        if (nee) {
            sa.sh_o = ctx->shadow_q[0].p + p0;
            sa.sh_d = ctx->shadow_q[1].p + p0;
            sa.sh_c = ctx->shadow_q[2].p + p0;
        }
        uint32_t* const bcounts = split ? ctx->queue_counts.p + split_at + (size_t)band * (waves + 1) : nullptr;
        sa.out_cap = sa.in_cap = bpix;
        sa.split_cos = 0.01f * (float)ctx->opt_ray_split;  // option value = 100 * the cosine below which a ray counts as long
        uint32_t wave = 0;
        for (uint32_t s = 0; s < spp; s++) {
            P.first_sample = s == 0;
            P.vertex = 0;
            int q = 0;
            sa.in_o = sa.in_d = nullptr;
            sa.out_o = ray_o[q];
            sa.out_d = ray_d[q];
            sa.out_p = ray_p[q];
            sa.out_s = ray_s[q];
            sa.out_count = qcounts + wave;
            sa.out_back = split ? bcounts + wave : nullptr;
            sa.in_back = nullptr;
            sa.sh_count = nee ? sh_counts + wave : nullptr;
            k_shade<true><<<first_grid, 256, 0, st>>>(sa, nullptr);
            MRT_LAUNCHED(ctx);
            for (uint32_t b = 1; b <= bounces; b++) {
                const uint32_t* in_count = qcounts + wave;
                if (nee) {
                    // the shadow rays of the vertex just shaded: any-hit traversal, free rays add their sun light to the
                    // accumulator; on the same stream, i.e. before the shade stage below touches the same pixels
                    ShadowJob SJ{sa.sh_o, sa.sh_d, sa.sh_c, sh_counts + wave, sa.accum};
                    if (ctx->opt_shadow_coherent)
                        k_shadow_coherent<<<div_up(bpix, TRACE_BLOCK), TRACE_BLOCK, 0, st>>>(SJ, sa.bvh, ctx->visit_counters.p + 4, ctx->opt_count_visits,
                                                                                             ctx->total_rays.p);
                    else
                        k_trace<ShadowJob><<<tgrid, TRACE_BLOCK, 0, st>>>(SJ, sa.bvh, sh_work + wave, ctx->visit_counters.p + 4,
                                                                          ctx->opt_count_visits, ctx->total_rays.p, 0ull);
                    MRT_LAUNCHED(ctx);
                }
                const bool timed = ctx->opt_trace_timing && ctx->trace_ev_used < kMaxTimedLaunches;
                while (timed && ctx->trace_ev.size() < 2 * (size_t)(ctx->trace_ev_used + 1)) {
                    cudaEvent_t e;
                    MRT_CUDA(ctx, cudaEventCreate(&e));
                    ctx->trace_ev.push_back(e);
                }
                if (timed) cudaEventRecord(ctx->trace_ev[2 * ctx->trace_ev_used], st);
                const uint32_t* order = nullptr;
                if (sort_mode == 2) {
                    MRT_TRY(dev_reserve(ctx, ctx->sort_keys, bpix));
                    MRT_TRY(dev_reserve(ctx, ctx->sort_keys_alt, bpix));
                    MRT_TRY(dev_reserve(ctx, ctx->sort_vals, bpix));
                    MRT_TRY(dev_reserve(ctx, ctx->sort_vals_alt, bpix));
                    k_ray_keys<<<div_up(bpix, 256), 256, 0, st>>>(ray_o[q], ray_d[q], in_count, sa.bvh, bpix, ctx->sort_keys.p, ctx->sort_vals.p);
                    MRT_LAUNCHED(ctx);
                    bool in_alt = false;
                    MRT_TRY(radix_sort_pairs_u64(ctx, ctx->sort_keys.p, ctx->sort_keys_alt.p, ctx->sort_vals.p, ctx->sort_vals_alt.p, bpix, 0, 24, &in_alt));
                    order = in_alt ? ctx->sort_vals_alt.p : ctx->sort_vals.p;
                } else if (sort) {
                    const uint32_t tiles = div_up(npix, OB_TILE);
                    MRT_TRY(dev_reserve(ctx, ctx->sort_vals, npix));
                    MRT_TRY(dev_reserve(ctx, ctx->sort_vals_alt, 8 * (size_t)tiles));
                    k_oct_hist<<<tiles, OB_THREADS, 0, st>>>(ray_d[q], in_count, ctx->sort_vals_alt.p, tiles);
                    MRT_LAUNCHED(ctx);
                    MRT_TRY(scan_exclusive_u32(ctx, ctx->sort_vals_alt.p, ctx->sort_vals_alt.p, 8 * (size_t)tiles));
                    k_oct_scatter<<<tiles, OB_THREADS, 0, st>>>(ray_d[q], in_count, ctx->sort_vals_alt.p, tiles, ctx->sort_vals.p);
                    MRT_LAUNCHED(ctx);
                    order = ctx->sort_vals.p;
                }
                QueueJob J{ray_o[q], ray_d[q], in_count, sa.hits, order, split ? bcounts + wave : nullptr, bpix, ray_p[q], ray_s[q]};
                const unsigned long long extra = wave == 0 ? (unsigned long long)bpix : 0ull;
                // the shade stage of this wave: vertex b of the paths; the last vertex emits nothing (its counter slot stays 0)
                P.vertex = b;
                sa.in_o = ray_o[q];
                sa.in_d = ray_d[q];
                sa.out_o = ray_o[q ^ 1];
                sa.out_d = ray_d[q ^ 1];
                sa.out_p = ray_p[q ^ 1];
                sa.out_s = ray_s[q ^ 1];
                sa.out_count = qcounts + (wave + 1 < waves ? wave + 1 : waves);
                sa.in_back = split ? bcounts + wave : nullptr;
                sa.out_back = split ? bcounts + (wave + 1 < waves ? wave + 1 : waves) : nullptr;
                sa.sh_count = nee ? sh_counts + (wave + 1 < waves ? wave + 1 : 0) : nullptr;  // the last vertex emits none
                if (fused) {
                    k_trace_shade<<<tgrid, TRACE_BLOCK, 0, st>>>(J, sa.bvh, work_counters + wave, ctx->visit_counters.p + 4,
                                                                 ctx->opt_count_visits, ctx->total_rays.p, extra, sa, shade_counters + wave);
                    MRT_LAUNCHED(ctx);
                } else {
                    k_trace<QueueJob><<<tgrid, TRACE_BLOCK, 0, st>>>(J, sa.bvh, work_counters + wave, ctx->visit_counters.p + 4,
                                                                     ctx->opt_count_visits, ctx->total_rays.p, extra);
                    MRT_LAUNCHED(ctx);
                    ctx->hits_dirty = true;
                }
                if (timed) {
                    cudaEventRecord(ctx->trace_ev[2 * ctx->trace_ev_used + 1], st);
                    ctx->trace_ev_used++;
                }
                if (!fused) {
                    k_shade<false><<<shade_grid, 256, 0, st>>>(sa, in_count);
                    MRT_LAUNCHED(ctx);
                }
                wave++;
                q ^= 1;
            }
        }
        if (bands > 1) {
            MRT_CUDA(ctx, cudaEventRecord(ctx->band_done[band], st));
            MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->band_done[band], 0));
        }
    }
    if (ctx->total_rays.p && waves == 0) {
        k_sum_rays<<<1, 32, 0, ctx->stream>>>(ctx->queue_counts.p, 0, (unsigned long long)npix, ctx->total_rays.p);
        MRT_LAUNCHED(ctx);
    }
    return mrt_check_cuda(ctx, cudaGetLastError(), "mesh_secondary");
}

int mesh_trace_rays(mrt_context* ctx, const float* o, const float* d, uint32_t n, uint32_t* ids, float* t, int brute) {
    if (n == 0) return MRT_OK;
    // context-owned scratch (grown on demand, freed with the context): no per-call cudaMalloc, nothing to leak on an
    // early return
    MRT_TRY(dev_reserve(ctx, ctx->query_o, 3 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->query_d, 3 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->query_t, n));
    MRT_TRY(dev_reserve(ctx, ctx->query_ids, n));
    float *d_o = ctx->query_o.p, *d_d = ctx->query_d.p, *d_t = ctx->query_t.p;
    uint32_t* d_ids = ctx->query_ids.p;
    MRT_CUDA(ctx, cudaMemcpyAsync(d_o, o, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MRT_CUDA(ctx, cudaMemcpyAsync(d_d, d, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MRT_TRY(reserve_visit_counters(ctx));
    MRT_TRY(dev_reserve(ctx, ctx->counters, 16));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->visit_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->counters.p + 8, 0, sizeof(uint32_t), ctx->stream));
    if (brute) {
        k_trace_brute<<<div_up(n, 128), 128, 0, ctx->stream>>>(ctx->pos.p, ctx->idx.p, ctx->ntris, d_o, d_d, n, d_ids, d_t);
    } else {
        QueryJob J{d_o, d_d, n, d_ids, d_t};
        k_trace<QueryJob><<<trace_grid(ctx, n), TRACE_BLOCK, 0, ctx->stream>>>(J, make_bvh(ctx), ctx->counters.p + 8,
                                                                              ctx->visit_counters.p, 1);
    }
    MRT_LAUNCHED(ctx);
    MRT_CUDA(ctx, cudaGetLastError());
    MRT_CUDA(ctx, cudaMemcpyAsync(ids, d_ids, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(ctx, cudaMemcpyAsync(t, d_t, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long vc[3] = {0, 0, 0};
    MRT_CUDA(ctx, cudaMemcpy(vc, ctx->visit_counters.p, sizeof vc, cudaMemcpyDeviceToHost));
    ctx->stats.node_visits = vc[0];
    ctx->stats.tri_tests = vc[1];
    ctx->stats.stack_overflows += (uint32_t)vc[2];
    return MRT_OK;
}
