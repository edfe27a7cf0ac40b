// trace.cuh -- closest-hit traversal of the compressed 8-wide BVH (north_star rows n3, n4).
//
// Execution model for incoherent rays (bounce waves): persistent warps.  Every lane owns one ray; the
// warp runs a warp-synchronous state machine whose step is chosen by ballot:
//     refill  : lanes whose ray finished take the next ray of the warp's chunk (one global atomic per
//               TRACE_CHUNK rays) once TRACE_REFILL_MIN lanes are idle -- bounce rays finish at very
//               different times, refilling keeps the warp populated;
//     node    : lanes with a pending node fetch it (five 128-bit ld.global.nc) and test its 8 children;
//     triangle: lanes with pending leaf triangles test up to TRACE_TRI_PER_STEP of them; triangle steps wait
//               until TRACE_TRI_MIN lanes have one (or no lane has node work), so they run batched.  A lane may
//               hold ONE postponed triangle group and keep taking node steps meanwhile (TRACE_POSTPONE).
// Coherent rays (the primary pass) use the plain per-lane loop trace_coherent().
// History (profiles/r1_results.md): v0 "one node then its triangles per lane" ran bounce rays at 7.8 of 32
// active lanes; the state machine reached ~19; the node encoding below halved the instructions of a node step.
//
// Node step, ~215 SASS instructions, 18 per child:
//   * decode + slab plane in ONE FMA: far planes -- PRMT builds the float 2^15 + q from the quantised byte
//     (0x47000000 | q << 8); near planes -- IDP.4A builds 2^15 + q/2 (0x47000000 + 128 q) so that the 48
//     decodes split between the ALU pipe and the FMA-heavy pipe; t = fma(decoded, step/d, (origin - o)/d -
//     2^15 * step/d).  The constant term carries an error of at most step/256, which the builder's outward
//     rounding (slack in bvh_build.cu k_emit_nodes) covers; near planes use 1/d * (1 - 2^-21), far planes
//     1/d * (1 + 2^-21), so fp32 rounding can never cull a box the exact ray touches;
//   * hit <=> tmin <= tmax && tmin <= tlimit && tmax >= 0: the three sign bits are ORed (one LOP3) and
//     funnel-shifted (SHF.L.W) into an 8-bit miss mask -- no predicates, no branches;
//   * inner hits are permuted into "slot ^ inverse ray octant" order by a 2 KB shared-memory table so that
//     the highest set bit is the child that lies first along the ray; leaf hits are expanded to the node's
//     fixed 3-bits-per-slot triangle mask by a second table (the LSU is idle here, the ALU pipe is not).
// The stack holds (child_base, ordered inner hits << 24 | imask) groups: the first TRACE_SM_STACK entries
// per lane live in shared memory (column layout, conflict-free), deeper ones spill to local memory.
//
// Ray/triangle: Woop-Benthin-Wald watertight test, fp32, no contraction (file built with -fmad=false;
// the only FMAs are the explicit fmaf() of the slab test), double fallback on zero edge functions, no
// culling, accept t >= 0.  Closest hit = lexicographic min of (t, primitive id), the tie rule of
// primaryRay.comp:28.  The axis permutation uses predicated selects (no divergent branches).
#pragma once
#include "context.cuh"

#ifndef TRACE_BLOCK
#define TRACE_BLOCK 128  // threads per traversal CTA
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 6   // resident CTAs per SM the register budget is capped for (80 regs/thread)
#endif
#ifndef TRACE_SM_STACK
#define TRACE_SM_STACK 10    // stack entries per lane kept in shared memory
#endif
#ifndef TRACE_LOCAL_STACK
#define TRACE_LOCAL_STACK 38 // deeper entries spill to local memory
#endif
#ifndef TRACE_CHUNK
#define TRACE_CHUNK 32       // rays a warp takes per global atomic
#endif
#ifndef TRACE_REFILL_MIN
#define TRACE_REFILL_MIN 6   // idle lanes that trigger a refill (2..12 measured; 5..8 flat)
#endif
#ifndef TRACE_TRI_MIN
#define TRACE_TRI_MIN 12     // lanes with pending triangles that trigger a triangle step
#endif
#ifndef TRACE_TRI_PER_STEP
#define TRACE_TRI_PER_STEP 2 // triangles a lane may test in one triangle step (1: lanes with leftovers idle; 3+: long steps)
#endif
#ifndef TRACE_PREFETCH
#define TRACE_PREFETCH 0     // prefetch the next child node to L1 at the end of a node step (A/B measured)
#endif
#ifndef TRACE_DRAIN_PREFETCH
#define TRACE_DRAIN_PREFETCH 0  // prefetch the next child node in the drain phase only (A/B measured: -1.5 %, off)
#endif
#ifndef TRACE_DRAIN_TRI
#define TRACE_DRAIN_TRI 3    // drain phase: a triangle step runs once 1/n of the working lanes want one (measured 2..32: flat)
#endif
#ifndef TRACE_NODE_REPEAT
#define TRACE_NODE_REPEAT 1  // node steps per trip round the persistent loop (A/B: 2)
#endif
#ifndef TRACE_POSTPONE
#define TRACE_POSTPONE 1     // 1: a lane of the persistent loop may hold one postponed triangle group and keep taking node steps
#endif
#ifndef TRACE_NEAREST_FIRST
#define TRACE_NEAREST_FIRST 0  // 1: of a node's hit inner children the one whose box the ray enters first is visited first, the rest in octant order (A/B)
#endif
#ifndef TRACE_DEFER_STORE
#define TRACE_DEFER_STORE 0  // 1: a finished ray's hit record is written when its lane takes the next ray (A/B measured: neutral)
#endif
#ifndef TRACE_DP4A_NEAR
#define TRACE_DP4A_NEAR 1    // decode near planes with IDP.4A (FMA-heavy pipe), far planes with PRMT (ALU pipe)
#endif

struct TraceHit {
    float t;
    uint32_t tri;   // index into BvhDev::tris / 3, or MRT_MISS_ID
    uint32_t prim;  // upload-order primitive id, or MRT_MISS_ID
};

// component k (0,1,2) of (x,y,z) with predicated selects
MRT_D float sel3(float x, float y, float z, int k) {
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\t"
        "selp.f32 %0, %2, %3, p1;\n\tselp.f32 %0, %1, %0, p0;\n\t}"
        : "=f"(r)
        : "f"(x), "f"(y), "f"(z), "r"(k));
    return r;
}

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

MRT_D RayShear make_shear(float3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz == 2 ? 0 : r.kz + 1;
    r.ky = r.kx == 2 ? 0 : r.kx + 1;
    float dz = sel3(d.x, d.y, d.z, r.kz);
    if (dz < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = sel3(d.x, d.y, d.z, r.kx) / dz;
    r.Sy = sel3(d.x, d.y, d.z, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// returns true and t on a hit (t >= 0); barycentrics are not needed by the flat-shaded path
MRT_D bool tri_test(float3 o, const RayShear& rs, float3 p0, float3 p1, float3 p2, float& t) {
    float3 A = p0 - o, B = p1 - o, C = p2 - o;
    float Akz = sel3(A.x, A.y, A.z, rs.kz), Bkz = sel3(B.x, B.y, B.z, rs.kz), Ckz = sel3(C.x, C.y, C.z, rs.kz);
    float Ax = sel3(A.x, A.y, A.z, rs.kx) - rs.Sx * Akz, Ay = sel3(A.x, A.y, A.z, rs.ky) - rs.Sy * Akz;
    float Bx = sel3(B.x, B.y, B.z, rs.kx) - rs.Sx * Bkz, By = sel3(B.x, B.y, B.z, rs.ky) - rs.Sy * Bkz;
    float Cx = sel3(C.x, C.y, C.z, rs.kx) - rs.Sx * Ckz, Cy = sel3(C.x, C.y, C.z, rs.ky) - rs.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;
    float Az = rs.Sz * Akz, Bz = rs.Sz * Bkz, Cz = rs.Sz * Ckz;
    float T = U * Az + V * Bz + W * Cz;
    float tt = T / det;
    if (!(tt >= 0.0f)) return false;
    t = tt;
    return true;
}

MRT_D void hit_consider(TraceHit& h, float t, uint32_t tri, uint32_t prim) {
    if (h.prim == MRT_MISS_ID || t < h.t || (t == h.t && prim < h.prim)) {
        h.t = t; h.tri = tri; h.prim = prim;
    }
}

struct TraceCounters { unsigned nodes, tris, overflow; };

// Per-lane traversal state.
struct LaneState {
    float3 o;
    float3 idn, idf;   // 1/d * (1 - 2^-21) for near planes, 1/d * (1 + 2^-21) for far planes
    RayShear rs;
    unsigned oct_inv;  // 7 - octant; bit a clear <=> direction negative on axis a
    uint2 ng;          // pending node group: child_base, ordered inner hits << 24 | imask
    uint2 tg;          // pending triangle group: tri_base, hit triangle bits
    unsigned tgmask;   // leafmask24 of the node tg came from
    uint2 tg2;         // postponed (older) triangle group (persistent loop with TRACE_POSTPONE only)
    unsigned tg2mask;
    int sp;
    TraceHit hit;
    float tlimit;      // hit.t * (1 + 2^-21)
};

#define TRACE_KNEAR 0.99999952f
#define TRACE_KFAR 1.00000048f

// What lane_begin derives from a ray's direction alone: six IEEE divisions (1/d per axis, the shear) and the octant.  Rays
// the shade stage generates can carry it (option "prepared_rays": computed there by full warps, read back here as two
// 16-byte loads) instead of having it recomputed by the ~7 lanes of a refill.
struct RayPre {
    float3 idir;
    RayShear rs;
    unsigned oct_inv;
};
MRT_D RayPre ray_prepare(float3 d) {
    const float tiny = 1e-20f;
    float3 dd = f3(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x), fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y),
                   fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
    RayPre r;
    r.idir = f3(1.0f / dd.x, 1.0f / dd.y, 1.0f / dd.z);
    r.oct_inv = 7u - ((r.idir.x < 0.0f ? 1u : 0u) | (r.idir.y < 0.0f ? 2u : 0u) | (r.idir.z < 0.0f ? 4u : 0u));
    r.rs = make_shear(d);
    return r;
}
// the two queue words next to origin and direction: (1/d, Sx) and (Sy, Sz, kx | ky << 2 | kz << 4 | oct_inv << 8, -)
MRT_D void ray_pre_pack(const RayPre& r, float4& a, float4& b) {
    a = make_float4(r.idir.x, r.idir.y, r.idir.z, r.rs.Sx);
    b = make_float4(r.rs.Sy, r.rs.Sz, __uint_as_float((unsigned)r.rs.kx | ((unsigned)r.rs.ky << 2) | ((unsigned)r.rs.kz << 4) | (r.oct_inv << 8)), 0.0f);
}
MRT_D RayPre ray_pre_unpack(float4 a, float4 b) {
    RayPre r;
    r.idir = f3(a.x, a.y, a.z);
    const unsigned k = __float_as_uint(b.z);
    r.rs.kx = (int)(k & 3u); r.rs.ky = (int)((k >> 2) & 3u); r.rs.kz = (int)((k >> 4) & 3u);
    r.rs.Sx = a.w; r.rs.Sy = b.x; r.rs.Sz = b.y;
    r.oct_inv = k >> 8;
    return r;
}

MRT_D void lane_begin_prepared(LaneState& L, float3 o, const RayPre& pre) {
    L.o = o;
    L.idn = pre.idir * TRACE_KNEAR;
    L.idf = pre.idir * TRACE_KFAR;
    L.oct_inv = pre.oct_inv;
    L.rs = pre.rs;
    L.ng = make_uint2(0u, 0x80000000u);  // root "group": node 0, one pending inner hit
    L.tg = make_uint2(0u, 0u);
    L.tgmask = 0u;
    L.tg2 = make_uint2(0u, 0u);
    L.tg2mask = 0u;
    L.sp = 0;
    L.hit.t = 3.0e38f; L.hit.tri = MRT_MISS_ID; L.hit.prim = MRT_MISS_ID;
    L.tlimit = 3.0e38f;
}
MRT_D void lane_begin(LaneState& L, float3 o, float3 d) { lane_begin_prepared(L, o, ray_prepare(d)); }

// float 2^15 + (byte `sel` of w).  K = 0x47000000 is passed in a register so that the selector can be the
// immediate operand of PRMT (otherwise ptxas keeps four selectors in uniform registers and copies them).
MRT_D float q_plane(unsigned w, unsigned K, unsigned sel) { return __uint_as_float(__byte_perm(w, K, 0x7404u | (sel << 4))); }
// float 2^15 + (byte `sel` of w) / 2 through the integer dot-product unit (FMA-heavy pipe instead of the
// ALU pipe PRMT runs on): 0x47000000 + 128 * byte.  The node step decodes near planes this way and far
// planes with PRMT so that neither pipe carries all 48 decodes (ncu: ALU pipe 65 % busy, FMA 22 %).
MRT_D float q_plane_half(unsigned w, unsigned K, unsigned sel) { return __uint_as_float(__dp4a(w, 128u << (8 * sel), K)); }

#ifndef TRACE_FFMA2
#define TRACE_FFMA2 1  // slab planes of two children per FFMA2 / FADD2 (packed fp32, sm_100)
#endif
MRT_D unsigned long long pack2(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
MRT_D float2 fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    float2 o;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}
MRT_D float2 add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    float2 o;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}

// Per-CTA shared memory of the traversal kernels: the lanes' stack columns and two bit-shuffle tables.
#ifndef TRACE_COOP_TRI
#define TRACE_COOP_TRI 0     // triangle steps of the persistent loop: pending (ray, triangle) pairs spread over ALL lanes of the warp
#endif
struct TraceShared {
#if TRACE_COOP_TRI
    uint2 pairs[TRACE_BLOCK / 32][64];   // per warp: (triangle, owner lane) of the pairs of one triangle step
#endif
    uint2 stack[TRACE_SM_STACK][TRACE_BLOCK];
    uint32_t expand3[256];        // bit j -> bits 3j..3j+2
    unsigned char perm[8][256];   // perm[c][x]: bit (s ^ c) = bit s of x
};

// The tables are compile-time constants in global memory; each CTA copies them (3 KB) into shared memory.
struct TraceTables {
    uint32_t expand3[256];
    unsigned char perm[8][256];
    constexpr TraceTables() : expand3(), perm() {
        for (unsigned x = 0; x < 256u; x++) {
            unsigned e = 0;
            for (unsigned j = 0; j < 8; j++)
                if (x & (1u << j)) e |= 7u << (3 * j);
            expand3[x] = e;
            for (unsigned c = 0; c < 8; c++) {
                unsigned p = 0;
                for (unsigned j = 0; j < 8; j++)
                    if (x & (1u << j)) p |= 1u << (j ^ c);
                perm[c][x] = (unsigned char)p;
            }
        }
    }
};
static __device__ const TraceTables g_trace_tables = TraceTables();

MRT_D void trace_shared_init(TraceShared& S) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&g_trace_tables);
    uint32_t* dst = S.expand3;  // expand3[256] and perm[8][256] are contiguous in both structs
    static_assert(sizeof(TraceTables) == 3072, "table layout");
    for (unsigned i = threadIdx.x; i < sizeof(TraceTables) / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// One node step: take the nearest pending child of L.ng, fetch it, test its 8 children.
// POSTPONE: the lane may arrive with untested triangles in L.tg; they move to the (free) postponed slot.
template <bool POSTPONE>
MRT_D void lane_node_step(LaneState& L, const BvhDev& bvh, TraceShared& S, uint2* spill, TraceCounters& cnt, bool prefetch = false) {
    uint2* const sm = &S.stack[0][threadIdx.x];
    if (POSTPONE && L.tg.y) { L.tg2 = L.tg; L.tg2mask = L.tgmask; }
#if TRACE_NEAREST_FIRST
    // bits 8..11 of ng.y: octant-order position + 1 of the child whose box the ray enters first (set by the step that made the group)
    const unsigned first = (L.ng.y >> 8) & 15u;
    const unsigned bit = first ? 23u + first : 31u - __clz(L.ng.y);
    L.ng.y &= ~((1u << bit) | 0xF00u);
#else
    const unsigned bit = 31u - __clz(L.ng.y);
    L.ng.y &= ~(1u << bit);
#endif
    if (L.ng.y & 0xFF000000u) {  // siblings remain: keep the group for later
        if (L.sp < TRACE_SM_STACK) sm[L.sp * TRACE_BLOCK] = L.ng;
        else if (L.sp < TRACE_SM_STACK + TRACE_LOCAL_STACK) spill[L.sp - TRACE_SM_STACK] = L.ng;
        else cnt.overflow++;
        L.sp = min(L.sp + 1, TRACE_SM_STACK + TRACE_LOCAL_STACK);
    }
    const unsigned slot = (bit - 24u) ^ L.oct_inv;
    const unsigned rel = __popc(L.ng.y & ~(0xFFFFFFFFu << slot));
    const uint4* np = reinterpret_cast<const uint4*>(bvh.nodes + (L.ng.x + rel));
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    cnt.nodes++;

    // per-axis slope and offset of  t(q) = (origin + q * step - o) / d  evaluated at the float 2^15 + q
    const float stx = __uint_as_float((n0.w & 0xFFu) << 23), sty = __uint_as_float((n0.w & 0xFF00u) << 15),
                stz = __uint_as_float((n0.w & 0xFF0000u) << 7);
    const float dx = __uint_as_float(n0.x) - L.o.x, dy = __uint_as_float(n0.y) - L.o.y, dz = __uint_as_float(n0.z) - L.o.z;
    const float snx = stx * L.idn.x, sny = sty * L.idn.y, snz = stz * L.idn.z;
    const float sfx = stx * L.idf.x, sfy = sty * L.idf.y, sfz = stz * L.idf.z;
#if TRACE_DP4A_NEAR
    // near planes are decoded as 2^15 + q/2: slope 2 * step/d, offset -2^16 * step/d (error <= step/256)
    const float bnx = fmaf(-65536.0f, snx, dx * L.idn.x), bny = fmaf(-65536.0f, sny, dy * L.idn.y),
                bnz = fmaf(-65536.0f, snz, dz * L.idn.z);
    const float snx2 = snx * 2.0f, sny2 = sny * 2.0f, snz2 = snz * 2.0f;
#else
    const float bnx = fmaf(-32768.0f, snx, dx * L.idn.x), bny = fmaf(-32768.0f, sny, dy * L.idn.y),
                bnz = fmaf(-32768.0f, snz, dz * L.idn.z);
#endif
    const float bfx = fmaf(-32768.0f, sfx, dx * L.idf.x), bfy = fmaf(-32768.0f, sfy, dy * L.idf.y),
                bfz = fmaf(-32768.0f, sfz, dz * L.idf.z);
    // near / far quantised planes per axis, chosen by the ray octant
    const bool px = L.oct_inv & 1u, py = L.oct_inv & 2u, pz = L.oct_inv & 4u;  // direction positive on the axis
    const unsigned nx0 = px ? n2.x : n3.z, nx1 = px ? n2.y : n3.w, fx0 = px ? n3.z : n2.x, fx1 = px ? n3.w : n2.y;
    const unsigned ny0 = py ? n2.z : n4.x, ny1 = py ? n2.w : n4.y, fy0 = py ? n4.x : n2.z, fy1 = py ? n4.y : n2.w;
    const unsigned nz0 = pz ? n3.x : n4.z, nz1 = pz ? n3.y : n4.w, fz0 = pz ? n4.z : n3.x, fz1 = pz ? n4.w : n3.y;
    const float tlimit = L.tlimit;

    const unsigned K = bvh.prmt_k;  // 0x47000000 from the kernel parameters (a constant-bank operand of PRMT)
    unsigned miss = 0u;  // after the loop: bit j set <=> child j missed
#if defined(TRACE_COUNT_EMPTY) && TRACE_COUNT_EMPTY == 2
    unsigned miss_nolimit = 0u;
#endif
#if TRACE_FFMA2
    // Two children per instruction: Blackwell's packed fp32 pipe (FFMA2 / FADD2, PTX fma.rn.f32x2) evaluates the slab
    // planes of children j and j-1 in one issue slot each -- the kernel is issue-bound (ncu: 70 % of the issue slots, FMA
    // pipe 25 %), so halving the FMA and FADD instruction count of the node step is worth the register pairing.  Every
    // component is the same IEEE fused multiply-add as the scalar code: results are bit-identical.
    // near planes are evaluated NEGATED (slope and offset negated: fma(q, -s, -b) = -fma(q, s, b) exactly), so that
    // -tmin = min of the three and the hit test's subtractions become packed additions
    const unsigned long long sn2x = pack2(-snx2, -snx2), sn2y = pack2(-sny2, -sny2), sn2z = pack2(-snz2, -snz2);
    const unsigned long long bn2x = pack2(-bnx, -bnx), bn2y = pack2(-bny, -bny), bn2z = pack2(-bnz, -bnz);
    const unsigned long long sf2x = pack2(sfx, sfx), sf2y = pack2(sfy, sfy), sf2z = pack2(sfz, sfz);
    const unsigned long long bf2x = pack2(bfx, bfx), bf2y = pack2(bfy, bfy), bf2z = pack2(bfz, bfz);
    const unsigned long long tl2 = pack2(tlimit, tlimit);
#if TRACE_NEAREST_FIRST
    float nearest_key = -3.4e38f;
    const unsigned imask_early = n0.w >> 24;
#endif
#pragma unroll
    for (int j = 7; j >= 1; j -= 2) {
        const unsigned wnx = j < 4 ? nx0 : nx1, wny = j < 4 ? ny0 : ny1, wnz = j < 4 ? nz0 : nz1;
        const unsigned wfx = j < 4 ? fx0 : fx1, wfy = j < 4 ? fy0 : fy1, wfz = j < 4 ? fz0 : fz1;
        const int a = j & 3, b = (j - 1) & 3;  // child j in the high half... (x = child j, y = child j-1)
        float2 t0x = fma2(pack2(q_plane_half(wnx, K, a), q_plane_half(wnx, K, b)), sn2x, bn2x);
        float2 t0y = fma2(pack2(q_plane_half(wny, K, a), q_plane_half(wny, K, b)), sn2y, bn2y);
        float2 t0z = fma2(pack2(q_plane_half(wnz, K, a), q_plane_half(wnz, K, b)), sn2z, bn2z);
        float2 t1x = fma2(pack2(q_plane(wfx, K, a), q_plane(wfx, K, b)), sf2x, bf2x);
        float2 t1y = fma2(pack2(q_plane(wfy, K, a), q_plane(wfy, K, b)), sf2y, bf2y);
        float2 t1z = fma2(pack2(q_plane(wfz, K, a), q_plane(wfz, K, b)), sf2z, bf2z);
        const float ntminA = fminf(fminf(t0x.x, t0y.x), t0z.x), tmaxA = fminf(fminf(t1x.x, t1y.x), t1z.x);  // -tmin, tmax
        const float ntminB = fminf(fminf(t0x.y, t0y.y), t0z.y), tmaxB = fminf(fminf(t1x.y, t1y.y), t1z.y);
        const unsigned long long tmin2 = pack2(ntminA, ntminB);
        const float2 d1 = add2(pack2(tmaxA, tmaxB), tmin2), d2 = add2(tl2, tmin2);  // tmax - tmin, tlimit - tmin
        const unsigned negA = __float_as_uint(d1.x) | __float_as_uint(d2.x) | __float_as_uint(tmaxA);
        const unsigned negB = __float_as_uint(d1.y) | __float_as_uint(d2.y) | __float_as_uint(tmaxB);
        miss = __funnelshift_l(negA, miss, 1);
        miss = __funnelshift_l(negB, miss, 1);
#if defined(TRACE_COUNT_EMPTY) && TRACE_COUNT_EMPTY == 2  // diagnostic: the same test without the closest-hit limit
        miss_nolimit = __funnelshift_l(__float_as_uint(d1.x) | __float_as_uint(tmaxA), miss_nolimit, 1);
        miss_nolimit = __funnelshift_l(__float_as_uint(d1.y) | __float_as_uint(tmaxB), miss_nolimit, 1);
#endif
#if TRACE_NEAREST_FIRST
        {   // largest -tmin among the inner children that are hit; the slot rides in the three low mantissa bits
            const float kA = __uint_as_float((__float_as_uint(ntminA) & ~7u) | (unsigned)j);
            const float kB = __uint_as_float((__float_as_uint(ntminB) & ~7u) | (unsigned)(j - 1));
            const bool okA = !(negA >> 31) && ((imask_early >> j) & 1u), okB = !(negB >> 31) && ((imask_early >> (j - 1)) & 1u);
            nearest_key = fmaxf(nearest_key, fmaxf(okA ? kA : -3.4e38f, okB ? kB : -3.4e38f));
        }
#endif
    }
#else
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        const unsigned wnx = j < 4 ? nx0 : nx1, wny = j < 4 ? ny0 : ny1, wnz = j < 4 ? nz0 : nz1;
        const unsigned wfx = j < 4 ? fx0 : fx1, wfy = j < 4 ? fy0 : fy1, wfz = j < 4 ? fz0 : fz1;
#if TRACE_DP4A_NEAR
        const float t0x = fmaf(q_plane_half(wnx, K, j & 3), snx2, bnx), t0y = fmaf(q_plane_half(wny, K, j & 3), sny2, bny),
                    t0z = fmaf(q_plane_half(wnz, K, j & 3), snz2, bnz);
#else
        const float t0x = fmaf(q_plane(wnx, K, j & 3), snx, bnx), t0y = fmaf(q_plane(wny, K, j & 3), sny, bny),
                    t0z = fmaf(q_plane(wnz, K, j & 3), snz, bnz);
#endif
        const float t1x = fmaf(q_plane(wfx, K, j & 3), sfx, bfx), t1y = fmaf(q_plane(wfy, K, j & 3), sfy, bfy),
                    t1z = fmaf(q_plane(wfz, K, j & 3), sfz, bfz);
        // hit <=> tmin <= tmax && tmin <= tlimit && tmax >= 0: OR the three sign bits (one LOP3) instead of
        // clamping with two more FMNMX -- the subtractions run on the under-used FMA pipe
        const float tmin = fmaxf(fmaxf(t0x, t0y), t0z);
        const float tmax = fminf(fminf(t1x, t1y), t1z);
        const unsigned neg = __float_as_uint(tmax - tmin) | __float_as_uint(tlimit - tmin) | __float_as_uint(tmax);
        miss = __funnelshift_l(neg, miss, 1);  // (miss << 1) | any sign set
    }
#endif
    const unsigned imask = n0.w >> 24;
    const unsigned hit8 = ~miss & 0xFFu;
#ifdef TRACE_COUNT_EMPTY  // diagnostic build: node steps that hit no child are counted as "stack overflows" (tools/count_empty.py)
#if TRACE_COUNT_EMPTY == 2    // ... only those that would have hit a child without the closest-hit limit: stale stack entries
    if (hit8 == 0u && (~miss_nolimit & 0xFFu) != 0u) cnt.overflow++;
#else
    if (hit8 == 0u) cnt.overflow++;
#endif
#endif
    // inner hits -> bit (slot ^ oct_inv); leaf hits -> 3 bits per slot, masked by the triangles that exist.
    // Both are 256-entry shared-memory tables (the LSU is idle here, the ALU pipe is the bottleneck).
    const unsigned inner = S.perm[L.oct_inv][hit8 & imask];
    const unsigned leaf = S.expand3[hit8 & ~imask];
#if TRACE_NEAREST_FIRST && TRACE_FFMA2
    {
        const unsigned pos = ((__float_as_uint(nearest_key) & 7u) ^ L.oct_inv) + 1u;  // octant-order position of the nearest child
        L.ng = make_uint2(n1.x, (inner << 24) | imask | (nearest_key > -3.0e38f ? pos << 8 : 0u));
    }
#else
    L.ng = make_uint2(n1.x, (inner << 24) | imask);
#endif
    L.tg = make_uint2(n1.y, leaf & n1.z);
    L.tgmask = n1.z;
    // The child this lane visits next is already known: pull its 80 bytes towards L1 while the warp does its triangle
    // step / loop bookkeeping.  In the body of a launch this costs more issue slots than it hides latency (+5 %, the
    // kernel is issue-bound there); in the DRAIN phase (queue exhausted, a few thin warps per SM, every step waits out
    // the full L2 latency) it is the latency that counts: TRACE_DRAIN_PREFETCH passes prefetch = true then.
    if ((TRACE_PREFETCH || prefetch) && inner) {
        const unsigned nbit = 31u - __clz(inner << 24);
        const unsigned nslot = (nbit - 24u) ^ L.oct_inv;
        const unsigned nrel = __popc(imask & ~(0xFFFFFFFFu << nslot));
        const char* p = reinterpret_cast<const char*>(bvh.nodes + (n1.x + nrel));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 64));
    }
}

// One triangle of the lane's pending group (POSTPONE: of the older group first).
template <bool POSTPONE>
MRT_D void lane_tri_step(LaneState& L, const BvhDev& bvh, TraceCounters& cnt) {
    uint32_t tri;
    if (POSTPONE) {
        const bool old = L.tg2.y != 0u;
        const unsigned bits = old ? L.tg2.y : L.tg.y;
        const unsigned bit = __ffs(bits) - 1;
        tri = (old ? L.tg2.x : L.tg.x) + __popc((old ? L.tg2mask : L.tgmask) & ((1u << bit) - 1u));
        if (old) L.tg2.y = bits & (bits - 1);
        else L.tg.y = bits & (bits - 1);
    } else {
        const unsigned bit = __ffs(L.tg.y) - 1;
        L.tg.y &= L.tg.y - 1;
        tri = L.tg.x + __popc(L.tgmask & ((1u << bit) - 1u));
    }
    const float4* tp = bvh.tris + 3 * (size_t)tri;
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    cnt.tris++;
    float t;
    if (tri_test(L.o, L.rs, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), t)) {
        hit_consider(L.hit, t, tri, __float_as_uint(v0.w));
        L.tlimit = L.hit.t * TRACE_KFAR;
    }
}

#if TRACE_COOP_TRI
// Cooperative triangle step (all 32 lanes call it).  ncu: triangle steps are 22 % of the bounce kernel's issue slots and run
// with ~11 of 32 lanes -- the lanes that hold pending triangles test up to two of them one after the other while the rest of
// the warp waits.  Here every lane with pending triangles contributes up to two (ray, triangle) PAIRS; the pairs are numbered
// with two ballots, published through shared memory, and lane w tests pair w: it fetches the owner's ray (origin, shear) with
// shuffles, runs the same watertight test, and the owner reads the results of its pairs back with shuffles and merges them
// with hit_consider -- closest hit is the lexicographic minimum of (t, primitive id), so who tested what does not matter.
// ~20 pairs from ~12 lanes become ONE test round with ~20 lanes instead of two rounds with 12 and 9.
MRT_D void warp_tri_step_coop(LaneState& L, const BvhDev& bvh, TraceCounters& cnt, bool want_tri, uint2* pairs) {
    const unsigned FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    uint32_t mine[2] = {0u, 0u};
    unsigned n = 0;
    if (want_tri) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (k == 1 && !(L.tg.y | L.tg2.y)) break;
            const bool old = L.tg2.y != 0u;
            const unsigned bits = old ? L.tg2.y : L.tg.y;
            const unsigned bit = __ffs(bits) - 1;
            mine[k] = (old ? L.tg2.x : L.tg.x) + __popc((old ? L.tg2mask : L.tgmask) & ((1u << bit) - 1u));
            if (old) L.tg2.y = bits & (bits - 1);
            else L.tg.y = bits & (bits - 1);
            n = k + 1;
        }
    }
    const unsigned b1 = __ballot_sync(FULL, n >= 1), b2 = __ballot_sync(FULL, n == 2);
    const unsigned pre = __popc(b1 & lt) + __popc(b2 & lt);
    const unsigned P = __popc(b1) + __popc(b2);
    if (n >= 1) pairs[pre] = make_uint2(mine[0], lane);
    if (n == 2) pairs[pre + 1] = make_uint2(mine[1], lane);
    __syncwarp();
    const int kpack = L.rs.kx | (L.rs.ky << 2) | (L.rs.kz << 4);
#pragma unroll 1
    for (unsigned base = 0; base < P; base += 32) {  // one round unless more than 32 pairs are pending
        const unsigned w = base + lane;
        const bool work = w < P;
        uint2 pr = make_uint2(0u, lane);
        if (work) pr = pairs[w];
        const float3 o = f3(__shfl_sync(FULL, L.o.x, pr.y), __shfl_sync(FULL, L.o.y, pr.y), __shfl_sync(FULL, L.o.z, pr.y));
        RayShear rs;
        rs.Sx = __shfl_sync(FULL, L.rs.Sx, pr.y);
        rs.Sy = __shfl_sync(FULL, L.rs.Sy, pr.y);
        rs.Sz = __shfl_sync(FULL, L.rs.Sz, pr.y);
        const int kp = __shfl_sync(FULL, kpack, pr.y);
        rs.kx = kp & 3; rs.ky = (kp >> 2) & 3; rs.kz = kp >> 4;
        float t = 0.0f;
        uint32_t prim = MRT_MISS_ID;
        if (work) {
            const float4* tp = bvh.tris + 3 * (size_t)pr.x;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            cnt.tris++;
            if (tri_test(o, rs, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), t)) prim = __float_as_uint(v0.w);
        }
        // results travel back to the owners (slots pre, pre + 1 if they fall into this round)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned slot = pre + k - base;
            const float tk = __shfl_sync(FULL, t, slot & 31u);
            const uint32_t pk = __shfl_sync(FULL, prim, slot & 31u);
            if ((unsigned)k < n && slot < 32u && pk != MRT_MISS_ID) hit_consider(L.hit, tk, mine[k], pk);
        }
    }
    __syncwarp();  // the pair list is rewritten by the next triangle step
    if (L.hit.prim != MRT_MISS_ID) L.tlimit = L.hit.t * TRACE_KFAR;
}
#endif

// Persistent warp loop.  Job supplies the rays and consumes the hits:
//   uint32_t Job::count() const;                          rays in this wave
//   bool     Job::load(uint32_t i, float3& o, float3& d); false => ray i does not exist (padding)
//   bool     Job::load_prepared(uint32_t i, float3& o, RayPre& pre);  true => the queue carries ray_prepare's results for ray i
//   void     Job::store(uint32_t i, const TraceHit& h);
//   static constexpr bool Job::ANY_HIT;                   true: occlusion query, the ray ends at its first hit
// work_counter: global counter of handed-out rays (zeroed before the launch).
template <class Job>
MRT_D void trace_persistent(const BvhDev& bvh, Job& job, uint32_t* work_counter, TraceShared& S, TraceCounters& cnt) {
    uint2* const sm = &S.stack[0][threadIdx.x];
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t total = job.count();
    uint2 spill[TRACE_LOCAL_STACK];
    LaneState L;
    L.ng = L.tg = L.tg2 = make_uint2(0u, 0u);
    L.sp = 0;
    bool have_ray = false;
    uint32_t ray_index = 0;
    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform chunk of ray indices
    bool exhausted = false;
#if TRACE_DEFER_STORE
    bool store_pending = false;
#endif

    for (;;) {
        // ---- refill idle lanes from the warp's chunk
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have_ray);
        if (idle && !exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= TRACE_REFILL_MIN)) {
            if (pool_next >= pool_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, (uint32_t)TRACE_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                pool_next = base;
                pool_end = min(base + (uint32_t)TRACE_CHUNK, total);
                if (base >= total) { exhausted = true; pool_next = pool_end = 0; }
            }
            if (!exhausted) {
                const uint32_t mine = pool_next + __popc(idle & lt_mask);
                if (!have_ray && mine < pool_end) {
                    float3 o, d;
#if TRACE_DEFER_STORE
                    if (store_pending) { job.store(ray_index, L.hit); store_pending = false; }  // the lane's previous ray
#endif
                    ray_index = mine;
                    RayPre pre;
                    if (job.load_prepared(mine, o, pre)) {  // (a job without prepared rays returns false here, at compile time)
                        lane_begin_prepared(L, o, pre);
                        if (bvh.num_nodes == 0) L.ng.y = 0u;
                        have_ray = true;
                    } else if (job.load(mine, o, d)) {
                        lane_begin(L, o, d);
                        if (bvh.num_nodes == 0) L.ng.y = 0u;  // empty scene: finishes as a miss below
                        have_ray = true;
                    }
                }
                pool_next = min(pool_next + (uint32_t)__popc(idle), pool_end);
            }
        }
        if (exhausted && __ballot_sync(0xFFFFFFFFu, have_ray) == 0u) break;

        // ---- lanes with a ray but no pending work: pop, or finish the ray
#if TRACE_POSTPONE
        // a lane with a free postponed slot keeps walking nodes (pops included); it finishes only with no work left
        if (have_ray && !(L.ng.y & 0xFF000000u) && !(L.tg.y && L.tg2.y)) {
            if (L.sp == 0) {
                if (!(L.tg.y | L.tg2.y)) {
#if TRACE_DEFER_STORE
                    store_pending = true;  // written when the lane takes its next ray (or after the loop): rays end one or
                                           // two lanes at a time, refills come ~7 lanes at a time
#else
                    job.store(ray_index, L.hit);
#endif
                    have_ray = false;
                }
            } else {
                L.sp--;
                L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
            }
        }
        const bool want_tri = have_ray && (L.tg.y | L.tg2.y) != 0u;
        const bool want_node = have_ray && !(L.tg.y && L.tg2.y) && (L.ng.y & 0xFF000000u);
#else
        if (have_ray && !(L.ng.y & 0xFF000000u) && L.tg.y == 0u) {
            if (L.sp == 0) {
#if TRACE_DEFER_STORE
                store_pending = true;
#else
                job.store(ray_index, L.hit);
#endif
                have_ray = false;
            } else {
                L.sp--;
                L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
            }
        }
        const bool want_tri = have_ray && L.tg.y != 0u;
        const bool want_node = have_ray && !want_tri && (L.ng.y & 0xFF000000u);
#endif
        const unsigned tmask = __ballot_sync(0xFFFFFFFFu, want_tri);
        const unsigned nmask = __ballot_sync(0xFFFFFFFFu, want_node);
        // drain phase (queue empty, the warp thins out): the fixed batch size would make the triangle lanes wait
        // for every node lane, so a triangle step runs once a third of the lanes still working want one
        const int tri_min = exhausted ? min(TRACE_TRI_MIN, max(1, (__popc(tmask | nmask) + TRACE_DRAIN_TRI - 1) / TRACE_DRAIN_TRI)) : TRACE_TRI_MIN;
        if (tmask && (nmask == 0u || __popc(tmask) >= tri_min)) {
#if TRACE_COOP_TRI && TRACE_POSTPONE
            warp_tri_step_coop(L, bvh, cnt, want_tri, S.pairs[threadIdx.x >> 5]);
            if (Job::ANY_HIT && L.hit.prim != MRT_MISS_ID) {  // shadow rays: the first hit ends the ray
                L.ng.y = 0u; L.tg.y = 0u; L.tg2.y = 0u; L.sp = 0;
            }
#else
            if (want_tri) {
                constexpr bool PP = TRACE_POSTPONE != 0;
                lane_tri_step<PP>(L, bvh, cnt);
#if TRACE_TRI_PER_STEP > 1
#pragma unroll 1
                for (int k = 1; k < TRACE_TRI_PER_STEP && (L.tg.y | (PP ? L.tg2.y : 0u)); k++) lane_tri_step<PP>(L, bvh, cnt);
#endif
                if (Job::ANY_HIT && L.hit.prim != MRT_MISS_ID) {  // shadow rays: the first hit ends the ray
                    L.ng.y = 0u; L.tg.y = 0u; L.tg2.y = 0u; L.sp = 0;
                }
            }
#endif
        } else if (nmask) {
#if TRACE_NODE_REPEAT > 1 && TRACE_POSTPONE
            // several node steps per trip round the loop: the refill check, the two ballots and the step choice (~35 warp
            // instructions with all 32 lanes) are paid once per TRACE_NODE_REPEAT node steps; a lane takes the next step
            // (popping its stack if the group is used up) as long as it has a free triangle slot
            if (want_node) {
                lane_node_step<true>(L, bvh, S, spill, cnt, TRACE_DRAIN_PREFETCH && exhausted);
#pragma unroll 1
                for (int rep = 1; rep < TRACE_NODE_REPEAT; rep++) {
                    if (L.tg.y && L.tg2.y) break;
                    if (!(L.ng.y & 0xFF000000u)) {
                        if (L.sp == 0) break;
                        L.sp--;
                        L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
                    }
                    lane_node_step<true>(L, bvh, S, spill, cnt, TRACE_DRAIN_PREFETCH && exhausted);
                }
            }
#else
            if (want_node) lane_node_step<TRACE_POSTPONE != 0>(L, bvh, S, spill, cnt, TRACE_DRAIN_PREFETCH && exhausted);
#endif
        }
    }
#if TRACE_DEFER_STORE
    if (store_pending) job.store(ray_index, L.hit);
#endif
}

// Per-lane loop for COHERENT rays (the primary pass): one node, then its triangles, per iteration.
// Rays of an 8x4 pixel tile walk the same nodes and reach leaves together, so the warp stays
// converged without the state machine, and testing triangles at once tightens t early.
template <bool ANY_HIT = false>
MRT_D TraceHit trace_coherent(const BvhDev& bvh, float3 o, float3 d, TraceShared& S, TraceCounters& cnt) {
    uint2* const sm = &S.stack[0][threadIdx.x];
    uint2 spill[TRACE_LOCAL_STACK];
    LaneState L;
    lane_begin(L, o, d);
    if (bvh.num_nodes == 0) return L.hit;
    for (;;) {
        if (L.ng.y & 0xFF000000u) lane_node_step<false>(L, bvh, S, spill, cnt);
        while (L.tg.y) {
            lane_tri_step<false>(L, bvh, cnt);
            if (ANY_HIT && L.hit.prim != MRT_MISS_ID) return L.hit;  // occlusion query: the first hit ends the ray
        }
        if (!(L.ng.y & 0xFF000000u)) {
            if (L.sp == 0) break;
            L.sp--;
            L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
        }
    }
    return L.hit;
}

// Coherent rays through the warp-synchronous step machine, WITHOUT refill: the 32 rays of a tile stay together (that is
// what makes them coherent), but a lane keeps its leaf triangles pending (one postponed group, as in trace_persistent)
// until a share of the working lanes wants a triangle step.  In the per-lane loop above the triangle tests run right
// after the node step that found them -- with ~10 of 32 lanes (ncu: 27 % of the primary pass' issue slots).
// Must be called by all 32 lanes of the warp; `active` = this lane has a ray.
#ifndef COHERENT_TRI_DIV
#define COHERENT_TRI_DIV 2   // a triangle step runs once 1/n of the lanes still working want one
#endif
MRT_D TraceHit trace_coherent_batched(const BvhDev& bvh, bool active, float3 o, float3 d, TraceShared& S, TraceCounters& cnt) {
    uint2* const sm = &S.stack[0][threadIdx.x];
    uint2 spill[TRACE_LOCAL_STACK];
    LaneState L;
    lane_begin(L, o, d);
    bool have_ray = active && bvh.num_nodes != 0;
    for (;;) {
        if (have_ray && !(L.ng.y & 0xFF000000u) && !(L.tg.y && L.tg2.y)) {
            if (L.sp == 0) {
                if (!(L.tg.y | L.tg2.y)) have_ray = false;
            } else {
                L.sp--;
                L.ng = L.sp < TRACE_SM_STACK ? sm[L.sp * TRACE_BLOCK] : spill[L.sp - TRACE_SM_STACK];
            }
        }
        const bool want_tri = have_ray && (L.tg.y | L.tg2.y) != 0u;
        const bool want_node = have_ray && !(L.tg.y && L.tg2.y) && (L.ng.y & 0xFF000000u);
        const unsigned tmask = __ballot_sync(0xFFFFFFFFu, want_tri);
        const unsigned nmask = __ballot_sync(0xFFFFFFFFu, want_node);
        if (!(tmask | nmask)) break;
        const int tri_min = max(1, (__popc(tmask | nmask) + COHERENT_TRI_DIV - 1) / COHERENT_TRI_DIV);
        if (tmask && (nmask == 0u || __popc(tmask) >= tri_min)) {
            if (want_tri) {
                lane_tri_step<true>(L, bvh, cnt);
#pragma unroll 1
                for (int k = 1; k < TRACE_TRI_PER_STEP && (L.tg.y | L.tg2.y); k++) lane_tri_step<true>(L, bvh, cnt);
            }
        } else {
            if (want_node) lane_node_step<true>(L, bvh, S, spill, cnt);
        }
    }
    return L.hit;
}

// geometric normal of a hit triangle, flipped to face the incoming ray (two-sided surfaces)
MRT_D float3 tri_facing_normal(const BvhDev& bvh, uint32_t tri, float3 d, uint32_t* prim_out) {
    const float4* tp = bvh.tris + 3 * (size_t)tri;
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    float3 p0 = f3(v0.x, v0.y, v0.z);
    float3 n = normalize3(cross3(f3(v1.x, v1.y, v1.z) - p0, f3(v2.x, v2.y, v2.z) - p0));
    if (dot3(n, d) > 0.0f) n = -n;
    if (prim_out) *prim_out = __float_as_uint(v0.w);
    return n;
}
