// trace.cuh -- closest-hit traversal of the compressed 8-wide BVH (north_star rows n3, n4).
//
// Node fetch: five 128-bit read-only loads (ld.global.nc.v4) per 80-byte node.  Child boxes are
// decoded with one PRMT + FADD per byte (0x4B000000 | q is the float 2^23 + q) and tested with one
// FMA per slab plane against the pre-scaled inverse direction.  Hits are gathered into a 32-bit
// mask: inner children land in bits 24..31 at position (slot ^ inverse ray octant), so the highest
// set bit is the child that lies first along the ray; leaf triangles land in bits 0..23.
// The traversal stack holds (child_base, hit bits | imask) groups: the first TRACE_SM_STACK entries
// per lane live in shared memory (column layout, conflict-free), deeper ones spill to local memory.
//
// Ray/triangle: Woop-Benthin-Wald watertight test, fp32, no contraction (file built with
// -fmad=false; the only FMAs are the explicit fmaf() of the slab test), double fallback on zero
// edge functions, no culling, accept t >= 0.  Closest hit = lexicographic min of (t, primitive id),
// the tie rule of primaryRay.comp:28.
#pragma once
#include "context.cuh"

#define TRACE_BLOCK 128
#define TRACE_SM_STACK 12
#define TRACE_LOCAL_STACK 36

struct TraceHit {
    float t;
    uint32_t tri;   // index into BvhDev::tris / 3, or MRT_MISS_ID
    float u, v;
    uint32_t prim;  // upload-order primitive id, or MRT_MISS_ID
};

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

MRT_D RayShear make_shear(float3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz == 2 ? 0 : r.kz + 1;
    r.ky = r.kx == 2 ? 0 : r.kx + 1;
    float dz = comp3(d, r.kz);
    if (dz < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = comp3(d, r.kx) / dz;
    r.Sy = comp3(d, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

MRT_D bool tri_test(float3 o, const RayShear& rs, float3 p0, float3 p1, float3 p2, float& t, float& u, float& v) {
    float3 A = p0 - o, B = p1 - o, C = p2 - o;
    float Akz = comp3(A, rs.kz), Bkz = comp3(B, rs.kz), Ckz = comp3(C, rs.kz);
    float Ax = comp3(A, rs.kx) - rs.Sx * Akz, Ay = comp3(A, rs.ky) - rs.Sy * Akz;
    float Bx = comp3(B, rs.kx) - rs.Sx * Bkz, By = comp3(B, rs.ky) - rs.Sy * Bkz;
    float Cx = comp3(C, rs.kx) - rs.Sx * Ckz, Cy = comp3(C, rs.ky) - rs.Sy * Ckz;
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
    u = V / det;
    v = W / det;
    return true;
}

MRT_D void hit_consider(TraceHit& h, float t, float u, float v, uint32_t tri, uint32_t prim) {
    if (h.prim == MRT_MISS_ID || t < h.t || (t == h.t && prim < h.prim)) {
        h.t = t; h.u = u; h.v = v; h.tri = tri; h.prim = prim;
    }
}

MRT_D float q_to_float(unsigned w, unsigned sel) {
    // byte `sel` of w -> float, via the 2^23 + q bit pattern
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | sel)) - 8388608.0f;
}

// stack: per-lane column in shared memory (stride TRACE_BLOCK) + local spill
struct TraceStack {
    uint2* sm;  // &shared[0][lane column]
    uint2 spill[TRACE_LOCAL_STACK];
    int sp;
};

MRT_D void stack_push(TraceStack& S, uint2 e, unsigned& overflow) {
    if (S.sp < TRACE_SM_STACK) S.sm[S.sp * TRACE_BLOCK] = e;
    else if (S.sp < TRACE_SM_STACK + TRACE_LOCAL_STACK) S.spill[S.sp - TRACE_SM_STACK] = e;
    else { overflow++; return; }
    S.sp++;
}
MRT_D uint2 stack_pop(TraceStack& S) {
    S.sp--;
    return S.sp < TRACE_SM_STACK ? S.sm[S.sp * TRACE_BLOCK] : S.spill[S.sp - TRACE_SM_STACK];
}

struct TraceCounters { unsigned nodes, tris, overflow; };

// sm_column: this lane's column of a __shared__ uint2[TRACE_SM_STACK][TRACE_BLOCK] array
MRT_D TraceHit bvh_trace(const BvhDev& bvh, float3 o, float3 d, uint2* sm_column, TraceCounters& cnt) {
    TraceHit hit;
    hit.t = 3.0e38f; hit.tri = MRT_MISS_ID; hit.u = hit.v = 0.0f; hit.prim = MRT_MISS_ID;
    if (bvh.num_nodes == 0) return hit;

    const float tiny = 1e-20f;
    float3 dd = f3(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x), fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y),
                   fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
    const float3 idir = f3(1.0f / dd.x, 1.0f / dd.y, 1.0f / dd.z);
    const bool negx = idir.x < 0.0f, negy = idir.y < 0.0f, negz = idir.z < 0.0f;
    const unsigned oct_inv = 7u - ((negx ? 1u : 0u) | (negy ? 2u : 0u) | (negz ? 4u : 0u));
    const RayShear rs = make_shear(d);

    TraceStack S;
    S.sm = sm_column;
    S.sp = 0;
    uint2 ng = make_uint2(0u, 0x80000000u);  // root "group": node 0, one pending inner hit
    uint2 tg = make_uint2(0u, 0u);

    for (;;) {
        if (ng.y & 0xFF000000u) {
            const unsigned bit = 31u - __clz(ng.y);
            ng.y &= ~(1u << bit);
            if (ng.y & 0xFF000000u) stack_push(S, ng, cnt.overflow);
            const unsigned slot = (bit - 24u) ^ oct_inv;
            const unsigned rel = __popc(ng.y & ~(0xFFFFFFFFu << slot));
            const uint4* np = reinterpret_cast<const uint4*>(bvh.nodes + (ng.x + rel));
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            cnt.nodes++;

            const float adx = __uint_as_float((n0.w & 0xFFu) << 23) * idir.x;
            const float ady = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idir.y;
            const float adz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idir.z;
            const float ox = (__uint_as_float(n0.x) - o.x) * idir.x;
            const float oy = (__uint_as_float(n0.y) - o.y) * idir.y;
            const float oz = (__uint_as_float(n0.z) - o.z) * idir.z;
            // near / far quantised planes per axis, chosen by the ray octant
            const uint2 qlx = make_uint2(n2.x, n2.y), qly = make_uint2(n2.z, n2.w), qlz = make_uint2(n3.x, n3.y);
            const uint2 qhx = make_uint2(n3.z, n3.w), qhy = make_uint2(n4.x, n4.y), qhz = make_uint2(n4.z, n4.w);
            const uint2 nx = negx ? qhx : qlx, fx = negx ? qlx : qhx;
            const uint2 ny = negy ? qhy : qly, fy = negy ? qly : qhy;
            const uint2 nz = negz ? qhz : qlz, fz = negz ? qlz : qhz;
            const float tlimit = hit.t;

            unsigned hitmask = 0u;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const unsigned meta4 = half ? n1.w : n1.z;
                const unsigned wnx = half ? nx.y : nx.x, wny = half ? ny.y : ny.x, wnz = half ? nz.y : nz.x;
                const unsigned wfx = half ? fx.y : fx.x, wfy = half ? fy.y : fy.x, wfz = half ? fz.y : fz.x;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float t0x = fmaf(q_to_float(wnx, j), adx, ox), t1x = fmaf(q_to_float(wfx, j), adx, ox);
                    const float t0y = fmaf(q_to_float(wny, j), ady, oy), t1y = fmaf(q_to_float(wfy, j), ady, oy);
                    const float t0z = fmaf(q_to_float(wnz, j), adz, oz), t1z = fmaf(q_to_float(wfz, j), adz, oz);
                    // padded so fp32 rounding can never cull a box the exact ray touches
                    const float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f)) * 0.9999995f;
                    const float tmax = fminf(fminf(t1x, t1y), fminf(t1z, tlimit)) * 1.0000005f;
                    if (tmin <= tmax) {
                        const unsigned m = (meta4 >> (8 * j)) & 0xFFu;
                        const unsigned inner = ((m & 0x18u) == 0x18u) ? 7u : 0u;
                        const unsigned bit_index = (m ^ (oct_inv & inner)) & 31u;
                        hitmask |= (m >> 5) << bit_index;
                    }
                }
            }
            ng = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
            tg = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
        } else {
            tg = ng;
            ng = make_uint2(0u, 0u);
        }

        while (tg.y) {
            const unsigned bit = __ffs(tg.y) - 1;
            tg.y &= tg.y - 1;
            const uint32_t tri = tg.x + bit;
            const float4* tp = bvh.tris + 3 * (size_t)tri;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            cnt.tris++;
            float t, u, v;
            if (tri_test(o, rs, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), t, u, v))
                hit_consider(hit, t, u, v, tri, __float_as_uint(v0.w));
        }

        if (!(ng.y & 0xFF000000u)) {
            if (S.sp == 0) break;
            ng = stack_pop(S);
        }
    }
    return hit;
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
