// bvh_build.cu -- GPU BVH construction (north_star row n2; no reference counterpart):
//   1. per-primitive AABBs + centroid bounds          (k_prim_bounds)
//   2. 60-bit Morton keys (20 bits/axis)               (k_morton)
//   3. LSD radix sort of (key, primitive)              (sort.cu)
//   4. binary LBVH hierarchy, Karras 2012              (k_karras)
//   5. bottom-up AABBs with per-node arrival counters  (k_bin_boxes)
//   6. level-by-level collapse into 8-wide nodes: greedy surface-area expansion of each binary
//      subtree into <= 8 slots, octant-aware slot assignment, child ranges allocated by an
//      exclusive scan so the layout is deterministic        (k_collapse_expand / k_collapse_emit)
//   7. quantise child boxes to 8 bits against the node origin + per-axis power-of-two scale,
//      write 80-byte nodes and the triangles in node-leaf order          (k_emit_nodes)
// Refit (animated scenes) reruns 1, 5 and 7 on the kept topology.
// Every pass streams its arrays once: the build is HBM-bound (DESIGN.md lists bytes per triangle).
#include <utility>

#include "context.cuh"

namespace {

MRT_D uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
MRT_D float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void k_init_bounds(uint32_t* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;   // min
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;       // max
}

__global__ void __launch_bounds__(256) k_prim_bounds(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t n,
                                                     float4* __restrict__ prim_lo, float4* __restrict__ prim_hi,
                                                     uint32_t* __restrict__ bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 clo = f3s(3.0e38f), chi = f3s(-3.0e38f);
    if (i < n) {
        uint32_t i0 = idx[3 * (size_t)i], i1 = idx[3 * (size_t)i + 1], i2 = idx[3 * (size_t)i + 2];
        float3 a = f3(pos[3 * (size_t)i0], pos[3 * (size_t)i0 + 1], pos[3 * (size_t)i0 + 2]);
        float3 b = f3(pos[3 * (size_t)i1], pos[3 * (size_t)i1 + 1], pos[3 * (size_t)i1 + 2]);
        float3 c = f3(pos[3 * (size_t)i2], pos[3 * (size_t)i2 + 1], pos[3 * (size_t)i2 + 2]);
        float3 lo = f3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
        float3 hi = f3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
        prim_lo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        prim_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
        clo = chi = (lo + hi) * 0.5f;
    }
    // warp reduce, then one set of atomics per CTA
    __shared__ float sm[6][8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        clo.x = fminf(clo.x, __shfl_xor_sync(0xFFFFFFFFu, clo.x, off));
        clo.y = fminf(clo.y, __shfl_xor_sync(0xFFFFFFFFu, clo.y, off));
        clo.z = fminf(clo.z, __shfl_xor_sync(0xFFFFFFFFu, clo.z, off));
        chi.x = fmaxf(chi.x, __shfl_xor_sync(0xFFFFFFFFu, chi.x, off));
        chi.y = fmaxf(chi.y, __shfl_xor_sync(0xFFFFFFFFu, chi.y, off));
        chi.z = fmaxf(chi.z, __shfl_xor_sync(0xFFFFFFFFu, chi.z, off));
    }
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sm[0][warp] = clo.x; sm[1][warp] = clo.y; sm[2][warp] = clo.z;
        sm[3][warp] = chi.x; sm[4][warp] = chi.y; sm[5][warp] = chi.z;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = sm[threadIdx.x][0];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? fminf(v, sm[threadIdx.x][w]) : fmaxf(v, sm[threadIdx.x][w]);
        if (threadIdx.x < 3) atomicMin(&bounds[threadIdx.x], float_to_ordered(v));
        else atomicMax(&bounds[threadIdx.x], float_to_ordered(v));
    }
}

MRT_D uint64_t expand_bits_21(uint64_t x) {
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi, uint32_t n,
                                                const uint32_t* __restrict__ bounds, uint64_t* __restrict__ keys,
                                                uint32_t* __restrict__ order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 smin = f3(ordered_to_float(bounds[0]), ordered_to_float(bounds[1]), ordered_to_float(bounds[2]));
    float3 smax = f3(ordered_to_float(bounds[3]), ordered_to_float(bounds[4]), ordered_to_float(bounds[5]));
    float4 lo = prim_lo[i], hi = prim_hi[i];
    float3 c = (f3(lo.x, lo.y, lo.z) + f3(hi.x, hi.y, hi.z)) * 0.5f;
    float3 ext = smax - smin;
    float3 q = f3(ext.x > 0.0f ? (c.x - smin.x) / ext.x : 0.0f, ext.y > 0.0f ? (c.y - smin.y) / ext.y : 0.0f,
                  ext.z > 0.0f ? (c.z - smin.z) / ext.z : 0.0f);
    const float S = 1048576.0f;  // 2^20 cells per axis
    uint64_t qx = (uint64_t)fminf(fmaxf(q.x * S, 0.0f), S - 1.0f);
    uint64_t qy = (uint64_t)fminf(fmaxf(q.y * S, 0.0f), S - 1.0f);
    uint64_t qz = (uint64_t)fminf(fmaxf(q.z * S, 0.0f), S - 1.0f);
    keys[i] = expand_bits_21(qx) | (expand_bits_21(qy) << 1) | (expand_bits_21(qz) << 2);
    order[i] = i;
}

// ---- Karras 2012: one thread per internal node ----
MRT_D int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)(i ^ j));
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) k_karras(const uint64_t* __restrict__ keys, int n, int32_t* __restrict__ left,
                                                int32_t* __restrict__ right, int32_t* __restrict__ parent,
                                                uint32_t* __restrict__ first, uint32_t* __restrict__ last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int L = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int R = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = L;
    right[i] = R;
    parent[L] = i;
    parent[R] = i;
    first[i] = (uint32_t)lo;
    last[i] = (uint32_t)hi;
    if (i == 0) parent[0] = -1;
}

// leaves write their box, then climb; the second thread to arrive at a node merges its children
__global__ void __launch_bounds__(256) k_bin_boxes(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                                   const uint32_t* __restrict__ order, int n, const int32_t* __restrict__ left,
                                                   const int32_t* __restrict__ right, const int32_t* __restrict__ parent,
                                                   float4* bin_lo, float4* bin_hi, uint32_t* flag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = n - 1 + k;
    uint32_t prim = order[k];
    float4 lo = prim_lo[prim], hi = prim_hi[prim];
    bin_lo[node] = lo;
    bin_hi[node] = hi;
    if (n == 1) return;
    __threadfence();
    int cur = parent[node];
    while (cur >= 0) {
        if (atomicAdd(&flag[cur], 1u) == 0u) return;
        __threadfence();
        int a = left[cur], b = right[cur];
        float4 alo = __ldcg(&bin_lo[a]), ahi = __ldcg(&bin_hi[a]);
        float4 blo = __ldcg(&bin_lo[b]), bhi = __ldcg(&bin_hi[b]);
        lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
        hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
        __stcg(&bin_lo[cur], lo);
        __stcg(&bin_hi[cur], hi);
        __threadfence();
        cur = parent[cur];
    }
}

// ---- collapse to 8-wide ----
struct BinTree {
    const int32_t* left;
    const int32_t* right;
    const uint32_t* first;
    const uint32_t* last;
    const float4* lo;
    const float4* hi;
    int n;  // primitives
};
MRT_D bool bin_is_leaf(const BinTree& T, int b) { return b >= T.n - 1; }
MRT_D uint32_t bin_count(const BinTree& T, int b) { return bin_is_leaf(T, b) ? 1u : (T.last[b] - T.first[b] + 1u); }
MRT_D uint32_t bin_first(const BinTree& T, int b) { return bin_is_leaf(T, b) ? (uint32_t)(b - (T.n - 1)) : T.first[b]; }
MRT_D float bin_area(const BinTree& T, int b) {
    float4 lo = T.lo[b], hi = T.hi[b];
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// One thread per wide node of the current level: choose its <= 8 slots.
__global__ void __launch_bounds__(128) k_collapse_expand(BinTree T, const uint2* __restrict__ items, uint32_t count,
                                                         int32_t* __restrict__ slot_node, uint32_t* __restrict__ node_nchild,
                                                         uint32_t* __restrict__ node_ntri) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int b = (int)items[k].x;
    uint32_t w = items[k].y;
    int slots[8];
    int ns;
    if (bin_is_leaf(T, b)) {
        slots[0] = b;
        ns = 1;
    } else {
        slots[0] = T.left[b];
        slots[1] = T.right[b];
        ns = 2;
    }
    // phase 1: open the largest subtree that is too big to be a leaf; phase 2: use spare slots
    // to split small multi-triangle leaves (tighter boxes, fewer triangle tests)
    for (int phase = 0; phase < 2; phase++) {
        while (ns < 8) {
            int best = -1;
            float bestA = -1.0f;
            for (int s = 0; s < ns; s++) {
                int c = slots[s];
                if (bin_is_leaf(T, c)) continue;
                bool big = bin_count(T, c) > MRT_MAX_LEAF_TRIS;
                if (phase == 0 ? !big : big) continue;
                float A = bin_area(T, c);
                if (A > bestA) { bestA = A; best = s; }
            }
            if (best < 0) break;
            int c = slots[best];
            slots[best] = T.left[c];
            slots[ns++] = T.right[c];
        }
    }
    // octant-aware slot assignment: slot s lies towards (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z)
    // of the node centre, so that traversal priority (slot ^ ray octant) approximates front-to-back
    float4 nlo = T.lo[b], nhi = T.hi[b];
    float3 nc = f3(nlo.x + nhi.x, nlo.y + nhi.y, nlo.z + nhi.z);
    float3 off[8];
    for (int s = 0; s < ns; s++) {
        float4 lo = T.lo[slots[s]], hi = T.hi[slots[s]];
        off[s] = f3(lo.x + hi.x, lo.y + hi.y, lo.z + hi.z) - nc;
    }
    int assigned[8];
    for (int s = 0; s < 8; s++) assigned[s] = -1;
    unsigned child_done = 0, slot_done = 0;
    for (int it = 0; it < ns; it++) {
        float bestC = -3.0e38f;
        int bc = -1, bs = -1;
        for (int c = 0; c < ns; c++) {
            if (child_done & (1u << c)) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_done & (1u << s)) continue;
                float cost = ((s & 1) ? off[c].x : -off[c].x) + ((s & 2) ? off[c].y : -off[c].y) +
                             ((s & 4) ? off[c].z : -off[c].z);
                if (cost > bestC) { bestC = cost; bc = c; bs = s; }
            }
        }
        assigned[bs] = slots[bc];
        child_done |= 1u << bc;
        slot_done |= 1u << bs;
    }
    uint32_t nchild = 0, ntri = 0;
    for (int s = 0; s < 8; s++) {
        int c = assigned[s];
        slot_node[(size_t)w * 8 + s] = c;
        if (c < 0) continue;
        uint32_t cnt = bin_count(T, c);
        if (cnt > MRT_MAX_LEAF_TRIS) nchild++;
        else ntri += cnt;
    }
    node_nchild[w] = nchild;
    node_ntri[w] = ntri;
}

// child_off: exclusive scan of node_nchild over this level (indexed like node_nchild)
__global__ void __launch_bounds__(128) k_collapse_emit(BinTree T, uint32_t level_start, uint32_t count, uint32_t next_start,
                                                       const int32_t* __restrict__ slot_node,
                                                       const uint32_t* __restrict__ child_off,
                                                       uint32_t* __restrict__ node_child_base, uint2* __restrict__ next_items) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint32_t w = level_start + k;
    uint32_t base = next_start + child_off[w];
    node_child_base[w] = base;
    uint32_t rel = 0;
    for (int s = 0; s < 8; s++) {
        int c = slot_node[(size_t)w * 8 + s];
        if (c < 0) continue;
        if (bin_count(T, c) > MRT_MAX_LEAF_TRIS) {
            next_items[base - next_start + rel] = make_uint2((uint32_t)c, base + rel);
            rel++;
        }
    }
}

// Quantisation grid of a wide node, per axis: 255 steps of size 2^e starting two steps below the node
// box, sized so that 250 steps span the box.  Every child plane then lies strictly inside the grid with
// room to spare, which lets the traversal fold the integer->float decode into its FMA (trace.cuh) at the
// price of a plane error of at most 1/512 step; planes are rounded outward with 1/128 step of slack.
MRT_D uint32_t grid_exponent(float ext) {
    // biased exponent e such that 2^(e-127) * 250 >= ext
    float s = ext / 250.0f;
    uint32_t b = __float_as_uint(s);
    uint32_t e = (b >> 23) & 0xFFu;
    if (b & 0x7FFFFFu) e += 1;
    while (e < 254u && __uint_as_float(e << 23) * 250.0f < ext) e++;  // guard against rounding in the division
    return e > 254u ? 254u : e;
}

// One thread per wide node: quantise, pack, and copy the node's triangles in leaf order.
// Leaf triangles: slot s owns bits 3s..3s+2 of leafmask24 (one bit per triangle present); the triangle
// behind bit b is tris[tri_base + popc(leafmask24 & ((1 << b) - 1))].  Empty slots get the inverted
// box lo = 255, hi = 0 and can never be hit.
__global__ void __launch_bounds__(128)
k_emit_nodes(BinTree T, uint32_t num_nodes, const int32_t* __restrict__ slot_node, const uint32_t* __restrict__ node_child_base,
             const uint32_t* __restrict__ node_tri_base, const uint32_t* __restrict__ order, const float* __restrict__ pos,
             const uint32_t* __restrict__ idx, WideNode* __restrict__ nodes, float4* __restrict__ tris) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_nodes) return;
    int sl[8];
    float3 nlo = f3s(3.0e38f), nhi = f3s(-3.0e38f);
    for (int s = 0; s < 8; s++) {
        sl[s] = slot_node[(size_t)w * 8 + s];
        if (sl[s] < 0) continue;
        float4 lo = T.lo[sl[s]], hi = T.hi[sl[s]];
        nlo = f3(fminf(nlo.x, lo.x), fminf(nlo.y, lo.y), fminf(nlo.z, lo.z));
        nhi = f3(fmaxf(nhi.x, hi.x), fmaxf(nhi.y, hi.y), fmaxf(nhi.z, hi.z));
    }
    uint32_t ex = grid_exponent(nhi.x - nlo.x), ey = grid_exponent(nhi.y - nlo.y), ez = grid_exponent(nhi.z - nlo.z);
    float sc[3] = {__uint_as_float(ex << 23), __uint_as_float(ey << 23), __uint_as_float(ez << 23)};
    float org[3] = {nlo.x - 2.0f * sc[0], nlo.y - 2.0f * sc[1], nlo.z - 2.0f * sc[2]};
    uint32_t qlo[3][2] = {{0, 0}, {0, 0}, {0, 0}}, qhi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t imask = 0, leafmask = 0;
    uint32_t tri_base = node_tri_base[w];
    uint32_t tri_off = 0;
    for (int s = 0; s < 8; s++) {
        int c = sl[s];
        if (c < 0) {
            for (int a = 0; a < 3; a++) qlo[a][s >> 2] |= 255u << (8 * (s & 3));
            continue;
        }
        float4 lo4 = T.lo[c], hi4 = T.hi[c];
        float clo[3] = {lo4.x, lo4.y, lo4.z}, chi[3] = {hi4.x, hi4.y, hi4.z};
        for (int a = 0; a < 3; a++) {
            float ql = 0.0f, qh = 255.0f;
            if (sc[a] > 0.0f) {
                const float slack = sc[a] * 0.015625f;  // 1/64 step; the traversal's decode error is <= 1/256 step
                ql = fminf(fmaxf(floorf((clo[a] - org[a]) / sc[a] - 0.02f), 0.0f), 255.0f);
                qh = fminf(fmaxf(ceilf((chi[a] - org[a]) / sc[a] + 0.02f), 0.0f), 255.0f);
                // outward rounding under the decode arithmetic (origin + q * scale)
                while (ql > 0.0f && org[a] + ql * sc[a] > clo[a] - slack) ql -= 1.0f;
                while (qh < 255.0f && org[a] + qh * sc[a] < chi[a] + slack) qh += 1.0f;
            }
            qlo[a][s >> 2] |= (uint32_t)ql << (8 * (s & 3));
            qhi[a][s >> 2] |= (uint32_t)qh << (8 * (s & 3));
        }
        uint32_t cnt = bin_count(T, c);
        if (cnt > MRT_MAX_LEAF_TRIS) {
            imask |= 1u << s;
        } else {
            leafmask |= ((1u << cnt) - 1u) << (3 * s);
            uint32_t f0 = bin_first(T, c);
            for (uint32_t t = 0; t < cnt; t++) {
                uint32_t prim = order[f0 + t];
                uint32_t i0 = idx[3 * (size_t)prim], i1 = idx[3 * (size_t)prim + 1], i2 = idx[3 * (size_t)prim + 2];
                size_t o = 3 * (size_t)(tri_base + tri_off + t);
                tris[o + 0] = make_float4(pos[3 * (size_t)i0], pos[3 * (size_t)i0 + 1], pos[3 * (size_t)i0 + 2], __uint_as_float(prim));
                tris[o + 1] = make_float4(pos[3 * (size_t)i1], pos[3 * (size_t)i1 + 1], pos[3 * (size_t)i1 + 2], 0.0f);
                tris[o + 2] = make_float4(pos[3 * (size_t)i2], pos[3 * (size_t)i2 + 1], pos[3 * (size_t)i2 + 2], 0.0f);
            }
            tri_off += cnt;
        }
    }
    WideNode N;
    N.w[0] = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]),
                        ex | (ey << 8) | (ez << 16) | (imask << 24));
    N.w[1] = make_uint4(node_child_base[w], tri_base, leafmask, 0u);
    N.w[2] = make_uint4(qlo[0][0], qlo[0][1], qlo[1][0], qlo[1][1]);
    N.w[3] = make_uint4(qlo[2][0], qlo[2][1], qhi[0][0], qhi[0][1]);
    N.w[4] = make_uint4(qhi[1][0], qhi[1][1], qhi[2][0], qhi[2][1]);
    nodes[w] = N;
}

__global__ void k_last_total(const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt, uint32_t last, uint32_t* out) {
    out[0] = off[last] + cnt[last];
}

BinTree make_tree(mrt_context* ctx) {
    BinTree T;
    T.left = ctx->bin_left.p; T.right = ctx->bin_right.p; T.first = ctx->bin_first.p; T.last = ctx->bin_last.p;
    T.lo = ctx->bin_lo.p; T.hi = ctx->bin_hi.p; T.n = (int)ctx->ntris;
    return T;
}

int compute_boxes(mrt_context* ctx) {
    const uint32_t n = ctx->ntris;
    k_init_bounds<<<1, 32, 0, ctx->stream>>>(ctx->scene_bounds.p);
    MRT_LAUNCHED(ctx);
    k_prim_bounds<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->pos.p, ctx->idx.p, n, ctx->prim_lo.p, ctx->prim_hi.p,
                                                           ctx->scene_bounds.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "prim_bounds");
}

int climb_boxes(mrt_context* ctx) {
    const uint32_t n = ctx->ntris;
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->bin_flag.p, 0, sizeof(uint32_t) * (size_t)(n ? n : 1), ctx->stream));
    k_bin_boxes<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->prim_lo.p, ctx->prim_hi.p, ctx->order.p, (int)n, ctx->bin_left.p,
                                                         ctx->bin_right.p, ctx->bin_parent.p, ctx->bin_lo.p, ctx->bin_hi.p,
                                                         ctx->bin_flag.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "bin_boxes");
}

int emit_nodes(mrt_context* ctx) {
    k_emit_nodes<<<div_up(ctx->num_nodes, 128), 128, 0, ctx->stream>>>(make_tree(ctx), ctx->num_nodes, ctx->slot_node.p,
                                                                       ctx->node_child_base.p, ctx->node_tri_base.p,
                                                                       ctx->order.p, ctx->pos.p, ctx->idx.p, ctx->nodes.p,
                                                                       ctx->tris.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "emit_nodes");
}

}  // namespace

int bvh_build_full(mrt_context* ctx) {
    const uint32_t n = ctx->ntris;
    ctx->bvh_valid = false;
    ctx->num_nodes = 0;
    ctx->num_leaf_tris = 0;
    if (n == 0) {
        ctx->bvh_valid = true;
        return MRT_OK;
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    MRT_TRY(dev_reserve(ctx, ctx->prim_lo, n));
    MRT_TRY(dev_reserve(ctx, ctx->prim_hi, n));
    MRT_TRY(dev_reserve(ctx, ctx->keys, n));
    MRT_TRY(dev_reserve(ctx, ctx->keys_alt, n));
    MRT_TRY(dev_reserve(ctx, ctx->order, n));
    MRT_TRY(dev_reserve(ctx, ctx->order_alt, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_left, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_right, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_parent, 2 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_first, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_last, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_lo, 2 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_hi, 2 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_flag, n));
    MRT_TRY(dev_reserve(ctx, ctx->scene_bounds, 8));
    MRT_TRY(dev_reserve(ctx, ctx->work_a, n));
    MRT_TRY(dev_reserve(ctx, ctx->work_b, n));
    MRT_TRY(dev_reserve(ctx, ctx->slot_node, 8 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->node_nchild, (size_t)n + 1));
    MRT_TRY(dev_reserve(ctx, ctx->node_ntri, (size_t)n + 1));
    MRT_TRY(dev_reserve(ctx, ctx->node_child_base, (size_t)n + 1));
    MRT_TRY(dev_reserve(ctx, ctx->node_tri_base, (size_t)n + 1));
    MRT_TRY(dev_reserve(ctx, ctx->counters, 16));

    // 1-3: boxes, Morton keys, sort
    MRT_TRY(compute_boxes(ctx));
    k_morton<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->prim_lo.p, ctx->prim_hi.p, n, ctx->scene_bounds.p, ctx->keys.p,
                                                      ctx->order.p);
    MRT_LAUNCHED(ctx);
    bool in_alt = false;
    MRT_TRY(radix_sort_pairs_u64(ctx, ctx->keys.p, ctx->keys_alt.p, ctx->order.p, ctx->order_alt.p, n, 0, 64, &in_alt));
    if (in_alt) {
        std::swap(ctx->keys, ctx->keys_alt);
        std::swap(ctx->order, ctx->order_alt);
    }
    // 4-5: hierarchy + boxes
    if (n > 1) {
        k_karras<<<div_up(n - 1, 256), 256, 0, ctx->stream>>>(ctx->keys.p, (int)n, ctx->bin_left.p, ctx->bin_right.p,
                                                              ctx->bin_parent.p, ctx->bin_first.p, ctx->bin_last.p);
        MRT_LAUNCHED(ctx);
    }
    MRT_TRY(climb_boxes(ctx));

    // 6: collapse, one level per iteration
    BinTree T = make_tree(ctx);
    uint2 root = make_uint2(n > 1 ? 0u : 0u /* single leaf = binary node n-1 = 0 */, 0u);
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->work_a.p, &root, sizeof root, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t level_start = 0, level_count = 1;
    DevArray<uint2>* cur = &ctx->work_a;
    DevArray<uint2>* nxt = &ctx->work_b;
    while (level_count > 0) {
        if ((size_t)level_start + level_count > n) return mrt_fail(ctx, MRT_ERR_INVALID, "wide BVH node budget exceeded");
        k_collapse_expand<<<div_up(level_count, 128), 128, 0, ctx->stream>>>(T, cur->p, level_count, ctx->slot_node.p,
                                                                             ctx->node_nchild.p, ctx->node_ntri.p);
        MRT_LAUNCHED(ctx);
        // child offsets of this level; node_child_base temporarily holds the scan
        MRT_TRY(scan_exclusive_u32(ctx, ctx->node_nchild.p + level_start, ctx->node_tri_base.p + level_start, level_count));
        k_last_total<<<1, 1, 0, ctx->stream>>>(ctx->node_tri_base.p + level_start, ctx->node_nchild.p + level_start,
                                               level_count - 1, ctx->counters.p);
        MRT_LAUNCHED(ctx);
        uint32_t next_count = 0;
        MRT_CUDA(ctx, cudaMemcpyAsync(&next_count, ctx->counters.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        uint32_t next_start = level_start + level_count;
        if ((size_t)next_start + next_count > n) return mrt_fail(ctx, MRT_ERR_INVALID, "wide BVH node budget exceeded");
        // k_collapse_emit indexes child_off by wide node id
        k_collapse_emit<<<div_up(level_count, 128), 128, 0, ctx->stream>>>(T, level_start, level_count, next_start,
                                                                           ctx->slot_node.p, ctx->node_tri_base.p,
                                                                           ctx->node_child_base.p, nxt->p);
        MRT_LAUNCHED(ctx);
        level_start = next_start;
        level_count = next_count;
        std::swap(cur, nxt);
    }
    ctx->num_nodes = level_start;
    ctx->num_leaf_tris = n;

    // 7: triangle ranges + final nodes
    MRT_TRY(scan_exclusive_u32(ctx, ctx->node_ntri.p, ctx->node_tri_base.p, ctx->num_nodes));
    MRT_TRY(dev_reserve(ctx, ctx->nodes, ctx->num_nodes));
    MRT_TRY(dev_reserve(ctx, ctx->tris, 3 * (size_t)n));
    MRT_TRY(emit_nodes(ctx));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev[0], ctx->ev[1]);
    ctx->stats.num_triangles = n;
    ctx->stats.num_wide_nodes = ctx->num_nodes;
    ctx->stats.bvh_bytes = (uint64_t)ctx->num_nodes * sizeof(WideNode) + (uint64_t)n * 48u;
    ctx->bvh_valid = true;
    return MRT_OK;
}

int bvh_refit(mrt_context* ctx) {
    if (!ctx->bvh_valid || ctx->num_nodes == 0) return bvh_build_full(ctx);
    cudaEventRecord(ctx->ev[0], ctx->stream);
    MRT_TRY(compute_boxes(ctx));
    MRT_TRY(climb_boxes(ctx));
    MRT_TRY(emit_nodes(ctx));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev[0], ctx->ev[1]);
    return MRT_OK;
}
