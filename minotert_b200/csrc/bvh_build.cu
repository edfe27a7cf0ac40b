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
#include <cooperative_groups.h>

#include <utility>

#include "context.cuh"

namespace cg = cooperative_groups;

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
                                                uint32_t* __restrict__ count) {
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
    count[i] = (uint32_t)(hi - lo + 1);
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

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) over the Morton order ----
// Each round, every cluster finds the neighbour within +-radius positions whose merged box has the
// smallest surface area; mutual nearest neighbours merge into a new internal node.  Node ids and the
// compacted cluster list come from exclusive scans, so the tree is deterministic.  Compared with the
// Karras hierarchy (splits dictated by Morton bits) this approaches SAH quality.
MRT_D float merged_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    float dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    float dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(256) k_ploc_init(uint32_t n, uint32_t* __restrict__ clusters, int32_t* __restrict__ parent) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    clusters[i] = n - 1 + i;  // leaf of sorted position i
    parent[n - 1 + i] = -1;
}

// nearest cluster of position i within +-radius positions (plain loads: the device-side loop mutates these arrays)
MRT_D uint32_t ploc_nearest_one(const uint32_t* clusters, uint32_t m, int radius, const float4* lo, const float4* hi, uint32_t i) {
    uint32_t ci = clusters[i];
    float4 ilo = lo[ci], ihi = hi[ci];
    int j0 = max((int)i - radius, 0), j1 = min((int)i + radius, (int)m - 1);
    // Ties are the rule on regular meshes (rows of identical quads).  Breaking them by "lowest position"
    // makes every cluster point down the row and only one pair per row is mutual per round (slow rounds,
    // caterpillar trees).  The tie-break below is symmetric in (i, j): nearer position first, then pairs
    // whose lower position is even -- in a uniform row that pairs (0,1), (2,3), ... in a single round.
    float best = 3.0e38f;
    uint32_t bj = i, bkey = 0xFFFFFFFFu;
    for (int j = j0; j <= j1; j++) {
        if (j == (int)i) continue;
        uint32_t cj = clusters[j];
        float a = merged_area(ilo, ihi, lo[cj], hi[cj]);
        uint32_t dist = (uint32_t)abs(j - (int)i);
        uint32_t lowpos = (uint32_t)min(j, (int)i);
        // (distance, parity of the lower position, lower position): a total order on the pairs around i that
        // both partners evaluate identically, so the globally best pair is always mutual (progress)
        uint32_t key = (dist << 26) | ((lowpos & 1u) << 25) | (lowpos & 0x1FFFFFFu);
        if (a < best || (a == best && key < bkey)) { best = a; bj = (uint32_t)j; bkey = key; }
    }
    return bj;
}

__global__ void __launch_bounds__(256) k_ploc_nearest(const uint32_t* __restrict__ clusters, uint32_t m, int radius,
                                                      const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                      uint32_t* __restrict__ nn) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    nn[i] = ploc_nearest_one(clusters, m, radius, lo, hi, i);
}

// flags: create[i] = 1 if i is the left partner of a mutual pair; keep[i] = 0 if i is the right partner
__global__ void __launch_bounds__(256) k_ploc_flags(const uint32_t* __restrict__ nn, uint32_t m, uint32_t* __restrict__ create,
                                                    uint32_t* __restrict__ keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t j = nn[i];
    bool mutual = j != i && nn[j] == i;
    create[i] = (mutual && i < j) ? 1u : 0u;
    keep[i] = (mutual && i > j) ? 0u : 1u;
}

__global__ void __launch_bounds__(256)
k_ploc_merge(const uint32_t* __restrict__ clusters, const uint32_t* __restrict__ nn, uint32_t m, const uint32_t* __restrict__ create,
             const uint32_t* __restrict__ keep, const uint32_t* __restrict__ create_scan, const uint32_t* __restrict__ keep_scan,
             uint32_t next_node, int nprims, int32_t* __restrict__ left, int32_t* __restrict__ right, int32_t* __restrict__ parent,
             uint32_t* __restrict__ count, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* __restrict__ clusters_out,
             uint32_t* __restrict__ totals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (i == m - 1) {
        totals[0] = keep_scan[i] + keep[i];      // clusters after this round
        totals[1] = create_scan[i] + create[i];  // nodes created
    }
    if (!keep[i]) return;
    uint32_t c = clusters[i];
    if (create[i]) {
        uint32_t j = nn[i], cj = clusters[j];
        uint32_t id = next_node + create_scan[i];
        float4 alo = lo[c], ahi = hi[c], blo = lo[cj], bhi = hi[cj];
        left[id] = (int32_t)c;
        right[id] = (int32_t)cj;
        parent[c] = (int32_t)id;
        parent[cj] = (int32_t)id;
        parent[id] = -1;
        uint32_t na = (int)c >= nprims - 1 ? 1u : count[c], nb = (int)cj >= nprims - 1 ? 1u : count[cj];
        count[id] = na + nb;
        lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
        hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
        c = id;
    }
    clusters_out[keep_scan[i]] = c;
}

__global__ void __launch_bounds__(256) k_leaf_boxes(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                                    const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ bin_lo,
                                                    float4* __restrict__ bin_hi) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t prim = order[k];
    bin_lo[n - 1 + k] = prim_lo[prim];
    bin_hi[n - 1 + k] = prim_hi[prim];
}

// ---- PLOC with the round loop on the device ----
// The host-driven loop above costs ~10 launches and one readback per round (40 rounds at 1 M triangles: 443
// launches, 3.5 ms, of which the arithmetic is a few per cent).  k_ploc_loop runs every round inside ONE
// cooperative launch: phases are separated by grid-wide barriers, the two stream compactions of a round
// (surviving clusters, created nodes) are chunked scans -- each CTA owns a contiguous range of positions, publishes
// its counts, and after the barrier sums the counts of the CTAs before it -- and once few clusters are left a
// single CTA finishes the remaining rounds with CTA barriers only.  Node ids, child order and boxes are produced
// by the same rules as the host-driven loop, so both build the same tree bit for bit (tested).
#define PLOC_MAX_RADIUS 32      // mrt_set_option("ploc_radius") range
#ifndef PLOC_TAIL
#define PLOC_TAIL LOOP_THREADS  // clusters at which CTA 0 takes over (one position per thread)
#endif

// Search radius of a round with m clusters left.  PLOC_TOP_RADIUS > 0: the rounds that build the TOP of the tree (few
// clusters left, each a large region) look further than the configured radius -- nearest-neighbour quality matters most
// where boxes are biggest, and a round over a few thousand clusters costs the same whatever the radius.
#ifndef PLOC_TOP_RADIUS
#define PLOC_TOP_RADIUS 0
#endif
#ifndef PLOC_TOP_CLUSTERS
#define PLOC_TOP_CLUSTERS 16384
#endif
MRT_D int ploc_radius_for(int radius, uint32_t m) {
#if PLOC_TOP_RADIUS > 0
    if (m <= (uint32_t)PLOC_TOP_CLUSTERS) return max(radius, PLOC_TOP_RADIUS);
#endif
    return radius;
}

struct PlocLoop {
    uint32_t n;
    int radius;
    uint32_t* clusters[2];
    uint32_t* nn;
    int32_t *left, *right, *parent;
    uint32_t* count;
    float4 *lo, *hi;
    uint2* block_sums;  // [gridDim.x]: (kept clusters, created nodes) of each CTA's range
    uint32_t* result;   // [0] internal nodes created, [1] status (0 ok, 1 no progress), [2] rounds
};

// CTA-wide helpers of the cooperative loops (LOOP_THREADS threads, one CTA per SM: the grid barrier costs one
// atomic per CTA, so few fat CTAs synchronise faster than many thin ones -- 1184 CTAs of 256: ~6 us per barrier).
#ifndef LOOP_THREADS
#define LOOP_THREADS 1024
#endif
constexpr unsigned LOOP_WARPS = LOOP_THREADS / 32;

// -DBUILD_PROFILE: thread 0 of CTA 0 prints the time between phase boundaries of the cooperative loops
#ifdef BUILD_PROFILE
// phase boundaries are time-stamped into a global table and printed when the kernel ends (a printf per phase costs ~30 us)
MRT_D unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct ProfRec { const char* label; unsigned a, b; unsigned long long dt; };
__device__ ProfRec g_prof[512];
__device__ unsigned g_prof_n;
#define PROF_DECL unsigned long long prof_t = gtimer(); if (blockIdx.x == 0 && threadIdx.x == 0) g_prof_n = 0;
#define PROF(label_, a_, b_) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long n_ = gtimer(); if (g_prof_n < 512) { g_prof[g_prof_n].label = label_; g_prof[g_prof_n].a = (unsigned)(a_); g_prof[g_prof_n].b = (unsigned)(b_); g_prof[g_prof_n].dt = n_ - prof_t; g_prof_n++; } prof_t = gtimer(); } } while (0)
#define PROF_DUMP() do { if (blockIdx.x == 0 && threadIdx.x == 0) for (unsigned k_ = 0; k_ < g_prof_n; k_++) printf("%s %u %u: %llu ns\n", g_prof[k_].label, g_prof[k_].a, g_prof[k_].b, g_prof[k_].dt); } while (0)
#else
#define PROF_DECL
#define PROF(label, a, b) do {} while (0)
#define PROF_DUMP() do {} while (0)
#endif

// exclusive scan of one 32-bit value per thread (packed counters: two 16-bit fields); *total = CTA sum
MRT_D uint32_t cta_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t ws[LOOP_WARPS];
    __shared__ uint32_t tot;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
        if (lane >= (unsigned)off) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < LOOP_WARPS ? ws[lane] : 0u, winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, off);
            if (lane >= (unsigned)off) winc += t;
        }
        if (lane < LOOP_WARPS) ws[lane] = winc - w;
        if (lane == 31) tot = winc;
    }
    __syncthreads();
    const uint32_t r = ws[warp] + inc - v;
    *total = tot;
    __syncthreads();
    return r;
}
// sum of (a, b) over the CTA, returned to every thread
MRT_D uint2 cta_sum2(uint32_t a, uint32_t b) {
    __shared__ uint2 ws[LOOP_WARPS];
    __shared__ uint2 tot;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_down_sync(0xFFFFFFFFu, a, off);
        b += __shfl_down_sync(0xFFFFFFFFu, b, off);
    }
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = make_uint2(a, b);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint2 t = threadIdx.x < LOOP_WARPS ? ws[threadIdx.x] : make_uint2(0u, 0u);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            t.x += __shfl_down_sync(0xFFFFFFFFu, t.x, off);
            t.y += __shfl_down_sync(0xFFFFFFFFu, t.y, off);
        }
        if (threadIdx.x == 0) tot = t;
    }
    __syncthreads();
    const uint2 r = tot;
    __syncthreads();
    return r;
}

// One PLOC round over positions [0, m).  nb CTAs take part (b = this CTA's index); SYNC is the barrier between
// phases: the grid barrier, or __syncthreads when a single CTA runs the tail.  Returns (clusters kept, nodes created).
template <class Sync>
MRT_D uint2 ploc_round(const PlocLoop& A, uint32_t m, int cur, uint32_t next_node, uint32_t b, uint32_t nb, float4* wlo, float4* whi,
                        Sync sync) {
    const uint32_t* C = A.clusters[cur];
    uint32_t* Cout = A.clusters[cur ^ 1];
    // phase 1: nearest neighbour of every position.  A CTA takes LOOP_THREADS consecutive positions at a time and stages
    // the boxes of that window (+- radius) in shared memory: one gather per position instead of one per pair.
    {
        const int radius = ploc_radius_for(A.radius, m);
        for (uint32_t t0 = b * LOOP_THREADS; t0 < m; t0 += nb * LOOP_THREADS) {
            const int w0 = (int)t0 - radius;                       // window = positions [w0, w0 + LOOP_THREADS + 2 radius)
            for (int q = threadIdx.x; q < LOOP_THREADS + 2 * radius; q += LOOP_THREADS) {
                const int pos = w0 + q;
                if (pos >= 0 && pos < (int)m) {
                    const uint32_t c = C[pos];
                    wlo[q] = A.lo[c];
                    whi[q] = A.hi[c];
                }
            }
            __syncthreads();
            const uint32_t i = t0 + threadIdx.x;
            if (i < m) {
                const int qi = (int)threadIdx.x + radius;
                const float4 ilo = wlo[qi], ihi = whi[qi];
                const int j0 = max((int)i - radius, 0), j1 = min((int)i + radius, (int)m - 1);
                float best = 3.0e38f;
                uint32_t bj = i, bkey = 0xFFFFFFFFu;
                for (int j = j0; j <= j1; j++) {  // same pair order and tie-break as ploc_nearest_one
                    if (j == (int)i) continue;
                    const float a = merged_area(ilo, ihi, wlo[j - w0], whi[j - w0]);
                    const uint32_t dist = (uint32_t)abs(j - (int)i);
                    const uint32_t lowpos = (uint32_t)min(j, (int)i);
                    const uint32_t key = (dist << 26) | ((lowpos & 1u) << 25) | (lowpos & 0x1FFFFFFu);
                    if (a < best || (a == best && key < bkey)) { best = a; bj = (uint32_t)j; bkey = key; }
                }
                A.nn[i] = bj;
            }
            __syncthreads();
        }
    }
    sync();
    // phase 2: counts of this CTA's range
    const uint32_t chunk = (m + nb - 1) / nb;
    const uint32_t r0 = min(m, b * chunk), r1 = min(m, r0 + chunk);
    uint32_t keepc = 0, createc = 0;
    for (uint32_t i = r0 + threadIdx.x; i < r1; i += LOOP_THREADS) {
        const uint32_t j = A.nn[i];
        const bool mutual = j != i && A.nn[j] == i;
        keepc += (mutual && i > j) ? 0u : 1u;
        createc += (mutual && i < j) ? 1u : 0u;
    }
    const uint2 mine = cta_sum2(keepc, createc);
    if (threadIdx.x == 0) A.block_sums[b] = mine;
    sync();
    // phase 3: offsets of the range, then compaction + merges tile by tile
    uint32_t kb = 0, cb = 0, kt = 0, ct = 0;
    for (uint32_t q = threadIdx.x; q < nb; q += LOOP_THREADS) {
        const uint2 sq = A.block_sums[q];
        kt += sq.x; ct += sq.y;
        if (q < b) { kb += sq.x; cb += sq.y; }
    }
    const uint2 before = cta_sum2(kb, cb), total = cta_sum2(kt, ct);
    uint32_t keep_base = before.x, create_base = before.y;
    const int nprims = (int)A.n;
    for (uint32_t base = r0; base < r1; base += LOOP_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < r1;
        uint32_t j = 0;
        bool keep = false, create = false;
        if (valid) {
            j = A.nn[i];
            const bool mutual = j != i && A.nn[j] == i;
            keep = !(mutual && i > j);
            create = mutual && i < j;
        }
        uint32_t tile_total;
        const uint32_t ex = cta_scan((keep ? 1u : 0u) | (create ? 0x10000u : 0u), &tile_total);
        if (keep) {
            uint32_t c = C[i];
            if (create) {
                const uint32_t cj = C[j];
                const uint32_t id = next_node + create_base + (ex >> 16);
                const float4 alo = A.lo[c], ahi = A.hi[c], blo = A.lo[cj], bhi = A.hi[cj];
                A.left[id] = (int32_t)c;
                A.right[id] = (int32_t)cj;
                A.parent[c] = (int32_t)id;
                A.parent[cj] = (int32_t)id;
                A.parent[id] = -1;
                const uint32_t na = (int)c >= nprims - 1 ? 1u : A.count[c], nbb = (int)cj >= nprims - 1 ? 1u : A.count[cj];
                A.count[id] = na + nbb;
                A.lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
                A.hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
                c = id;
            }
            Cout[keep_base + (ex & 0xFFFFu)] = c;
        }
        keep_base += tile_total & 0xFFFFu;
        create_base += tile_total >> 16;
    }
    sync();
    return total;
}

// A round in which every CTA's range of positions, with a halo of 2 radius on either side, fits one window of
// LOOP_THREADS positions (<= ~130 k clusters): the window's ids, boxes and counts are read once into shared memory;
// nearest neighbours are computed for the range plus a halo of one radius, so that the mutual-pair test of the range
// needs no other CTA's result -- one grid barrier (and four dependent trips to L2) less than ploc_round; the merges
// take their boxes and counts from shared memory.  Same pairs, ids and boxes as ploc_round (ids come from a scan over
// positions, whatever the chunking).
#ifndef PLOC_FUSED_ROUNDS
#define PLOC_FUSED_ROUNDS 1
#endif
template <class Sync>
MRT_D uint2 ploc_round_fused(const PlocLoop& A, uint32_t m, int cur, uint32_t next_node, uint32_t b, uint32_t nb, float4* wlo,
                              float4* whi, uint32_t* wid, uint32_t* wcnt, uint32_t* wnn, Sync sync) {
    const uint32_t* C = A.clusters[cur];
    uint32_t* Cout = A.clusters[cur ^ 1];
    const int radius = ploc_radius_for(A.radius, m);
    const uint32_t chunk = (m + nb - 1) / nb;  // the caller checked chunk + 4 radius <= LOOP_THREADS
    const uint32_t r0 = min(m, b * chunk), r1 = min(m, r0 + chunk);
    const int w0 = (int)r0 - 2 * radius;       // window = positions [w0, w0 + LOOP_THREADS)
    const int q = (int)threadIdx.x, pos = w0 + q;
    const int nprims = (int)A.n;
    if (pos >= 0 && pos < (int)min(m, r1 + 2u * (uint32_t)radius)) {
        const uint32_t c = C[pos];
        wid[q] = c;
        wlo[q] = A.lo[c];
        whi[q] = A.hi[c];
        wcnt[q] = (int)c >= nprims - 1 ? 1u : A.count[c];
    }
    __syncthreads();
    if (pos >= max(0, (int)r0 - radius) && pos < (int)min(m, r1 + (uint32_t)radius)) {
        const int i = pos;
        const float4 ilo = wlo[q], ihi = whi[q];
        const int j0 = max(i - radius, 0), j1 = min(i + radius, (int)m - 1);
        float best = 3.0e38f;
        uint32_t bj = (uint32_t)i, bkey = 0xFFFFFFFFu;
        for (int j = j0; j <= j1; j++) {  // same pair order and tie-break as ploc_round
            if (j == i) continue;
            const float a = merged_area(ilo, ihi, wlo[j - w0], whi[j - w0]);
            const uint32_t dist = (uint32_t)abs(j - i);
            const uint32_t lowpos = (uint32_t)min(j, i);
            const uint32_t key = (dist << 26) | ((lowpos & 1u) << 25) | (lowpos & 0x1FFFFFFu);
            if (a < best || (a == best && key < bkey)) { best = a; bj = (uint32_t)j; bkey = key; }
        }
        wnn[q] = bj;
    }
    __syncthreads();
    bool keep = false, create = false;
    uint32_t j = 0;
    if (pos >= (int)r0 && pos < (int)r1) {
        const uint32_t i = (uint32_t)pos;
        j = wnn[q];
        const bool mutual = j != i && wnn[(int)j - w0] == i;
        keep = !(mutual && i > j);
        create = mutual && i < j;
    }
    const uint2 mine = cta_sum2(keep ? 1u : 0u, create ? 1u : 0u);
    if (threadIdx.x == 0) A.block_sums[b] = mine;
    sync();
    uint32_t kb = 0, cb = 0, kt = 0, ct = 0;
    for (uint32_t k = threadIdx.x; k < nb; k += LOOP_THREADS) {
        const uint2 sq = A.block_sums[k];
        kt += sq.x; ct += sq.y;
        if (k < b) { kb += sq.x; cb += sq.y; }
    }
    const uint2 before = cta_sum2(kb, cb), total = cta_sum2(kt, ct);
    uint32_t tile_total;
    const uint32_t ex = cta_scan((keep ? 1u : 0u) | (create ? 0x10000u : 0u), &tile_total);
    if (keep) {
        uint32_t c = wid[q];
        if (create) {
            const int qj = (int)j - w0;
            const uint32_t cj = wid[qj];
            const uint32_t id = next_node + before.y + (ex >> 16);
            const float4 alo = wlo[q], ahi = whi[q], blo = wlo[qj], bhi = whi[qj];
            A.left[id] = (int32_t)c;
            A.right[id] = (int32_t)cj;
            A.parent[c] = (int32_t)id;
            A.parent[cj] = (int32_t)id;
            A.parent[id] = -1;
            A.count[id] = wcnt[q] + wcnt[qj];
            A.lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
            A.hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
            c = id;
        }
        Cout[before.x + (ex & 0xFFFFu)] = c;
    }
    sync();
    return total;
}

// The same idea for the LARGE rounds (a CTA's range spans several windows): per window, nearest neighbours for the window's
// own positions plus one radius of halo, the mutual-pair test from shared memory, and one packed word per own position
// -- keep, create, partner offset -- written where ploc_round keeps its neighbour index.  The counting pass of ploc_round
// (its phase 2: two dependent loads per position, a grid barrier) is gone, and the merge phase reads one word instead of
// chasing nn[i] and nn[nn[i]].  Same pairs, ids and boxes.
template <class Sync>
MRT_D uint2 ploc_round_windows(const PlocLoop& A, uint32_t m, int cur, uint32_t next_node, uint32_t b, uint32_t nb, float4* wlo,
                                float4* whi, uint32_t* wnn, Sync sync) {
    const uint32_t* C = A.clusters[cur];
    uint32_t* Cout = A.clusters[cur ^ 1];
    const int radius = ploc_radius_for(A.radius, m);
    const uint32_t chunk = (m + nb - 1) / nb;
    const uint32_t r0 = min(m, b * chunk), r1 = min(m, r0 + chunk);
    const uint32_t own = (uint32_t)LOOP_THREADS - 4u * (uint32_t)radius;  // own positions per window
    const int q = (int)threadIdx.x;
    uint32_t keepc = 0, createc = 0;
    for (uint32_t a = r0; a < r1; a += own) {
        const uint32_t e = min(a + own, r1);
        const int w0 = (int)a - 2 * radius, pos = w0 + q;
        if (pos >= 0 && pos < (int)min(m, e + 2u * (uint32_t)radius)) {
            const uint32_t c = C[pos];
            wlo[q] = A.lo[c];
            whi[q] = A.hi[c];
        }
        __syncthreads();
        if (pos >= max(0, (int)a - radius) && pos < (int)min(m, e + (uint32_t)radius)) {
            const int i = pos;
            const float4 ilo = wlo[q], ihi = whi[q];
            const int j0 = max(i - radius, 0), j1 = min(i + radius, (int)m - 1);
            float best = 3.0e38f;
            uint32_t bj = (uint32_t)i, bkey = 0xFFFFFFFFu;
            for (int j = j0; j <= j1; j++) {  // same pair order and tie-break as ploc_round
                if (j == i) continue;
                const float ar = merged_area(ilo, ihi, wlo[j - w0], whi[j - w0]);
                const uint32_t dist = (uint32_t)abs(j - i);
                const uint32_t lowpos = (uint32_t)min(j, i);
                const uint32_t key = (dist << 26) | ((lowpos & 1u) << 25) | (lowpos & 0x1FFFFFFu);
                if (ar < best || (ar == best && key < bkey)) { best = ar; bj = (uint32_t)j; bkey = key; }
            }
            wnn[q] = bj;
        }
        __syncthreads();
        if (pos >= (int)a && pos < (int)e) {
            const uint32_t i = (uint32_t)pos, j = wnn[q];
            const bool mutual = j != i && wnn[(int)j - w0] == i;
            const bool keep = !(mutual && i > j), create = mutual && i < j;
            keepc += keep ? 1u : 0u;
            createc += create ? 1u : 0u;
            A.nn[i] = (keep ? 1u : 0u) | (create ? 2u : 0u) | ((uint32_t)((int)j - (int)i + 64) << 2);  // |j - i| <= radius <= 32
        }
        __syncthreads();  // the window is refilled
    }
    const uint2 mine = cta_sum2(keepc, createc);
    if (threadIdx.x == 0) A.block_sums[b] = mine;
    sync();
    uint32_t kb = 0, cb = 0, kt = 0, ct = 0;
    for (uint32_t k = threadIdx.x; k < nb; k += LOOP_THREADS) {
        const uint2 sq = A.block_sums[k];
        kt += sq.x; ct += sq.y;
        if (k < b) { kb += sq.x; cb += sq.y; }
    }
    const uint2 before = cta_sum2(kb, cb), total = cta_sum2(kt, ct);
    uint32_t keep_base = before.x, create_base = before.y;
    const int nprims = (int)A.n;
    for (uint32_t base = r0; base < r1; base += LOOP_THREADS) {
        const uint32_t i = base + threadIdx.x;
        bool keep = false, create = false;
        uint32_t j = 0;
        if (i < r1) {
            const uint32_t f = A.nn[i];  // written by this CTA (this thread's CTA owns [r0, r1)) before the barrier
            keep = (f & 1u) != 0u;
            create = (f & 2u) != 0u;
            j = (uint32_t)((int)i + (int)(f >> 2) - 64);
        }
        uint32_t tile_total;
        const uint32_t ex = cta_scan((keep ? 1u : 0u) | (create ? 0x10000u : 0u), &tile_total);
        if (keep) {
            uint32_t c = C[i];
            if (create) {
                const uint32_t cj = C[j];
                const uint32_t id = next_node + create_base + (ex >> 16);
                const float4 alo = A.lo[c], ahi = A.hi[c], blo = A.lo[cj], bhi = A.hi[cj];
                A.left[id] = (int32_t)c;
                A.right[id] = (int32_t)cj;
                A.parent[c] = (int32_t)id;
                A.parent[cj] = (int32_t)id;
                A.parent[id] = -1;
                const uint32_t na = (int)c >= nprims - 1 ? 1u : A.count[c], nbb = (int)cj >= nprims - 1 ? 1u : A.count[cj];
                A.count[id] = na + nbb;
                A.lo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
                A.hi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
                c = id;
            }
            Cout[keep_base + (ex & 0xFFFFu)] = c;
        }
        keep_base += tile_total & 0xFFFFu;
        create_base += tile_total >> 16;
    }
    sync();
    return total;
}

// The last rounds (<= LOOP_THREADS clusters, CTA 0 alone) with the cluster list in SHARED memory: ids, boxes and primitive
// counts are loaded once, a round is a neighbour search, a mutual-pair test and an in-place compaction between CTA
// barriers, and the only global traffic left is the record of each new node going out (nothing waits for it).  The
// generic round above costs ~5.5 us here (about ten dependent round trips to L2 per round, ~22 rounds at one million
// triangles); this one ~1 us.  Same pairs, same node ids, same boxes: one thread per position, the scan of ploc_round.
#ifndef PLOC_SMEM_TAIL
#define PLOC_SMEM_TAIL 1
#endif
MRT_D void ploc_tail_shared(const PlocLoop& A, uint32_t& m, uint32_t& next_node, uint32_t& rounds, uint32_t& status, int cur,
                            float4* slo, float4* shi, uint32_t* scid, uint32_t* scnt, uint32_t* snn) {
    const uint32_t i = threadIdx.x;
    const int nprims = (int)A.n;
    if (i < m) {
        const uint32_t c = A.clusters[cur][i];
        scid[i] = c;
        slo[i] = A.lo[c];
        shi[i] = A.hi[c];
        scnt[i] = (int)c >= nprims - 1 ? 1u : A.count[c];
    }
    __syncthreads();
    while (status == 0 && m > 1) {
        const int radius = ploc_radius_for(A.radius, m);
        if (i < m) {
            const float4 ilo = slo[i], ihi = shi[i];
            const int j0 = max((int)i - radius, 0), j1 = min((int)i + radius, (int)m - 1);
            float best = 3.0e38f;
            uint32_t bj = i, bkey = 0xFFFFFFFFu;
            for (int j = j0; j <= j1; j++) {  // same pair order and tie-break as ploc_round
                if (j == (int)i) continue;
                const float a = merged_area(ilo, ihi, slo[j], shi[j]);
                const uint32_t dist = (uint32_t)abs(j - (int)i);
                const uint32_t lowpos = (uint32_t)min(j, (int)i);
                const uint32_t key = (dist << 26) | ((lowpos & 1u) << 25) | (lowpos & 0x1FFFFFFu);
                if (a < best || (a == best && key < bkey)) { best = a; bj = (uint32_t)j; bkey = key; }
            }
            snn[i] = bj;
        }
        __syncthreads();
        bool keep = false, create = false;
        uint32_t c = 0, cj = 0, cnt = 0;
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        if (i < m) {
            const uint32_t j = snn[i];
            const bool mutual = j != i && snn[j] == i;
            keep = !(mutual && i > j);
            create = mutual && i < j;
            c = scid[i]; lo = slo[i]; hi = shi[i]; cnt = scnt[i];
            if (create) {
                cj = scid[j];
                const float4 blo = slo[j], bhi = shi[j];
                lo = make_float4(fminf(lo.x, blo.x), fminf(lo.y, blo.y), fminf(lo.z, blo.z), 0.0f);
                hi = make_float4(fmaxf(hi.x, bhi.x), fmaxf(hi.y, bhi.y), fmaxf(hi.z, bhi.z), 0.0f);
                cnt += scnt[j];
            }
        }
        uint32_t tile_total;
        const uint32_t ex = cta_scan((keep ? 1u : 0u) | (create ? 0x10000u : 0u), &tile_total);  // its barriers separate the reads above from the writes below
        if (keep) {
            if (create) {
                const uint32_t id = next_node + (ex >> 16);
                A.left[id] = (int32_t)c;
                A.right[id] = (int32_t)cj;
                A.parent[c] = (int32_t)id;
                A.parent[cj] = (int32_t)id;
                A.parent[id] = -1;
                A.count[id] = cnt;
                A.lo[id] = lo;
                A.hi[id] = hi;
                c = id;
            }
            const uint32_t pos = ex & 0xFFFFu;
            scid[pos] = c; slo[pos] = lo; shi[pos] = hi; scnt[pos] = cnt;
        }
        __syncthreads();
        const uint32_t kept = tile_total & 0xFFFFu, created = tile_total >> 16;
        rounds++;
        if (created == 0 || kept >= m) { status = 1; break; }
        next_node += created;
        m = kept;
    }
}

__global__ void __launch_bounds__(LOOP_THREADS, 1) k_ploc_loop(PlocLoop A) {
    __shared__ float4 wlo[LOOP_THREADS + 2 * PLOC_MAX_RADIUS], whi[LOOP_THREADS + 2 * PLOC_MAX_RADIUS];
#if PLOC_SMEM_TAIL
    __shared__ uint32_t tail_cid[LOOP_THREADS], tail_cnt[LOOP_THREADS], tail_nn[LOOP_THREADS];
    static_assert(PLOC_TAIL <= LOOP_THREADS, "the shared-memory tail keeps one cluster per thread");
#endif
    cg::grid_group grid = cg::this_grid();
    const uint32_t nb = gridDim.x, b = blockIdx.x;
    for (uint32_t i = b * LOOP_THREADS + threadIdx.x; i < A.n; i += nb * LOOP_THREADS) {  // k_ploc_init
        A.clusters[0][i] = A.n - 1 + i;
        A.parent[A.n - 1 + i] = -1;
    }
    grid.sync();
    uint32_t m = A.n, next_node = 0, rounds = 0, status = 0;
    int cur = 0;
    PROF_DECL
    while (m > (uint32_t)PLOC_TAIL) {
        PROF("ploc grid round/m", rounds, m);
#if PLOC_FUSED_ROUNDS && PLOC_SMEM_TAIL
        const bool fused = (m + nb - 1) / nb + 4u * (uint32_t)ploc_radius_for(A.radius, m) <= (uint32_t)LOOP_THREADS;
        const uint2 t = fused ? ploc_round_fused(A, m, cur, next_node, b, nb, wlo, whi, tail_cid, tail_cnt, tail_nn, [&] { grid.sync(); })
                              : ploc_round_windows(A, m, cur, next_node, b, nb, wlo, whi, tail_nn, [&] { grid.sync(); });
#else
        const uint2 t = ploc_round(A, m, cur, next_node, b, nb, wlo, whi, [&] { grid.sync(); });
#endif
        rounds++;
        if (t.y == 0 || t.x >= m) { status = 1; break; }  // every CTA sees the same totals
        next_node += t.y;
        m = t.x;
        cur ^= 1;
    }
    if (b != 0) return;
    PROF("ploc grid done round/m", rounds, m);
#if PLOC_SMEM_TAIL
    ploc_tail_shared(A, m, next_node, rounds, status, cur, wlo, whi, tail_cid, tail_cnt, tail_nn);
#endif
    while (status == 0 && m > 1) {
        const uint2 t = ploc_round(A, m, cur, next_node, 0u, 1u, wlo, whi, [] { __syncthreads(); });
        rounds++;
        if (t.y == 0 || t.x >= m) { status = 1; break; }
        next_node += t.y;
        m = t.x;
        cur ^= 1;
    }
    PROF("ploc tail done round/m", rounds, m);
    PROF_DUMP();
    if (threadIdx.x == 0) {
        A.result[0] = next_node;
        A.result[1] = status;
        A.result[2] = rounds;
    }
}

// ---- collapse to 8-wide ----
// Binary tree over the Morton-sorted primitives, produced by either builder (Karras LBVH or PLOC):
// internal nodes [0, n-1), leaf of sorted position k = node n-1+k.
struct BinTree {
    const int32_t* left;
    const int32_t* right;
    const uint32_t* count;  // primitives below each internal node
    const float4* lo;
    const float4* hi;
    int n;     // primitives
    int root;  // 0 for the LBVH, n-2 for PLOC (last node created); n-1 when n == 1
};
MRT_D bool bin_is_leaf(const BinTree& T, int b) { return b >= T.n - 1; }
MRT_D uint32_t bin_count(const BinTree& T, int b) { return bin_is_leaf(T, b) ? 1u : T.count[b]; }
// sorted positions of the (at most MRT_MAX_LEAF_TRIS = 3) primitives below b, left to right
MRT_D uint32_t bin_gather(const BinTree& T, int b, uint32_t out[3]) {
    uint32_t n = 0;
    int stack[3];
    int sp = 0;
    stack[sp++] = b;
    while (sp) {
        int c = stack[--sp];
        if (bin_is_leaf(T, c)) {
            if (n < 3) out[n++] = (uint32_t)(c - (T.n - 1));
        } else {
            stack[sp++] = T.right[c];
            stack[sp++] = T.left[c];
        }
    }
    return n;
}
MRT_D float bin_area(const BinTree& T, int b) {
    float4 lo = T.lo[b], hi = T.hi[b];
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Choose the <= 8 slots of wide node w, which stands for binary node b.
MRT_D void collapse_expand_one(const BinTree& T, int b, uint32_t w, int32_t* slot_node, uint32_t* node_nchild, uint32_t* node_ntri) {
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

// The same choice made by EIGHT lanes per wide node (lane s of the group = slot s during the expansion, child s
// during the slot assignment): the single-thread version above walks ~500 (child, slot) pairs and keeps its arrays
// in local memory, 15-25 us of dependent work per thread and per level whatever the level's size.  Every pick
// reproduces the serial scan order -- largest area / cost first, ties to the lowest slot, then the lowest child --
// so both versions build the same node bit for bit (tested through build_device_loop 0 vs 1).
// All 32 lanes of the warp call this together; groups without a node pass active = false.
// rec (k_collapse_loop only): one 16-byte record per internal binary node -- (left | big << 31, right | big << 31, area of left,
// area of right), big = more than MRT_MAX_LEAF_TRIS primitives below -- so that opening a slot is ONE dependent load
// instead of the child index followed by its count and box (the expansion is a chain of such loads, ~17 round trips
// to L2 per node without the records, ~10 with them).  Same decisions: the areas are bin_area's own values.
MRT_D uint32_t collapse_expand_group(const BinTree& T, const uint4* rec, bool active, int b, uint32_t w, int32_t* slot_node,
                                     uint32_t* node_nchild, uint32_t* node_ntri) {  // returns the node's child count (0 if !active)
    const unsigned FULL = 0xFFFFFFFFu;
    const int s = threadIdx.x & 7;
    int slot = -1;      // binary node in slot s
    int ns = 1;
    // cached per lane: can this slot be opened, is it "big", its area
    bool inner = false, big = false;
    float area = 0.0f;
    auto refresh = [&]() {
        inner = slot >= 0 && !bin_is_leaf(T, slot);
        big = inner && T.count[slot] > MRT_MAX_LEAF_TRIS;
        area = inner ? bin_area(T, slot) : 0.0f;
    };
    auto take = [&](const uint4& r, bool right) {  // this lane's slot becomes a child of the record's node
        const uint32_t v = right ? r.y : r.x;
        slot = (int)(v & 0x7FFFFFFFu);
        inner = !bin_is_leaf(T, slot);
        big = inner && (v >> 31) != 0u;
        area = inner ? __uint_as_float(right ? r.w : r.z) : 0.0f;
    };
    if (bin_is_leaf(T, b)) {
        if (s == 0) slot = b;
        refresh();
    } else if (rec) {
        const uint4 r = rec[b];
        if (s == 0) take(r, false);
        if (s == 1) take(r, true);
        ns = 2;
    } else {
        if (s == 0) slot = T.left[b];
        if (s == 1) slot = T.right[b];
        ns = 2;
        refresh();
    }
    // phase 0: open the largest subtree that is too big to be a leaf; phase 1: spend spare slots on small leaves
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int it = 0; it < 6; it++) {
            const bool eligible = inner && (phase == 0 ? big : !big);
            float key = eligible ? area : -1.0f;
            int who = s;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const float ok = __shfl_xor_sync(FULL, key, d, 8);
                const int ow = __shfl_xor_sync(FULL, who, d, 8);
                if (ok > key || (ok == key && ow < who)) { key = ok; who = ow; }
            }
            const bool open = key >= 0.0f && ns < 8;
            if (!__any_sync(FULL, open)) break;  // no group of this warp can open a slot in this phase any more
            // the chosen slot takes its left child, slot ns its right child
            const int chosen = __shfl_sync(FULL, slot, who, 8);
            if (open) {
                if (rec) {
                    if (s == who || s == ns) take(rec[chosen], s != who);
                } else {
                    if (s == who) { slot = T.left[chosen]; refresh(); }
                    else if (s == ns) { slot = T.right[chosen]; refresh(); }
                }
                ns++;
            }
        }
    }
    // octant-aware slot assignment: slot q lies towards (q&1 ? +x : -x, q&2 ? +y : -y, q&4 ? +z : -z) of the
    // node centre, so that traversal priority (slot ^ ray octant) approximates front-to-back.  Lane s = child s.
    const float4 nlo = T.lo[b], nhi = T.hi[b];
    const float3 nc = f3(nlo.x + nhi.x, nlo.y + nhi.y, nlo.z + nhi.z);
    float3 off = f3s(0.0f);
    if (slot >= 0) {
        const float4 lo = T.lo[slot], hi = T.hi[slot];
        off = f3(lo.x + hi.x, lo.y + hi.y, lo.z + hi.z) - nc;
    }
    unsigned slot_done = 0;
    bool child_done = slot < 0;
    int my_slot = -1;
#pragma unroll 1
    for (int it = 0; it < 8; it++) {
        float bestC = -3.0e38f;
        int bs = -1;
        if (!child_done) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (slot_done & (1u << q)) continue;
                const float cost = ((q & 1) ? off.x : -off.x) + ((q & 2) ? off.y : -off.y) + ((q & 4) ? off.z : -off.z);
                if (cost > bestC) { bestC = cost; bs = q; }
            }
        }
        // first maximum in (child, slot) scan order: highest cost, then lowest child
        float key = bs >= 0 ? bestC : -3.4e38f;
        int who = bs >= 0 ? s : 8;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const float ok = __shfl_xor_sync(FULL, key, d, 8);
            const int ow = __shfl_xor_sync(FULL, who, d, 8);
            if (ok > key || (ok == key && ow < who)) { key = ok; who = ow; }
        }
        if (!__any_sync(FULL, who < 8)) break;  // every child of every group is placed
        const int win_slot = __shfl_sync(FULL, bs, who & 7, 8);
        if (who < 8) {
            if (s == who) { my_slot = win_slot; child_done = true; }
            slot_done |= 1u << win_slot;
        }
    }
    const uint32_t cnt = slot >= 0 ? bin_count(T, slot) : 0u;
    const bool is_child = slot >= 0 && cnt > MRT_MAX_LEAF_TRIS;
    uint32_t nchild = is_child ? 1u : 0u, ntri = (slot >= 0 && !is_child) ? cnt : 0u;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        nchild += __shfl_xor_sync(FULL, nchild, d, 8);
        ntri += __shfl_xor_sync(FULL, ntri, d, 8);
    }
    if (!active) return 0u;
    if (slot >= 0 && my_slot >= 0) slot_node[(size_t)w * 8 + my_slot] = slot;
    if (!(slot_done & (1u << s))) slot_node[(size_t)w * 8 + s] = -1;
    if (s == 0) {
        node_nchild[w] = nchild;
        node_ntri[w] = ntri;
    }
    return nchild;
}

// One thread per wide node of the current level.
__global__ void __launch_bounds__(128) k_collapse_expand(BinTree T, const uint2* __restrict__ items, uint32_t count,
                                                         int32_t* __restrict__ slot_node, uint32_t* __restrict__ node_nchild,
                                                         uint32_t* __restrict__ node_ntri) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    collapse_expand_one(T, (int)items[k].x, items[k].y, slot_node, node_nchild, node_ntri);
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

// ---- collapse with the level loop on the device ----
// Same idea as k_ploc_loop: the host-driven loop launches 6 kernels and reads one counter back per level of the
// wide tree; here all levels run inside one cooperative launch.  Per level: expand (one thread per node), grid
// barrier, chunked scan of the child counts fused with the emission of the next level's work items, grid barrier.
// After the last level the per-node triangle counts are scanned into node_tri_base the same way.
struct CollapseLoop {
    BinTree T;
    uint2* items[2];          // work items (binary node, wide node) of the current / next level
    int32_t* slot_node;
    uint32_t *node_nchild, *node_ntri, *node_child_base, *node_tri_base;
    uint32_t* block_sums;     // [gridDim.x]
    uint32_t* result;         // [0] wide nodes, [1] status (0 ok, 1 node budget exceeded), [2] levels
    uint32_t* level_starts;   // [MAX_WIDE_LEVELS + 1]: first node of each level, then the node count
    const uint32_t* ploc_result;  // k_ploc_loop's result words when it built the hierarchy just before (else null)
    uint4* rec;                   // [n - 1] child records of the internal binary nodes (collapse_expand_group), filled here
    uint32_t n;               // primitives = node budget
};
#define MAX_WIDE_LEVELS 1023u
#ifndef COLLAPSE_CHUNKED
#define COLLAPSE_CHUNKED 0  // A/B (1 M / 260 k / 10.4 M triangles): 1.282 vs 1.202, 0.778 vs 0.719, 8.15 vs 7.92 ms -- one barrier less per level, yet slower: off
#endif
#ifndef COLLAPSE_TOP_IN_ONE_CTA
#define COLLAPSE_TOP_IN_ONE_CTA 0  // A/B (1 M / 260 k triangles): 1.257 vs 1.265 ms, 0.763 vs 0.748 ms -- no gain, off
#endif


// exclusive scan of in[first .. first+count) across the grid; calls emit(index, exclusive prefix) for every element
// and returns the total.  Two grid barriers (counts published, then consumed); block_sums is reused by the caller.
template <class Emit>
MRT_D uint32_t grid_scan(cg::grid_group& grid, const uint32_t* in, uint32_t first, uint32_t count, uint32_t* block_sums, Emit emit) {
    const uint32_t nb = gridDim.x, b = blockIdx.x;
    const uint32_t chunk = (count + nb - 1) / nb;
    const uint32_t r0 = min(count, b * chunk), r1 = min(count, r0 + chunk);
    uint32_t local = 0;
    for (uint32_t k = r0 + threadIdx.x; k < r1; k += LOOP_THREADS) local += in[first + k];
    const uint32_t mine = cta_sum2(local, 0u).x;
    if (threadIdx.x == 0) block_sums[b] = mine;
    grid.sync();
    uint32_t before = 0, all = 0;
    for (uint32_t q = threadIdx.x; q < nb; q += LOOP_THREADS) {
        const uint32_t sq = block_sums[q];
        all += sq;
        if (q < b) before += sq;
    }
    const uint2 sums = cta_sum2(before, all);
    uint32_t base = sums.x;
    const uint32_t total = sums.y;
    for (uint32_t t0 = r0; t0 < r1; t0 += LOOP_THREADS) {
        const uint32_t k = t0 + threadIdx.x;
        const uint32_t v = k < r1 ? in[first + k] : 0u;
        uint32_t tile_total;
        const uint32_t ex = cta_scan(v, &tile_total);
        if (k < r1) emit(first + k, base + ex);
        base += tile_total;
    }
    grid.sync();
    return total;
}

__global__ void __launch_bounds__(LOOP_THREADS, 1) k_collapse_loop(CollapseLoop A) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * LOOP_THREADS + threadIdx.x, gsize = gridDim.x * LOOP_THREADS;
    uint32_t level_start = 0, level_count = 1, levels = 0, status = 0;
    int cur = 0;
    PROF_DECL
    if (gtid == 0) A.items[0][0] = make_uint2((uint32_t)A.T.root, 0u);
    // the host has not looked at k_ploc_loop's result yet: an incomplete hierarchy is not walked (status 2, every thread alike)
    if (A.ploc_result && A.n > 1 && (A.ploc_result[1] != 0u || A.ploc_result[0] != A.n - 1u)) { status = 2; level_count = 0; }
    if (status == 0 && A.rec) {
        for (uint32_t i = gtid; i + 1 < A.n; i += gsize) {
            const int l = A.T.left[i], r = A.T.right[i];
            const uint32_t bl = bin_count(A.T, l) > MRT_MAX_LEAF_TRIS ? 0x80000000u : 0u, br = bin_count(A.T, r) > MRT_MAX_LEAF_TRIS ? 0x80000000u : 0u;
            A.rec[i] = make_uint4((uint32_t)l | bl, (uint32_t)r | br, __float_as_uint(bin_area(A.T, l)), __float_as_uint(bin_area(A.T, r)));
        }
    }
    grid.sync();
#if COLLAPSE_TOP_IN_ONE_CTA
    // The top of the wide tree (levels of <= 128 nodes: the first three of a large scene, all of a small one) inside CTA 0,
    // with CTA barriers and a CTA-wide scan: a level costs its expansion chain and ~1 us instead of three grid barriers and
    // a grid-wide scan (~15 us less per level).  Same nodes in the same order.  The other CTAs wait for the state at one barrier.
    {
        uint32_t* const state = A.level_starts + MAX_WIDE_LEVELS + 1;  // level_start, level_count, levels, cur | status << 8
        if (blockIdx.x == 0) {
            while (status == 0 && level_count > 0 && level_count <= LOOP_THREADS / 8u) {
                if ((size_t)level_start + level_count > A.n) { status = 1; break; }
                if (threadIdx.x == 0 && levels < MAX_WIDE_LEVELS) A.level_starts[levels] = level_start;
                const uint2* items = A.items[cur];
                uint2* next_items = A.items[cur ^ 1];
                {
                    const uint32_t k = threadIdx.x >> 3;
                    const bool active = k < level_count;
                    const uint2 item = active ? items[k] : make_uint2((uint32_t)A.T.root, 0u);
                    collapse_expand_group(A.T, A.rec, active, (int)item.x, item.y, A.slot_node, A.node_nchild, A.node_ntri);
                }
                __syncthreads();
                const uint32_t next_start = level_start + level_count;
                const bool mine = threadIdx.x < level_count;
                const uint32_t w = level_start + threadIdx.x;
                uint32_t next_count;
                const uint32_t child_off = cta_scan(mine ? A.node_nchild[w] : 0u, &next_count);
                if (mine) {
                    const uint32_t base = next_start + child_off;
                    A.node_child_base[w] = base;
                    uint32_t rel = 0;
                    for (int s = 0; s < 8; s++) {
                        const int c = A.slot_node[(size_t)w * 8 + s];
                        if (c < 0) continue;
                        if (bin_count(A.T, c) > MRT_MAX_LEAF_TRIS) {
                            next_items[base - next_start + rel] = make_uint2((uint32_t)c, base + rel);
                            rel++;
                        }
                    }
                }
                __syncthreads();
                if ((size_t)next_start + next_count > A.n) { status = 1; break; }
                level_start = next_start;
                level_count = next_count;
                cur ^= 1;
                levels++;
            }
            if (threadIdx.x == 0) {
                state[0] = level_start; state[1] = level_count; state[2] = levels; state[3] = (uint32_t)cur | (status << 8);
            }
        }
        grid.sync();
        level_start = state[0]; level_count = state[1]; levels = state[2];
        cur = (int)(state[3] & 0xFFu); status = state[3] >> 8;
        if (status != 0) level_count = 0;
    }
#endif
    while (level_count > 0) {
        if ((size_t)level_start + level_count > A.n) { status = 1; break; }
        if (gtid == 0 && levels < MAX_WIDE_LEVELS) A.level_starts[levels] = level_start;
        const uint2* items = A.items[cur];
        uint2* next_items = A.items[cur ^ 1];
#if COLLAPSE_CHUNKED
        // Every CTA expands a contiguous range of the level -- the range whose child counts it scans afterwards -- so its
        // partial sum is known without reading the counts back: one grid barrier and one global round trip per level less
        // than expand | barrier | grid_scan (sum, barrier, scan).
        const uint32_t nbk = gridDim.x, bk = blockIdx.x;
        const uint32_t chunk = (level_count + nbk - 1) / nbk;
        const uint32_t r0 = min(level_count, bk * chunk), r1 = min(level_count, r0 + chunk);
        uint32_t local = 0;
        for (uint32_t k0 = r0; k0 < r1; k0 += LOOP_THREADS / 8u) {  // 8 lanes per node
            const uint32_t k = k0 + (threadIdx.x >> 3);
            const bool active = k < r1;
            const uint2 item = active ? items[k] : make_uint2((uint32_t)A.T.root, 0u);
            const uint32_t nc = collapse_expand_group(A.T, A.rec, active, (int)item.x, item.y, A.slot_node, A.node_nchild, A.node_ntri);
            if ((threadIdx.x & 7u) == 0u) local += nc;
        }
        const uint32_t mine = cta_sum2(local, 0u).x;
        if (threadIdx.x == 0) A.block_sums[bk] = mine;
        PROF("collapse expand(cta0) level/count", levels, level_count);
        grid.sync();
        PROF("collapse expand-wait level/count", levels, level_count);
        const uint32_t next_start = level_start + level_count;
        uint32_t next_count;
        {
            uint32_t before = 0, all = 0;
            for (uint32_t q = threadIdx.x; q < nbk; q += LOOP_THREADS) {
                const uint32_t sq = A.block_sums[q];
                all += sq;
                if (q < bk) before += sq;
            }
            const uint2 sums = cta_sum2(before, all);
            uint32_t base_off = sums.x;
            next_count = sums.y;
            for (uint32_t t0 = r0; t0 < r1; t0 += LOOP_THREADS) {
                const uint32_t k = t0 + threadIdx.x;
                const uint32_t w = level_start + k;
                const uint32_t v = k < r1 ? A.node_nchild[w] : 0u;  // written by this CTA, before the barrier
                uint32_t tile_total;
                const uint32_t ex = cta_scan(v, &tile_total);
                if (k < r1) {
                    const uint32_t base = next_start + base_off + ex;
                    A.node_child_base[w] = base;
                    uint32_t rel = 0;
                    for (int sl = 0; sl < 8; sl++) {
                        const int c = A.slot_node[(size_t)w * 8 + sl];
                        if (c < 0) continue;
                        if (bin_count(A.T, c) > MRT_MAX_LEAF_TRIS) {
                            next_items[base - next_start + rel] = make_uint2((uint32_t)c, base + rel);
                            rel++;
                        }
                    }
                }
                base_off += tile_total;
            }
            grid.sync();
        }
#else
        for (uint32_t k0 = (gtid >> 5) * 4u; k0 < level_count; k0 += (gsize >> 5) * 4u) {  // a warp takes 4 nodes, 8 lanes each
            const uint32_t k = k0 + ((threadIdx.x & 31) >> 3);
            const bool active = k < level_count;
            const uint2 item = active ? items[k] : make_uint2((uint32_t)A.T.root, 0u);
            collapse_expand_group(A.T, A.rec, active, (int)item.x, item.y, A.slot_node, A.node_nchild, A.node_ntri);
        }
        PROF("collapse expand(cta0) level/count", levels, level_count);
        grid.sync();
        PROF("collapse expand-wait level/count", levels, level_count);
        const uint32_t next_start = level_start + level_count;
        // k_collapse_emit fused into the scan of the child counts
        const uint32_t next_count = grid_scan(grid, A.node_nchild, level_start, level_count, A.block_sums,
            [&](uint32_t w, uint32_t child_off) {
                const uint32_t base = next_start + child_off;
                A.node_child_base[w] = base;
                uint32_t rel = 0;
                for (int s = 0; s < 8; s++) {
                    const int c = A.slot_node[(size_t)w * 8 + s];
                    if (c < 0) continue;
                    if (bin_count(A.T, c) > MRT_MAX_LEAF_TRIS) {
                        next_items[base - next_start + rel] = make_uint2((uint32_t)c, base + rel);
                        rel++;
                    }
                }
            });
#endif
        PROF("collapse scan+emit level/next", levels, next_count);
        if ((size_t)next_start + next_count > A.n) { status = 1; break; }
        level_start = next_start;
        level_count = next_count;
        cur ^= 1;
        levels++;
    }
    uint32_t num_nodes = level_start;
    if (status == 0)
        grid_scan(grid, A.node_ntri, 0u, num_nodes, A.block_sums, [&](uint32_t w, uint32_t off) { A.node_tri_base[w] = off; });
    PROF("collapse tri scan levels/nodes", levels, num_nodes);
    PROF_DUMP();
    if (gtid == 0) {
        A.result[0] = num_nodes;
        A.result[1] = status;
        A.result[2] = levels;
        if (levels <= MAX_WIDE_LEVELS) A.level_starts[levels] = num_nodes;
    }
}

// Quantisation grid of a wide node, per axis: 255 steps of size 2^e starting two steps below the node
// box, sized so that 250 steps span the box.  Every child plane then lies strictly inside the grid with
// room to spare, which lets the traversal fold the integer->float decode into its FMA (trace.cuh) at the
// price of a plane error of at most 1/256 step; planes are rounded outward with 1/64 step of slack.
// The step is never finer than 2 ulp of the largest coordinate magnitude m on the axis (nor than 2^-126): fp32
// coordinates cannot resolve a finer grid, the grid origin nlo - 2 step would round by more than a fraction of a
// step, and a flat node (extent 0) still gets a slab of finite thickness.  The CPU restatement of this quantiser and
// of the traversal's slab test (oracle/minote_oracle.c orc_wide_node_*) is what tests/test_oracle_wide_node.py uses
// to check, without a GPU, that no child box the exact ray touches is ever culled.
MRT_D uint32_t grid_exponent(float ext, float m) {
    // biased exponent e such that 2^(e-127) * 250 >= ext
    float s = ext / 250.0f;
    uint32_t b = __float_as_uint(s);
    uint32_t e = (b >> 23) & 0xFFu;
    if (b & 0x7FFFFFu) e += 1;
    while (e < 254u && __uint_as_float(e << 23) * 250.0f < ext) e++;  // guard against rounding in the division
    const uint32_t em = (__float_as_uint(m) >> 23) & 0xFFu;
    const uint32_t emin = em > 23u ? em - 22u : 1u;
    if (e < emin) e = emin;
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
    uint32_t ex = grid_exponent(nhi.x - nlo.x, fmaxf(fabsf(nlo.x), fabsf(nhi.x))), ey = grid_exponent(nhi.y - nlo.y, fmaxf(fabsf(nlo.y), fabsf(nhi.y))),
             ez = grid_exponent(nhi.z - nlo.z, fmaxf(fabsf(nlo.z), fabsf(nhi.z)));
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
            // outward rounding with 1/64 step of slack (the traversal's decode error is <= 1/256 step).  The plane
            // origin + q * step is evaluated in double, i.e. exactly: that is the plane the traversal's
            // (origin - o) / d + q * (step / d) sees, whereas an fp32 sum would round to the coordinates' ulp
            const double o = (double)org[a], st = (double)sc[a], slack = st * 0.015625;
            float ql = fminf(fmaxf(floorf((clo[a] - org[a]) / sc[a] - 0.02f), 0.0f), 255.0f);
            float qh = fminf(fmaxf(ceilf((chi[a] - org[a]) / sc[a] + 0.02f), 0.0f), 255.0f);
            while (ql > 0.0f && o + (double)ql * st > (double)clo[a] - slack) ql -= 1.0f;
            while (qh < 255.0f && o + (double)qh * st < (double)chi[a] + slack) qh += 1.0f;
            qlo[a][s >> 2] |= (uint32_t)ql << (8 * (s & 3));
            qhi[a][s >> 2] |= (uint32_t)qh << (8 * (s & 3));
        }
        uint32_t cnt = bin_count(T, c);
        if (cnt > MRT_MAX_LEAF_TRIS) {
            imask |= 1u << s;
        } else {
            leafmask |= ((1u << cnt) - 1u) << (3 * s);
            uint32_t where[3];
            bin_gather(T, c, where);
            for (uint32_t t = 0; t < cnt; t++) {
                uint32_t prim = order[where[t]];
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

// ---- emission and refit on the WIDE tree, level by level, eight lanes per node (option "wide_refit", default) ----
// k_emit_nodes above walks a node's eight slots in one thread: 8 x (two gathers of binary boxes, 6 quantisations with
// a double-precision outward check, up to 3 triangles copied through 12 dependent loads) -- ~100 us per million
// triangles with 1/8 of the lanes a node could use, and MRT_BUILD_REFIT pays it after a bottom-up climb of the BINARY
// tree (k_bin_boxes: one atomic + fence + dependent loads per level of a ~40-level tree, 176 us per million triangles).
// The wide tree is stored level by level (children behind their parents, each level contiguous), and a slot's box is
// the exact union -- fmin / fmax, no rounding -- of the triangles below it whichever tree it is accumulated on.  So:
//   full build:  k_emit_topology writes what the collapse decided (imask, child_base, tri_base, leafmask; the primitive id
//                of every leaf triangle), then the refit below fills in everything that depends on coordinates;
//   refit:       for each level from the deepest up, k_wide_refit -- lane s of a group of 8 owns slot s: a leaf slot
//                re-reads its <= 3 triangles from the (new) vertex positions and rewrites them, an inner slot takes the
//                box its child node stored one launch earlier; the node box, the grid exponents and the packed plane
//                words come out of 8-lane shuffles; the quantiser is k_emit_nodes' own, expression for expression.
// Nodes and leaf triangles are bit-identical to the round-1 path (tests/test_gpu_mesh.py::test_wide_refit_*).
struct WideRefit {
    WideNode* nodes;
    float4* tris;
    const float* pos;
    const uint32_t* idx;
    float4 *node_lo, *node_hi;
    uint32_t first, count;  // this level: nodes [first, first + count)
};

__global__ void __launch_bounds__(256) k_wide_refit(WideRefit A) {
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool active = g < A.count;
    const uint32_t w = A.first + (active ? g : A.count - 1u);  // idle groups shadow the last node (shuffles need all lanes)
    const int s = threadIdx.x & 7;
    const uint32_t n0w = A.nodes[w].w[0].w;
    const uint4 n1 = A.nodes[w].w[1];
    const uint32_t imask = active ? n0w >> 24 : 0u, leafmask = active ? n1.z : 0u;  // idle groups touch nothing else
    float3 lo = f3s(3.0e38f), hi = f3s(-3.0e38f);
    bool present = false;
    if ((imask >> s) & 1u) {
        const uint32_t child = n1.x + __popc(imask & ((1u << s) - 1u));
        const float4 clo = A.node_lo[child], chi = A.node_hi[child];
        lo = f3(clo.x, clo.y, clo.z);
        hi = f3(chi.x, chi.y, chi.z);
        present = true;
    } else {
        const uint32_t bits = (leafmask >> (3 * s)) & 7u;
        if (bits) {
            const uint32_t cnt = __popc(bits), first = n1.y + __popc(leafmask & ((1u << (3 * s)) - 1u));
            for (uint32_t t = 0; t < cnt; t++) {
                const size_t o = 3 * (size_t)(first + t);
                const uint32_t prim = __float_as_uint(A.tris[o].w);
                const uint32_t i0 = A.idx[3 * (size_t)prim], i1 = A.idx[3 * (size_t)prim + 1], i2 = A.idx[3 * (size_t)prim + 2];
                const float3 a = f3(A.pos[3 * (size_t)i0], A.pos[3 * (size_t)i0 + 1], A.pos[3 * (size_t)i0 + 2]);
                const float3 b = f3(A.pos[3 * (size_t)i1], A.pos[3 * (size_t)i1 + 1], A.pos[3 * (size_t)i1 + 2]);
                const float3 c = f3(A.pos[3 * (size_t)i2], A.pos[3 * (size_t)i2 + 1], A.pos[3 * (size_t)i2 + 2]);
                A.tris[o + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(prim));
                A.tris[o + 1] = make_float4(b.x, b.y, b.z, 0.0f);
                A.tris[o + 2] = make_float4(c.x, c.y, c.z, 0.0f);
                lo = f3(fminf(lo.x, fminf(a.x, fminf(b.x, c.x))), fminf(lo.y, fminf(a.y, fminf(b.y, c.y))), fminf(lo.z, fminf(a.z, fminf(b.z, c.z))));
                hi = f3(fmaxf(hi.x, fmaxf(a.x, fmaxf(b.x, c.x))), fmaxf(hi.y, fmaxf(a.y, fmaxf(b.y, c.y))), fmaxf(hi.z, fmaxf(a.z, fmaxf(b.z, c.z))));
            }
            present = true;
        }
    }
    // node box over the 8 slots
    float3 nlo = lo, nhi = hi;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        nlo.x = fminf(nlo.x, __shfl_xor_sync(FULL, nlo.x, d, 8)); nlo.y = fminf(nlo.y, __shfl_xor_sync(FULL, nlo.y, d, 8));
        nlo.z = fminf(nlo.z, __shfl_xor_sync(FULL, nlo.z, d, 8));
        nhi.x = fmaxf(nhi.x, __shfl_xor_sync(FULL, nhi.x, d, 8)); nhi.y = fmaxf(nhi.y, __shfl_xor_sync(FULL, nhi.y, d, 8));
        nhi.z = fmaxf(nhi.z, __shfl_xor_sync(FULL, nhi.z, d, 8));
    }
    const uint32_t ex = grid_exponent(nhi.x - nlo.x, fmaxf(fabsf(nlo.x), fabsf(nhi.x))), ey = grid_exponent(nhi.y - nlo.y, fmaxf(fabsf(nlo.y), fabsf(nhi.y))),
                   ez = grid_exponent(nhi.z - nlo.z, fmaxf(fabsf(nlo.z), fabsf(nhi.z)));
    const float sc[3] = {__uint_as_float(ex << 23), __uint_as_float(ey << 23), __uint_as_float(ez << 23)};
    const float org[3] = {nlo.x - 2.0f * sc[0], nlo.y - 2.0f * sc[1], nlo.z - 2.0f * sc[2]};
    // this slot's planes (k_emit_nodes' quantiser), shifted to its byte of the packed words
    uint32_t vlo[3] = {255u, 255u, 255u}, vhi[3] = {0u, 0u, 0u};
    if (present) {
        const float clo[3] = {lo.x, lo.y, lo.z}, chi[3] = {hi.x, hi.y, hi.z};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double o = (double)org[a], st = (double)sc[a], slack = st * 0.015625;
            float ql = fminf(fmaxf(floorf((clo[a] - org[a]) / sc[a] - 0.02f), 0.0f), 255.0f);
            float qh = fminf(fmaxf(ceilf((chi[a] - org[a]) / sc[a] + 0.02f), 0.0f), 255.0f);
            while (ql > 0.0f && o + (double)ql * st > (double)clo[a] - slack) ql -= 1.0f;
            while (qh < 255.0f && o + (double)qh * st < (double)chi[a] + slack) qh += 1.0f;
            vlo[a] = (uint32_t)ql;
            vhi[a] = (uint32_t)qh;
        }
    }
    uint32_t own_lo[3], own_hi[3], oth_lo[3], oth_hi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        uint32_t l = vlo[a] << (8 * (s & 3)), h = vhi[a] << (8 * (s & 3));
        l |= __shfl_xor_sync(FULL, l, 1, 8); h |= __shfl_xor_sync(FULL, h, 1, 8);
        l |= __shfl_xor_sync(FULL, l, 2, 8); h |= __shfl_xor_sync(FULL, h, 2, 8);
        own_lo[a] = l; own_hi[a] = h;
        oth_lo[a] = __shfl_xor_sync(FULL, l, 4, 8); oth_hi[a] = __shfl_xor_sync(FULL, h, 4, 8);
    }
    if (!active) return;
    const bool low = s < 4;  // this lane's own words are those of slots 0..3
#define QLO(a, h) ((h == 0) == low ? own_lo[a] : oth_lo[a])
#define QHI(a, h) ((h == 0) == low ? own_hi[a] : oth_hi[a])
    uint4* const out = A.nodes[w].w;
    if (s == 0) {
        out[0] = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]), ex | (ey << 8) | (ez << 16) | (imask << 24));
        A.node_lo[w] = make_float4(nlo.x, nlo.y, nlo.z, 0.0f);
        A.node_hi[w] = make_float4(nhi.x, nhi.y, nhi.z, 0.0f);
    } else if (s == 1) {
        out[2] = make_uint4(QLO(0, 0), QLO(0, 1), QLO(1, 0), QLO(1, 1));
    } else if (s == 2) {
        out[3] = make_uint4(QLO(2, 0), QLO(2, 1), QHI(0, 0), QHI(0, 1));
    } else if (s == 3) {
        out[4] = make_uint4(QHI(1, 0), QHI(1, 1), QHI(2, 0), QHI(2, 1));
    }
#undef QLO
#undef QHI
}

// What the collapse decided, per wide node (8 lanes, lane s = slot s): imask, child_base, tri_base, leafmask, and the
// primitive id of every leaf triangle in leaf order.  Coordinates follow from k_wide_refit.
__global__ void __launch_bounds__(256)
k_emit_topology(BinTree T, uint32_t num_nodes, const int32_t* __restrict__ slot_node, const uint32_t* __restrict__ node_child_base,
                const uint32_t* __restrict__ node_tri_base, const uint32_t* __restrict__ order, WideNode* __restrict__ nodes,
                float4* __restrict__ tris) {
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool active = g < num_nodes;
    const uint32_t w = active ? g : num_nodes - 1u;
    const int s = threadIdx.x & 7;
    const int c = slot_node[(size_t)w * 8 + s];
    const uint32_t cnt = c < 0 ? 0u : bin_count(T, c);
    const bool inner = cnt > MRT_MAX_LEAF_TRIS;
    uint32_t imask = inner ? 1u << s : 0u;
    uint32_t leafmask = (!inner && cnt) ? ((1u << cnt) - 1u) << (3 * s) : 0u;
    const uint32_t ntri = inner ? 0u : cnt;
    uint32_t inc = ntri;  // inclusive prefix of the leaf triangles over the slots
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        imask |= __shfl_xor_sync(FULL, imask, d, 8);
        leafmask |= __shfl_xor_sync(FULL, leafmask, d, 8);
        const uint32_t t = __shfl_up_sync(FULL, inc, d, 8);
        if (s >= d) inc += t;
    }
    if (!active) return;
    const uint32_t tri_base = node_tri_base[w];
    if (ntri) {
        uint32_t where[3];
        bin_gather(T, c, where);
        for (uint32_t t = 0; t < ntri; t++)
            tris[3 * (size_t)(tri_base + inc - ntri + t)] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(order[where[t]]));
    }
    if (s == 0) {
        nodes[w].w[0] = make_uint4(0u, 0u, 0u, imask << 24);
        nodes[w].w[1] = make_uint4(node_child_base[w], tri_base, leafmask, 0u);
    }
}

// SAH cost of the wide tree (quality metric reported in mrt_stats.sah_cost): sum over wide nodes of
// area(node) / area(root) * 1 (one node step) + sum over leaf slots of area(slot box) / area(root) * ntris.
__global__ void __launch_bounds__(256) k_sah_cost(const WideNode* __restrict__ nodes, uint32_t num_nodes, float* __restrict__ out) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    float node_cost = 0.0f, tri_cost = 0.0f;
    if (w < num_nodes) {
        const uint4 n0 = nodes[w].w[0], n1 = nodes[w].w[1], n2 = nodes[w].w[2], n3 = nodes[w].w[3], n4 = nodes[w].w[4];
        float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
              sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
        unsigned imask = n0.w >> 24, leafmask = n1.z;
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        for (int s = 0; s < 8; s++) {
            unsigned sh = 8 * (s & 3);
            float qlx = (float)(((s < 4 ? n2.x : n2.y) >> sh) & 0xFF), qly = (float)(((s < 4 ? n2.z : n2.w) >> sh) & 0xFF);
            float qlz = (float)(((s < 4 ? n3.x : n3.y) >> sh) & 0xFF), qhx = (float)(((s < 4 ? n3.z : n3.w) >> sh) & 0xFF);
            float qhy = (float)(((s < 4 ? n4.x : n4.y) >> sh) & 0xFF), qhz = (float)(((s < 4 ? n4.z : n4.w) >> sh) & 0xFF);
            if (qlx > qhx) continue;  // empty slot
            float dx = (qhx - qlx) * sx, dy = (qhy - qly) * sy, dz = (qhz - qlz) * sz;
            float area = dx * dy + dy * dz + dz * dx;
            lo[0] = fminf(lo[0], qlx * sx); hi[0] = fmaxf(hi[0], qhx * sx);
            lo[1] = fminf(lo[1], qly * sy); hi[1] = fmaxf(hi[1], qhy * sy);
            lo[2] = fminf(lo[2], qlz * sz); hi[2] = fmaxf(hi[2], qhz * sz);
            if (!(imask & (1u << s))) tri_cost += area * (float)__popc((leafmask >> (3 * s)) & 7u);
        }
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        node_cost = dx * dy + dy * dz + dz * dx;
        if (w == 0) out[2] = node_cost;  // root area
    }
    for (int off = 16; off > 0; off >>= 1) {
        node_cost += __shfl_down_sync(0xFFFFFFFFu, node_cost, off);
        tri_cost += __shfl_down_sync(0xFFFFFFFFu, tri_cost, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], node_cost);
        atomicAdd(&out[1], tri_cost);
    }
}

__global__ void k_last_total(const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt, uint32_t last, uint32_t* out) {
    out[0] = off[last] + cnt[last];
}

BinTree make_tree(mrt_context* ctx) {
    BinTree T;
    T.left = ctx->bin_left.p; T.right = ctx->bin_right.p; T.count = ctx->bin_count.p; T.root = ctx->bin_root;
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

int build_ploc(mrt_context* ctx) {
    const uint32_t n = ctx->ntris;
    for (int k = 0; k < 2; k++) {
        MRT_TRY(dev_reserve(ctx, ctx->ploc_c[k], n));
        MRT_TRY(dev_reserve(ctx, ctx->ploc_flag[k], n));
        MRT_TRY(dev_reserve(ctx, ctx->ploc_scan[k], n));
    }
    MRT_TRY(dev_reserve(ctx, ctx->ploc_nn, n));
    k_leaf_boxes<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->prim_lo.p, ctx->prim_hi.p, ctx->order.p, n, ctx->bin_lo.p,
                                                          ctx->bin_hi.p);
    MRT_LAUNCHED(ctx);
    uint32_t next_node = 0;
    if (ctx->opt_build_device_loop) {
        // every round inside one cooperative launch (k_ploc_loop); k_ploc_init is folded into it
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        MRT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ploc_loop, LOOP_THREADS, 0));
        if (per_sm < 1) return mrt_fail(ctx, MRT_ERR_CUDA, "k_ploc_loop does not fit on an SM");
        unsigned grid = max(1u, min((unsigned)sms, div_up(n, LOOP_THREADS)));  // one CTA per SM
        MRT_TRY(dev_reserve(ctx, ctx->loop_sums, grid));
        PlocLoop A;
        A.n = n; A.radius = ctx->opt_ploc_radius;
        A.clusters[0] = ctx->ploc_c[0].p; A.clusters[1] = ctx->ploc_c[1].p; A.nn = ctx->ploc_nn.p;
        A.left = ctx->bin_left.p; A.right = ctx->bin_right.p; A.parent = ctx->bin_parent.p; A.count = ctx->bin_count.p;
        A.lo = ctx->bin_lo.p; A.hi = ctx->bin_hi.p; A.block_sums = ctx->loop_sums.p; A.result = ctx->counters.p + 4;
        void* args[] = {&A};
        MRT_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_ploc_loop, dim3(grid), dim3(LOOP_THREADS), args, 0, ctx->stream));
        MRT_LAUNCHED(ctx);
        // No readback here: the result words (counters[4..6]) are checked by k_collapse_loop on the device -- it refuses a
        // hierarchy that is not complete -- and by the host together with the collapse's own result: one round trip less
        ctx->bin_root = (int)n - 2;
        return MRT_OK;
    } else {
    k_ploc_init<<<div_up(n, 256), 256, 0, ctx->stream>>>(n, ctx->ploc_c[0].p, ctx->bin_parent.p);
    MRT_LAUNCHED(ctx);
    uint32_t m = n;
    int cur = 0;
    while (m > 1) {
        const unsigned g = div_up(m, 256);
        k_ploc_nearest<<<g, 256, 0, ctx->stream>>>(ctx->ploc_c[cur].p, m, ctx->opt_ploc_radius, ctx->bin_lo.p, ctx->bin_hi.p,
                                                   ctx->ploc_nn.p);
        MRT_LAUNCHED(ctx);
        k_ploc_flags<<<g, 256, 0, ctx->stream>>>(ctx->ploc_nn.p, m, ctx->ploc_flag[0].p, ctx->ploc_flag[1].p);
        MRT_LAUNCHED(ctx);
        MRT_TRY(scan_exclusive_u32(ctx, ctx->ploc_flag[0].p, ctx->ploc_scan[0].p, m));
        MRT_TRY(scan_exclusive_u32(ctx, ctx->ploc_flag[1].p, ctx->ploc_scan[1].p, m));
        k_ploc_merge<<<g, 256, 0, ctx->stream>>>(ctx->ploc_c[cur].p, ctx->ploc_nn.p, m, ctx->ploc_flag[0].p, ctx->ploc_flag[1].p,
                                                 ctx->ploc_scan[0].p, ctx->ploc_scan[1].p, next_node, (int)n, ctx->bin_left.p,
                                                 ctx->bin_right.p, ctx->bin_parent.p, ctx->bin_count.p, ctx->bin_lo.p,
                                                 ctx->bin_hi.p, ctx->ploc_c[cur ^ 1].p, ctx->counters.p);
        MRT_LAUNCHED(ctx);
        uint32_t totals[2] = {0, 0};
        MRT_CUDA(ctx, cudaMemcpyAsync(totals, ctx->counters.p, sizeof totals, cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (totals[1] == 0 || totals[0] >= m) return mrt_fail(ctx, MRT_ERR_INVALID, "PLOC made no progress at %u clusters", m);
        next_node += totals[1];
        m = totals[0];
        cur ^= 1;
    }
    }
    if (next_node != n - 1) return mrt_fail(ctx, MRT_ERR_INVALID, "PLOC produced %u internal nodes for %u primitives", next_node, n);
    ctx->bin_root = (int)n - 2;
    return MRT_OK;
}

// boxes, planes and leaf triangles of every level, deepest first (the topology words are in place)
int wide_refit_levels(mrt_context* ctx, cudaStream_t stream) {
    MRT_TRY(dev_reserve(ctx, ctx->node_lo, ctx->num_nodes));
    MRT_TRY(dev_reserve(ctx, ctx->node_hi, ctx->num_nodes));
    WideRefit A;
    A.nodes = ctx->nodes.p; A.tris = ctx->tris.p; A.pos = ctx->pos.p; A.idx = ctx->idx.p;
    A.node_lo = ctx->node_lo.p; A.node_hi = ctx->node_hi.p;
    for (size_t l = ctx->level_starts.size() - 1; l-- > 0;) {
        A.first = ctx->level_starts[l];
        A.count = ctx->level_starts[l + 1] - A.first;
        if (A.count == 0) continue;
        k_wide_refit<<<div_up(A.count, 256 / 8), 256, 0, stream>>>(A);
        MRT_LAUNCHED(ctx);
    }
    return mrt_check_cuda(ctx, cudaGetLastError(), "wide_refit");
}

int emit_nodes(mrt_context* ctx) {
    if (ctx->opt_wide_refit && ctx->level_starts.size() >= 2 && ctx->level_starts.back() == ctx->num_nodes) {
        k_emit_topology<<<div_up(ctx->num_nodes, 256 / 8), 256, 0, ctx->stream>>>(make_tree(ctx), ctx->num_nodes, ctx->slot_node.p,
                                                                                  ctx->node_child_base.p, ctx->node_tri_base.p,
                                                                                  ctx->order.p, ctx->nodes.p, ctx->tris.p);
        MRT_LAUNCHED(ctx);
        return wide_refit_levels(ctx, ctx->stream);
    }
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
    ctx->alt_valid = false;  // option async_update: the second copy of the tree holds another topology now
    ctx->num_nodes = 0;
    ctx->num_leaf_tris = 0;
    if (n == 0) {
        ctx->bvh_valid = true;
        return MRT_OK;
    }
    cudaEventRecord(ctx->ev[12], ctx->stream);
    MRT_TRY(dev_reserve(ctx, ctx->prim_lo, n));
    MRT_TRY(dev_reserve(ctx, ctx->prim_hi, n));
    MRT_TRY(dev_reserve(ctx, ctx->keys, n));
    MRT_TRY(dev_reserve(ctx, ctx->keys_alt, n));
    MRT_TRY(dev_reserve(ctx, ctx->order, n));
    MRT_TRY(dev_reserve(ctx, ctx->order_alt, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_left, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_right, n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_parent, 2 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->bin_count, n));
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
    if (ctx->counters.p == nullptr) {  // first use: not every word is written by every builder, all are read back
        MRT_TRY(dev_reserve(ctx, ctx->counters, 16));
        MRT_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(uint32_t), ctx->stream));
    }

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
    ctx->bin_root = 0;
    if (n > 1 && ctx->opt_builder == 1) {
        MRT_TRY(build_ploc(ctx));  // boxes are produced by the merges
    } else {
        if (n > 1) {
            k_karras<<<div_up(n - 1, 256), 256, 0, ctx->stream>>>(ctx->keys.p, (int)n, ctx->bin_left.p, ctx->bin_right.p,
                                                                  ctx->bin_parent.p, ctx->bin_count.p);
            MRT_LAUNCHED(ctx);
        }
        MRT_TRY(climb_boxes(ctx));
    }

    // 6: collapse into 8-wide nodes, level by level; 7a: triangle ranges
    BinTree T = make_tree(ctx);
    if (ctx->opt_build_device_loop) {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        MRT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse_loop, LOOP_THREADS, 0));
        if (per_sm < 1) return mrt_fail(ctx, MRT_ERR_CUDA, "k_collapse_loop does not fit on an SM");
        unsigned grid = max(1u, min((unsigned)sms, div_up(n, LOOP_THREADS / 8)));  // one CTA per SM, 8 lanes per wide node
        MRT_TRY(dev_reserve(ctx, ctx->loop_sums, grid));
        CollapseLoop A;
        A.T = T;
        A.items[0] = ctx->work_a.p; A.items[1] = ctx->work_b.p;
        A.slot_node = ctx->slot_node.p; A.node_nchild = ctx->node_nchild.p; A.node_ntri = ctx->node_ntri.p;
        A.node_child_base = ctx->node_child_base.p; A.node_tri_base = ctx->node_tri_base.p;
        A.block_sums = reinterpret_cast<uint32_t*>(ctx->loop_sums.p); A.result = ctx->counters.p; A.n = n;
        A.ploc_result = (n > 1 && ctx->opt_builder == 1) ? ctx->counters.p + 4 : nullptr;
        MRT_TRY(dev_reserve(ctx, ctx->bin_rec, n));
        A.rec = ctx->bin_rec.p;
        if (ctx->level_starts_dev.p == nullptr) {  // first use: the tail of the table beyond the tree's depth is copied to the host too
            MRT_TRY(dev_reserve(ctx, ctx->level_starts_dev, MAX_WIDE_LEVELS + 1 + 8));  // + the state CTA 0 hands to the grid
            MRT_CUDA(ctx, cudaMemsetAsync(ctx->level_starts_dev.p, 0, sizeof(uint32_t) * (MAX_WIDE_LEVELS + 1 + 8), ctx->stream));
        }
        A.level_starts = ctx->level_starts_dev.p;
        void* args[] = {&A};
        MRT_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_collapse_loop, dim3(grid), dim3(LOOP_THREADS), args, 0, ctx->stream));
        MRT_LAUNCHED(ctx);
        // results and level table land in page-locked memory: a pageable destination is staged by the driver (~15 us more)
        if (!ctx->build_results_host) MRT_CUDA(ctx, cudaHostAlloc(&ctx->build_results_host, sizeof(uint32_t) * (8 + MAX_WIDE_LEVELS + 1), cudaHostAllocDefault));
        uint32_t* const res = ctx->build_results_host;
        MRT_CUDA(ctx, cudaMemcpyAsync(res, ctx->counters.p, sizeof(uint32_t) * 7, cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(ctx, cudaMemcpyAsync(res + 8, ctx->level_starts_dev.p, sizeof(uint32_t) * (MAX_WIDE_LEVELS + 1), cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->level_starts.assign(res + 8, res + 8 + MAX_WIDE_LEVELS + 1);
        if (A.ploc_result && res[5] != 0) return mrt_fail(ctx, MRT_ERR_INVALID, "PLOC made no progress (round %u)", res[6]);
        if (A.ploc_result && res[4] != n - 1) return mrt_fail(ctx, MRT_ERR_INVALID, "PLOC produced %u internal nodes for %u primitives", res[4], n);
        if (res[1] != 0) return mrt_fail(ctx, MRT_ERR_INVALID, "wide BVH node budget exceeded");
        ctx->num_nodes = res[0];
        ctx->num_leaf_tris = n;
        if (res[2] <= MAX_WIDE_LEVELS) ctx->level_starts.resize(res[2] + 1);
        else ctx->level_starts.clear();  // deeper than the table: emission falls back to one thread per node
    } else {
    uint2 root = make_uint2((uint32_t)ctx->bin_root /* n == 1: the single leaf is binary node n-1 = 0 */, 0u);
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->work_a.p, &root, sizeof root, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t level_start = 0, level_count = 1;
    DevArray<uint2>* cur = &ctx->work_a;
    DevArray<uint2>* nxt = &ctx->work_b;
    ctx->level_starts.clear();
    while (level_count > 0) {
        if ((size_t)level_start + level_count > n) return mrt_fail(ctx, MRT_ERR_INVALID, "wide BVH node budget exceeded");
        ctx->level_starts.push_back(level_start);
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
    ctx->level_starts.push_back(level_start);
    MRT_TRY(scan_exclusive_u32(ctx, ctx->node_ntri.p, ctx->node_tri_base.p, ctx->num_nodes));
    }

    // 7b: final nodes
    MRT_TRY(dev_reserve(ctx, ctx->nodes, ctx->num_nodes));
    MRT_TRY(dev_reserve(ctx, ctx->tris, 3 * (size_t)n));
    MRT_TRY(emit_nodes(ctx));
    cudaEventRecord(ctx->ev[13], ctx->stream);
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev[12], ctx->ev[13]);
    ctx->build_time_pending = false;
    ctx->stats.num_triangles = n;
    ctx->stats.num_wide_nodes = ctx->num_nodes;
    ctx->stats.bvh_bytes = (uint64_t)ctx->num_nodes * sizeof(WideNode) + (uint64_t)n * 48u;
    {
        float* sah = reinterpret_cast<float*>(ctx->counters.p + 12);
        float h[3] = {0, 0, 0};
        cudaMemsetAsync(sah, 0, 3 * sizeof(float), ctx->stream);
        k_sah_cost<<<div_up(ctx->num_nodes, 256), 256, 0, ctx->stream>>>(ctx->nodes.p, ctx->num_nodes, sah);
        MRT_LAUNCHED(ctx);
        MRT_CUDA(ctx, cudaMemcpyAsync(h, sah, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stats.sah_node_cost = h[2] > 0.0f ? h[0] / h[2] : 0.0f;
        ctx->stats.sah_tri_cost = h[2] > 0.0f ? h[1] / h[2] : 0.0f;
    }
    ctx->bvh_valid = true;
    return MRT_OK;
}

// the refit that needs no host round trip: topology in place, level table known
bool bvh_refit_can_be_async(const mrt_context* ctx) {
    return ctx->bvh_valid && ctx->num_nodes != 0 && ctx->opt_wide_refit && ctx->level_starts.size() >= 2 && ctx->level_starts.back() == ctx->num_nodes;
}

int bvh_refit(mrt_context* ctx, bool wait, cudaStream_t side) {
    if (!ctx->bvh_valid || ctx->num_nodes == 0) return bvh_build_full(ctx);
    if (!wait) {  // asynchronous (bvh_refit_can_be_async): queued on the caller's scene-update stream; mrt_stats_get reads the time later
        cudaEventRecord(ctx->ev[12], side);
        MRT_TRY(wide_refit_levels(ctx, side));
        cudaEventRecord(ctx->ev[13], side);
        ctx->build_time_pending = true;
        return MRT_OK;
    }
    cudaEventRecord(ctx->ev[12], ctx->stream);
    if (ctx->opt_wide_refit && ctx->level_starts.size() >= 2 && ctx->level_starts.back() == ctx->num_nodes) {
        MRT_TRY(wide_refit_levels(ctx, ctx->stream));  // the topology words and the triangles' primitive ids are in place
    } else {
        MRT_TRY(compute_boxes(ctx));
        MRT_TRY(climb_boxes(ctx));
        MRT_TRY(emit_nodes(ctx));
    }
    cudaEventRecord(ctx->ev[13], ctx->stream);
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_build, ctx->ev[12], ctx->ev[13]);
    ctx->build_time_pending = false;
    return MRT_OK;
}
