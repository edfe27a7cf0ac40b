// sort.cu -- device-wide exclusive scan and LSD radix sort (64-bit keys, 32-bit payload).
// Used by the LBVH builder (Morton keys, row n2) and by the optional ray sort between bounces
// (row n5).  Warp-level primitives do the ranking: __match_any_sync groups lanes with the same
// digit, popc of the lower-lane mask gives a stable intra-warp rank, per-warp digit counters in
// shared memory are then prefix-summed across the 8 warps of a tile.  HBM-streaming kernels.
#include <cooperative_groups.h>

#include "context.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one value per thread across a block; returns block total through `total`
template <int THREADS>
MRT_D uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[THREADS / 32];
    __shared__ uint32_t block_total;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, off);
        if (lane >= (unsigned)off) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, winc, off);
            if (lane >= (unsigned)off) winc += n;
        }
        if (lane < THREADS / 32) warp_sums[lane] = winc - w;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    uint32_t r = warp_sums[warp] + inc - v;
    *total = block_total;
    __syncthreads();
    return r;
}

// in may alias out (in-place scan): every thread reads its items before any thread writes
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* in, uint32_t* out, uint32_t* tile_sums, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = base + j < n ? in[base + j] : 0u;
        sum += v[j];
    }
    uint32_t total;
    uint32_t prefix = block_exclusive_scan<SCAN_THREADS>(sum, &total);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = prefix;
        prefix += v[j];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* sums, uint32_t m) {
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < m ? sums[i] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan<1024>(v, &total);
        uint32_t carry = carry_s;
        if (i < m) sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums,
                                                           size_t n) {
    uint32_t add = tile_sums[blockIdx.x];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] += add;
}

// ---- radix sort, 8-bit digits ----
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per CTA
constexpr int RS_WARP_SPAN = 32 * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, size_t n, int shift,
                                                        uint32_t* __restrict__ hist, uint32_t num_tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        size_t i = base + (size_t)j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out, size_t n, int shift, const uint32_t* __restrict__ offsets, uint32_t num_tiles) {
    __shared__ uint32_t wc[RS_WARPS][256];
    __shared__ uint32_t gbase[256], texcl[256];
    __shared__ uint64_t skeys[RS_TILE];  // the tile in digit order: a warp's stores then walk runs of equal digits
    __shared__ uint32_t svals[RS_TILE];  // (8 keys on average at 2048 keys per tile) instead of 32 different sectors
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
    __syncthreads();

    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], loc[RS_ITEMS];
    const size_t tbase = (size_t)blockIdx.x * RS_TILE, wbase = tbase + (size_t)warp * RS_WARP_SPAN;
    const uint32_t nvalid = (uint32_t)min((size_t)RS_TILE, n - min(n, tbase));
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        bool valid = i < n;
        key[j] = valid ? keys_in[i] : 0ull;
        val[j] = valid ? vals_in[i] : 0u;
        unsigned d = valid ? ((unsigned)(key[j] >> shift) & 255u) : (256u + lane);
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) {
            base = wc[warp][d];
            wc[warp][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        loc[j] = base + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    uint32_t run = 0;
    {
        // exclusive prefix over the tile's warps for digit = threadIdx.x
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = wc[w][threadIdx.x];
            wc[w][threadIdx.x] = run;
            run += c;
        }
    }
    {
        uint32_t all;
        const uint32_t ex = block_exclusive_scan<RS_THREADS>(run, &all);  // where digit threadIdx.x starts inside the tile
        texcl[threadIdx.x] = ex;
        // global position of the tile's entry i of this digit: gbase + i (modulo 2^32)
        gbase[threadIdx.x] = offsets[(size_t)threadIdx.x * num_tiles + blockIdx.x] - ex;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        if (i < n) {
            unsigned d = (unsigned)(key[j] >> shift) & 255u;
            const uint32_t lp = texcl[d] + wc[warp][d] + loc[j];
            skeys[lp] = key[j];
            svals[lp] = val[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const uint32_t i = (uint32_t)j * RS_THREADS + threadIdx.x;
        if (i < nvalid) {
            const uint64_t k = skeys[i];
            const size_t pos = (size_t)(uint32_t)(gbase[(unsigned)(k >> shift) & 255u] + i);
            keys_out[pos] = k;
            vals_out[pos] = svals[i];
        }
    }
}

// ---- all passes of the sort in ONE cooperative launch (inputs whose tiles are all co-resident) ----
// The launch-by-launch path above spends 5 launches per 8-bit pass (histogram, three scan kernels, scatter): at one
// million keys each pass moves 24 MB -- 4 us of HBM time -- inside ~36 us of launch and drain latency, 8 times over.
// Here one CTA owns one tile of 2048 keys for the whole sort and the passes are separated by grid barriers instead of
// launches.  Per pass: rank the tile's keys (same __match_any_sync ranking as k_rs_scatter), publish the tile's 256
// digit counts and put the tile in digit order in shared memory | barrier | every CTA derives its 256 global offsets
// from the counts of all (<= 148) tiles and a block scan of the digit totals, and writes its runs out | barrier.  No spinning on other CTAs' flags (nothing that could hang the device), and the result is the same
// stable permutation, bit for bit, as the multi-launch path (tests/test_gpu_mesh.py::test_fused_sort_*).
struct FusedSort {
    uint64_t* keys[2];
    uint32_t* vals[2];
    uint32_t* counts;   // [tiles][256]
    uint32_t* groups;   // [ceil(tiles / 32)][256]
    uint32_t n, tiles;
    int begin_bit, end_bit;
};

// THREADS = 1024 (8192 keys per CTA) for large inputs: the grid barrier costs one arrival per CTA (148 x 1024 threads:
// ~2.5 us; 510 x 256: ~5 us, measured); THREADS = 256 while that still leaves most SMs without a tile.
// Before a tile's keys leave for HBM they are put in digit order in shared memory, so that the stores of a warp walk
// through runs of equal digits (32 keys on average at 8192 keys per tile) instead of hitting 32 different sectors.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_rs_fused(FusedSort A) {
    constexpr int WARPS = THREADS / 32, TILE = THREADS * RS_ITEMS;
    extern __shared__ __align__(16) unsigned char fused_smem[];
    uint64_t* const skeys = reinterpret_cast<uint64_t*>(fused_smem);                    // [TILE]
    uint32_t* const svals = reinterpret_cast<uint32_t*>(skeys + TILE);                  // [TILE]
    uint32_t (*const wc)[256] = reinterpret_cast<uint32_t (*)[256]>(svals + TILE);      // [WARPS][256]
    __shared__ uint32_t gbase[256], texcl[256];
    cg::grid_group grid = cg::this_grid();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = blockIdx.x, d = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t tbase = (size_t)t * TILE, wbase = tbase + (size_t)warp * RS_WARP_SPAN;
    const uint32_t nvalid = (uint32_t)min((size_t)TILE, (size_t)A.n - min((size_t)A.n, tbase));
    int cur = 0;
    for (int shift = A.begin_bit; shift < A.end_bit; shift += 8, cur ^= 1) {
        const uint64_t* keys_in = A.keys[cur];
        const uint32_t* vals_in = A.vals[cur];
        for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) (&wc[0][0])[i] = 0;
        __syncthreads();
        uint64_t key[RS_ITEMS];
        uint32_t val[RS_ITEMS], loc[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const size_t i = wbase + (size_t)j * 32 + lane;
            const bool valid = i < A.n;
            key[j] = valid ? keys_in[i] : 0ull;
            val[j] = valid ? vals_in[i] : 0u;
            const unsigned dig = valid ? ((unsigned)(key[j] >> shift) & 255u) : (256u + lane);
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, dig);
            const unsigned leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (valid && lane == leader) {
                base = wc[warp][dig];
                wc[warp][dig] = base + __popc(peers);
            }
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            loc[j] = base + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        uint32_t mine = 0;
        if (d < 256u) {   // exclusive prefix over the tile's warps for digit d; the tile's count of d goes out
#pragma unroll 8
            for (int w = 0; w < WARPS; w++) {
                const uint32_t c = wc[w][d];
                wc[w][d] = mine;
                mine += c;
            }
            A.counts[(size_t)t * 256 + d] = mine;
        }
        {
            uint32_t all;
            const uint32_t ex = block_exclusive_scan<THREADS>(mine, &all);  // where digit d starts inside the tile
            if (d < 256u) texcl[d] = ex;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const size_t i = wbase + (size_t)j * 32 + lane;
            if (i < A.n) {
                const unsigned dig = (unsigned)(key[j] >> shift) & 255u;
                const uint32_t lp = texcl[dig] + wc[warp][dig] + loc[j];
                skeys[lp] = key[j];
                svals[lp] = val[j];
            }
        }
        grid.sync();
        {
            // global offsets of this tile: for digit dd, the keys with a smaller digit anywhere plus the keys with digit dd
            // in the tiles before this one.  The THREADS / 256 thread groups take the tiles round robin, so that every
            // thread has all its (<= 37) loads in flight at once instead of walking 148 tiles one round trip at a time.
            constexpr int Q = THREADS / 256;
            const unsigned q = threadIdx.x >> 8, dd = threadIdx.x & 255u;
            uint32_t total = 0, before = 0;
#pragma unroll 8
            for (uint32_t k = q; k < A.tiles; k += Q) {
                const uint32_t v = A.counts[(size_t)k * 256 + dd];
                total += v;
                before += k < t ? v : 0u;
            }
            if (Q > 1) {  // combine the groups' partial sums through wc (free again: the tile sits in skeys / svals)
                __syncthreads();
                wc[2 * q][dd] = total;
                wc[2 * q + 1][dd] = before;
                __syncthreads();
                if (q == 0) {
#pragma unroll
                    for (int r = 1; r < Q; r++) { total += wc[2 * r][dd]; before += wc[2 * r + 1][dd]; }
                } else {
                    total = 0;
                }
            }
            uint32_t all;
            const uint32_t ex = block_exclusive_scan<THREADS>(total, &all);  // keys with a smaller digit (threads >= 256 add 0)
            if (d < 256u) gbase[d] = ex + before - texcl[d];  // global position of the tile's entry i of digit d: gbase[d] + i
        }
        __syncthreads();
        uint64_t* keys_out = A.keys[cur ^ 1];
        uint32_t* vals_out = A.vals[cur ^ 1];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const uint32_t i = (uint32_t)j * THREADS + threadIdx.x;
            if (i < nvalid) {
                const uint64_t k = skeys[i];
                const size_t pos = (size_t)(gbase[(unsigned)(k >> shift) & 255u] + i);  // modulo 2^32: gbase may have wrapped below 0
                keys_out[pos] = k;
                vals_out[pos] = svals[i];
            }
        }
        grid.sync();  // the next pass reads what this one wrote; counts / groups are reused
    }
}

template <int THREADS>
int launch_fused_sort(mrt_context* ctx, FusedSort& A, unsigned tiles) {
    constexpr int TILE = THREADS * RS_ITEMS;
    const size_t smem = (size_t)TILE * 12 + (size_t)(THREADS / 32) * 256 * 4;
    MRT_CUDA(ctx, cudaFuncSetAttribute(k_rs_fused<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&A};
    MRT_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_rs_fused<THREADS>, dim3(tiles), dim3(THREADS), args, smem, ctx->stream));
    MRT_LAUNCHED(ctx);
    return MRT_OK;
}

}  // namespace

int scan_exclusive_u32(mrt_context* ctx, const uint32_t* in, uint32_t* out, size_t n) {
    if (n == 0) return MRT_OK;
    unsigned tiles = div_up(n, SCAN_TILE);
    MRT_TRY(dev_reserve(ctx, ctx->scan_tmp, tiles));
    k_scan_tiles<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, ctx->scan_tmp.p, n);
    MRT_LAUNCHED(ctx);
    if (tiles > 1) {
        k_scan_sums<<<1, 1024, 0, ctx->stream>>>(ctx->scan_tmp.p, tiles);
        MRT_LAUNCHED(ctx);
        k_scan_add<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(out, ctx->scan_tmp.p, n);
        MRT_LAUNCHED(ctx);
    }
    return mrt_check_cuda(ctx, cudaGetLastError(), "scan_exclusive_u32");
}

// Sorts (keys, vals) ascending by key bits [begin_bit, end_bit), stable.  Ping-pongs between the
// primary and _alt buffers; *result_in_alt tells where the sorted data ended up.
int radix_sort_pairs_u64(mrt_context* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t n,
                         int begin_bit, int end_bit, bool* result_in_alt) {
    *result_in_alt = false;
    if (n == 0) return MRT_OK;
    unsigned tiles = div_up(n, RS_TILE);
    if (ctx->opt_fused_sort && begin_bit < end_bit) {
        // one cooperative launch, one CTA per SM at most: 256-thread CTAs (2048 keys) up to 148 tiles, 1024-thread CTAs
        // (8192 keys) up to 1.2 M keys on a B200; beyond that the launch-by-launch passes below
        if (ctx->fused_sort_capacity < 0) cudaDeviceGetAttribute(&ctx->fused_sort_capacity, cudaDevAttrMultiProcessorCount, ctx->device);
        const unsigned small = div_up(n, (size_t)256 * RS_ITEMS), large = div_up(n, (size_t)1024 * RS_ITEMS);
#ifndef FUSED_SORT_SMALL_MAX
#define FUSED_SORT_SMALL_MAX 1000000  // tiles of 2048 keys above which the 8192-key variant is used even though the small one would fit (A/B knob)
#endif
        const unsigned ftiles = ((int)small <= ctx->fused_sort_capacity && small <= (unsigned)FUSED_SORT_SMALL_MAX) ? small : large;
        if ((int)ftiles <= ctx->fused_sort_capacity) {
            const unsigned ngroups = div_up(ftiles, 32u);
            MRT_TRY(dev_reserve(ctx, ctx->hist, (size_t)256 * (ftiles + ngroups)));
            FusedSort A;
            A.keys[0] = keys; A.keys[1] = keys_alt; A.vals[0] = vals; A.vals[1] = vals_alt;
            A.counts = ctx->hist.p; A.groups = ctx->hist.p + (size_t)256 * ftiles;
            A.n = (uint32_t)n; A.tiles = ftiles; A.begin_bit = begin_bit; A.end_bit = end_bit;
            if (ftiles == small) MRT_TRY(launch_fused_sort<256>(ctx, A, ftiles));
            else MRT_TRY(launch_fused_sort<1024>(ctx, A, ftiles));
            *result_in_alt = (((end_bit - begin_bit) + 7) / 8) % 2 == 1;
            return mrt_check_cuda(ctx, cudaGetLastError(), "radix_sort_pairs_u64 (fused)");
        }
    }
    MRT_TRY(dev_reserve(ctx, ctx->hist, (size_t)256 * tiles));
    bool in_alt = false;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const uint64_t* kin = in_alt ? keys_alt : keys;
        const uint32_t* vin = in_alt ? vals_alt : vals;
        uint64_t* kout = in_alt ? keys : keys_alt;
        uint32_t* vout = in_alt ? vals : vals_alt;
        k_rs_hist<<<tiles, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, ctx->hist.p, tiles);
        MRT_LAUNCHED(ctx);
        // the tile layout of k_rs_hist must match k_rs_scatter's only per tile (counts), which it does
        MRT_TRY(scan_exclusive_u32(ctx, ctx->hist.p, ctx->hist.p, (size_t)256 * tiles));
        k_rs_scatter<<<tiles, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, ctx->hist.p, tiles);
        MRT_LAUNCHED(ctx);
        in_alt = !in_alt;
    }
    *result_in_alt = in_alt;
    return mrt_check_cuda(ctx, cudaGetLastError(), "radix_sort_pairs_u64");
}
