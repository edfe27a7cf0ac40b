// sort.cu -- device-wide exclusive scan and LSD radix sort (64-bit keys, 32-bit payload).
// Used by the LBVH builder (Morton keys, row n2) and by the optional ray sort between bounces
// (row n5).  Warp-level primitives do the ranking: __match_any_sync groups lanes with the same
// digit, popc of the lower-lane mask gives a stable intra-warp rank, per-warp digit counters in
// shared memory are then prefix-summed across the 8 warps of a tile.  HBM-streaming kernels.
#include "context.cuh"

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
    __shared__ uint32_t gbase[256];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
    gbase[threadIdx.x] = offsets[(size_t)threadIdx.x * num_tiles + blockIdx.x];
    __syncthreads();

    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], loc[RS_ITEMS];
    const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * RS_WARP_SPAN;
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
    {
        // exclusive prefix over the tile's warps for digit = threadIdx.x
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = wc[w][threadIdx.x];
            wc[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        if (i < n) {
            unsigned d = (unsigned)(key[j] >> shift) & 255u;
            size_t pos = (size_t)gbase[d] + wc[warp][d] + loc[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
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
