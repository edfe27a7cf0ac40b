// context.cuh -- the mrt_context: device memory, stream, scene and frame state.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/minotert.h"
#include "sky.cuh"
#include "vec.cuh"

// ---- wide BVH (row n3): 80-byte compressed 8-wide node, five 128-bit words ----
//  w0: grid origin x, y, z (fp32 bits), ex | ey<<8 | ez<<16 | imask<<24   (grid step on axis a = 2^(e_a-127))
//  w1: child_base, tri_base, leafmask24, 0
//  w2: qlo_x[0..7], qlo_y[0..7]        w3: qlo_z[0..7], qhi_x[0..7]       w4: qhi_y[0..7], qhi_z[0..7]
//  child plane = origin + q * step.  imask bit s: slot s is an inner node, the k-th set bit of imask is node
//  child_base + k.  leafmask24 bits 3s..3s+2: triangles of leaf slot s, the k-th set bit of leafmask24 is
//  triangle tri_base + k.  Empty slot: qlo = 255, qhi = 0.
struct WideNode { uint4 w[5]; };
static_assert(sizeof(WideNode) == 80, "wide node is 80 bytes");

#ifndef MRT_MAX_LEAF_TRIS
#define MRT_MAX_LEAF_TRIS 2  // triangles per leaf slot (the node format allows 3): 2 measured best of 1, 2, 3
#endif
#define MRT_MAX_SPHERES 16
#define MRT_MAX_BANDS 8

struct BvhDev {
    const WideNode* nodes;  // [num_nodes]
    const float4* tris;     // [3 * num_leaf_tris] : v0.xyz|prim id, v1.xyz|0, v2.xyz|0 in node-leaf order
    uint32_t num_nodes;
    uint32_t num_tris;
    uint32_t prmt_k;        // 0x47000000: exponent bytes of the float 2^15 + q built by PRMT in the node step
};

struct Partition { uint32_t rank, nranks, slab_rows; };

// local row -> row of the full image (mrt_set_partition)
MRT_HD uint32_t partition_local_to_y(const Partition& p, uint32_t lr) {
    uint32_t slab = lr / p.slab_rows, r = lr % p.slab_rows;
    return (slab * p.nranks + p.rank) * p.slab_rows + r;
}
static inline uint32_t partition_local_rows(const Partition& p, uint32_t H) {
    uint32_t n = 0;
    for (uint32_t y = 0; y < H; y++)
        if ((y / p.slab_rows) % p.nranks == p.rank) n++;
    return n;
}

struct Spheres { mrt_sphere s[MRT_MAX_SPHERES]; uint32_t n; };

template <typename T>
struct DevArray {
    T* p = nullptr;
    size_t cap = 0;  // elements
};

struct mrt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    bool poisoned = false;

    // options
    int opt_count_visits = 0;
    int opt_sort_rays = 0;
    int opt_persistent_primary = 0;  // run primary rays through the persistent state machine too (A/B switch)
    int opt_shade_tiles = 0;         // (A/B: no gain, off) vertex 0 shaded in 8x4-pixel tile order: the bounce queues hold compact patches, not row strips
    int opt_ray_split = 0;           // bounce queues filled from both ends: rays with n.d < value/100 (expected long) are traced first (0: off)
    int opt_trace_carveout = -1;     // cudaFuncAttributePreferredSharedMemoryCarveout of the traversal kernels (-1: driver default)
    int opt_shadow_coherent = 0;     // MRT_SECONDARY_NEE_SUN: shadow rays through the per-lane loop of the primary pass (A/B)
    int opt_primary_batched = 0;     // primary rays: warp-voted triangle steps (trace_coherent_batched) instead of the per-lane loop
    int opt_trace_timing = 1;        // CUDA event pair around every bounce-wave traversal launch (mrt_stats.ms_trace)
    int opt_primary_entry = 0;       // primary pass: walk the top of the BVH once per 32x16 tile against its frustum (entry list), then per ray; measured slower
    int opt_bands = 1;               // wavefront: pixel bands rendered on their own streams so that one band's traversal drain overlaps another band's kernels
    cudaStream_t band_stream[8] = {nullptr};
    cudaEvent_t band_done[8] = {nullptr}, band_fork = nullptr;
    int opt_path_kernel = 0;         // the secondary pass of a triangle scene as ONE persistent launch (mesh.cu k_path); 0: wavefront (trace + shade launches per bounce wave)
    bool secondary_was_path_kernel = false;
    int opt_fused_shade = 0;         // shade stage of a bounce wave inside the traversal kernel (mesh.cu k_trace_shade); A/B: +1.5 % at 1080p 1 spp, -2 % at 4K 8 spp
    int opt_trace_ctas_per_sm = 0;   // 0: as many as fit; n: persistent traversal grids use n CTAs per SM (co-running contexts)
    int opt_builder = 1;             // 0: Karras LBVH hierarchy, 1: PLOC (locally-ordered clustering) hierarchy
    int opt_prepared_rays = 1;       // the shade stage stores 1/d, the shear constants and the octant with the rays it queues (0: the traversal kernel derives them at refill)
    int opt_spheres_batched = 1;     // sphere path: 0 the shader's nested loops, 1 warp-synchronous state machine over samples and bounces (default), 2 ... with per-lane pixel refill (persistent grid; measured slower)
    int spheres_grid = 0;            // resident CTAs of k_spheres_secondary_persistent (queried once)
    int opt_fused_sort = 1;          // radix sort: all passes in one cooperative launch when the tiles are co-resident (0: 5 launches per pass)
    int fused_sort_capacity = -1;    // co-resident CTAs of k_rs_fused (queried once)
    int opt_async_update = 0;        // mrt_scene_update_positions + MRT_BUILD_REFIT return without waiting for the GPU (see minotert.h)
    int opt_wide_refit = 1;          // boxes, planes and leaf triangles level by level on the wide tree, 8 lanes per node (0: round 1's binary climb + one thread per node)
    int opt_build_device_loop = 1;   // PLOC rounds and collapse levels looped inside cooperative kernels (0: host-driven loops with a readback per round)
    int opt_ploc_radius = 6;         // +-positions searched for the nearest cluster (measured best of 2..32 on config 2)

    // inputs
    uchar4* bn = nullptr;
    uint32_t bnW = 0, bnH = 0;
    int scene_kind = 0;  // 0 none, 1 spheres, 2 mesh
    Spheres spheres{};

    // mesh (upload order)
    DevArray<float> pos;      // 3 * nverts
    DevArray<uint32_t> idx;   // 3 * ntris
    DevArray<float4> albedo;  // per primitive, rgb_
    uint32_t nverts = 0, ntris = 0;
    bool bvh_valid = false;
    bool scene_borrowed = false;  // pos/idx/albedo/nodes/tris alias another context's arrays (mrt_scene_share): never freed here
    mrt_context* scene_owner = nullptr;      // whose arrays those are
    std::vector<mrt_context*> borrowers;     // contexts that alias THIS context's scene: made stale before it changes

    // LBVH build scratch + result (bvh_build.cu)
    DevArray<float4> prim_lo, prim_hi;      // per primitive AABB
    DevArray<uint64_t> keys, keys_alt;      // Morton keys
    DevArray<uint32_t> order, order_alt;    // sorted primitive ids
    DevArray<uint32_t> hist;                // radix histograms
    DevArray<uint32_t> scan_tmp;
    DevArray<int32_t> bin_left, bin_right, bin_parent;  // binary nodes: [0,N-1) internal, [N-1,2N-1) leaves
    DevArray<uint32_t> bin_count;                        // primitives below each internal node
    int bin_root = 0;                                    // root of the binary tree (0: LBVH, n-2: PLOC)
    DevArray<uint32_t> ploc_c[2], ploc_nn, ploc_flag[2], ploc_scan[2];  // PLOC cluster lists and scratch
    DevArray<float4> bin_lo, bin_hi;                      // boxes of all 2N-1 binary nodes
    DevArray<uint32_t> bin_flag;
    DevArray<uint4> bin_rec;                              // child records of the internal binary nodes (k_collapse_loop)
    DevArray<uint32_t> scene_bounds;                      // 6 ordered-int floats
    DevArray<uint2> work_a, work_b;                       // collapse work items (binary node, wide node)
    DevArray<int32_t> slot_node;                          // [num_nodes][8] binary node behind each slot
    DevArray<uint32_t> node_nchild, node_ntri, node_child_base, node_tri_base;
    DevArray<WideNode> nodes;
    DevArray<float4> tris;
    // option async_update: vertex uploads on their own stream, ordered against refits and the borrowers' frames by events
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t ev_pos_ready = nullptr, ev_refit_done = nullptr, ev_frames_done = nullptr;
    bool pos_upload_pending = false, refit_recorded = false, build_time_pending = false, frames_marked = false;
    DevArray<WideNode> nodes_alt;                         // second copy of the tree: an asynchronous refit writes the copy no
    DevArray<float4> tris_alt;                            // frame in flight reads, then the copies swap roles
    bool alt_valid = false;                               // nodes_alt / tris_alt hold the current topology
    DevArray<float4> node_lo, node_hi;                    // boxes of the wide nodes (scratch of the level-wise emission / refit)
    DevArray<uint32_t> level_starts_dev;                  // first wide node of each level (+ the node count), written by k_collapse_loop
    uint32_t* build_results_host = nullptr;               // page-locked landing area of the build's result words + level table
    std::vector<uint32_t> level_starts;                   // host copy: level L = nodes [level_starts[L], level_starts[L + 1])
    DevArray<uint2> loop_sums;                            // per-CTA counts of the device-side build loops
    DevArray<uint32_t> counters;                          // misc device counters
    uint32_t num_nodes = 0, num_leaf_tris = 0;

    // sky
    mrt_atmosphere_params atmo{};
    bool have_atmo = false, have_view = false;
    DevArray<uint16_t> trans16, multi16;
    DevArray<uint32_t> view_packed;
    DevArray<float4> trans_f, multi_f, view_f;
    DevArray<uint16_t> aerial16;   // aerial-perspective camera volume, 32^3 RGBA16F (mrt_sky_aerial_perspective)
    DevArray<float4> aerial_f;     // ... decoded for the trilinear taps
    bool have_aerial = false;

    // frame
    Partition part{0, 1, 8};
    uint32_t W = 0, H = 0, local_rows = 0;
    size_t npix = 0;  // local pixels
    bool have_gbuffer = false, have_color = false, have_accum = false, have_ldr = false;
    bool secondary_done = false;   // a secondary pass has run (its events and queue counters can be read)
    bool have_frame_sum = false;   // a frame rendered with MRT_SECONDARY_FRAME_SUM waits in frame_sum (mrt_accum_commit)
    mrt_primary_constants pc{};
    DevArray<uint32_t> visibility;
    DevArray<uint16_t> depth, normal, motion, color16;
    DevArray<float> hit_t;
    DevArray<float4> accum;
    DevArray<float4> sun_e;        // MRT_SECONDARY_NEE_SUN: sun-centre radiance at the camera, one value per frame
    DevArray<float4> shadow_q[3];  // MRT_SECONDARY_NEE_SUN: shadow-ray queue (origin|pixel, direction, contribution)
    uint32_t back_counts_at = 0, num_back_counts = 0;      // option ray_split: sizes of the queues' back ends
    uint32_t shadow_counts_at = 0, num_shadow_counts = 0;  // where the shadow-queue sizes sit in queue_counts
    DevArray<float4> frame_sum;    // MRT_SECONDARY_FRAME_SUM: this frame's radiance sums (xyz) and samples (w)
    cudaEvent_t commit_ev[2] = {nullptr, nullptr};  // mrt_accum_commit: src rendered / dst consumed
    // bilateral denoiser (denoise.cu): RGBA8 output, tap list cached per (sigma, kSigma, image size)
    DevArray<uchar4> denoised;
    DevArray<float4> dn_taps;
    int dn_ntaps = 0, dn_ncols = 0;
    float dn_key_sigma = 0.0f, dn_key_ksigma = 0.0f;
    uint32_t dn_key_w = 0, dn_key_h = 0;
    bool have_denoised = false;
    // temporal reprojection (temporal.cu): double-buffered history = previous output (rgb, 1), its history length and
    // the previous frame's visibility ids
    DevArray<float4> tp_rgba[2];
    DevArray<float> tp_count[2];
    DevArray<uint32_t> tp_vis[2];
    int tp_cur = 0;
    uint32_t tp_w = 0, tp_h = 0;
    bool have_temporal = false;
    DevArray<uchar4> ldr_buf[2];         // double-buffered output framebuffer: an async readback of frame f
    int ldr_cur = 0;                     // overlaps the rendering of frame f+1 (the reference keeps 3 frames in flight)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ldr_ready = nullptr, copy_done[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    // wavefront state (triangle path)
    DevArray<float4> hit0_pos, hit0_n;   // primary hit position|prim id, normal|valid
    DevArray<float4> path_state;         // throughput rgb | rng state
    DevArray<float4> ray_o[2], ray_d[2]; // queues: origin|pixel, direction|tmax
    DevArray<float4> ray_p[2], ray_s[2]; // option prepared_rays: (1/d, Sx), (Sy, Sz, axes | octant) of the queued rays
    DevArray<unsigned long long> hits;   // t bits | tri slot << 32 (MRT_HIT_PENDING_TRI: not traced yet)
    bool hits_dirty = true;              // records are not all "pending": reset before the next fused wave
    DevArray<uint32_t> queue_counts;     // one counter per wave, then one work counter per trace launch
    uint32_t num_queue_counts = 0;
    DevArray<uint64_t> sort_keys, sort_keys_alt;
    DevArray<uint32_t> sort_vals, sort_vals_alt;
    DevArray<float> query_o, query_d, query_t;    // mrt_trace_rays scratch
    DevArray<uint32_t> query_ids;
    DevArray<unsigned long long> visit_counters;  // node visits, tri tests, stack overflows
    DevArray<unsigned long long> total_rays;      // running sum of traced rays since mrt_stats_reset

    // stats
    mrt_stats stats{};
    cudaEvent_t ev[14] = {nullptr};  // pairs: sky, primary, secondary, tonemap, denoise, temporal, BVH build/refit
    // The sky view of a frame is generated on a side stream so that it overlaps the primary pass
    // (Renderer::draw order: sky -> primary -> secondary); consumers join through sky_join().
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_fork = nullptr, sky_ready = nullptr;
    bool sky_pending = false;
    std::vector<cudaEvent_t> trace_ev;  // begin/end pairs around each traversal launch of the last frame
    uint32_t trace_ev_used = 0;
};

// ---- error plumbing ----
int mrt_fail(mrt_context* ctx, int code, const char* fmt, ...);
int mrt_check_cuda(mrt_context* ctx, cudaError_t e, const char* what);
#define MRT_CUDA(ctx, call)                                            \
    do {                                                               \
        int _s = mrt_check_cuda((ctx), (call), #call);                 \
        if (_s != MRT_OK) return _s;                                   \
    } while (0)
#define MRT_TRY(call)                    \
    do {                                 \
        int _s = (call);                 \
        if (_s != MRT_OK) return _s;     \
    } while (0)

template <typename T>
static inline int dev_reserve(mrt_context* ctx, DevArray<T>& a, size_t n) {
    if (n <= a.cap && a.p) return MRT_OK;
    if (a.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(a.p);
        a.p = nullptr;
        a.cap = 0;
    }
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)&a.p, n * sizeof(T));
    if (e != cudaSuccess) {
        a.p = nullptr;
        return mrt_fail(ctx, e == cudaErrorMemoryAllocation ? MRT_ERR_OOM : MRT_ERR_CUDA, "cudaMalloc(%zu bytes): %s",
                        n * sizeof(T), cudaGetErrorString(e));
    }
    a.cap = n;
    return MRT_OK;
}
template <typename T>
static inline void dev_free(DevArray<T>& a) {
    if (a.p) cudaFree(a.p);
    a.p = nullptr;
    a.cap = 0;
}

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
#define MRT_LAUNCHED(ctx) ((ctx)->stats.kernel_launches++)

// ---- stage entry points implemented across the .cu files ----
int sky_gen_atmosphere(mrt_context* ctx);
int sky_gen_view(mrt_context* ctx, const float probe[3], const float sunDir[3], const float sunIll[3]);
int spheres_primary(mrt_context* ctx);
int spheres_secondary(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp, uint32_t bounces, uint32_t flags);
int denoise_bilateral(mrt_context* ctx, float sigma, float kSigma, float threshold, float nearPlane, uint32_t frameCounter);
int temporal_accumulate(mrt_context* ctx, float maxHistory, bool reset);
int sky_gen_aerial(mrt_context* ctx, const mrt_primary_constants* c, const float cameraPos[3], const float sunDir[3], const float sunIll[3]);
int tonemap_run(mrt_context* ctx, int mode, float exposure, const float* params, uint32_t nparams, int source);
int bvh_build_full(mrt_context* ctx);
int bvh_refit(mrt_context* ctx, bool wait = true, cudaStream_t side = nullptr);  // wait = false: queued on `side` only (mrt_stats_get reads ms_build later)
bool bvh_refit_can_be_async(const mrt_context* ctx);
int mesh_primary(mrt_context* ctx);
int mesh_secondary(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp, uint32_t bounces, uint32_t flags);
int probe_sky_color(mrt_context* ctx, const float cameraPos[3], const float* dirs, uint32_t n, float* out);
int probe_bounce_stream(mrt_context* ctx, uint32_t frameCounter, uint32_t x, uint32_t y, const float pos[3], const float normal[3],
                        uint32_t n, float* out9);
int mesh_trace_rays(mrt_context* ctx, const float* o, const float* d, uint32_t n, uint32_t* ids, float* t, int brute);
// device-wide primitives (sort.cu)
int scan_exclusive_u32(mrt_context* ctx, const uint32_t* in, uint32_t* out, size_t n);
int radix_sort_pairs_u64(mrt_context* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t n,
                         int begin_bit, int end_bit, bool* result_in_alt);
