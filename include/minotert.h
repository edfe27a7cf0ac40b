/*
 * minotert.h -- C ABI of the B200-native MinoteRT ray-tracing hot path (libminotert.so).
 *
 * This is the drop-in boundary: the entry points a maintainer of Tearnote/MinoteRT would bind
 * in place of the vuk/Vulkan dispatches recorded by src/gfx/modules/{pathtracer,sky,tonemapper}.ixx.
 * Plain pointers and sizes only; no exceptions cross the seam.  Every function returns
 * MRT_OK (0) or a negative mrt_status; mrt_last_error() gives the sticky message.
 *
 * Threading: one host thread per context (the reference is single-threaded, src/main.cpp:20);
 * distinct contexts may be driven from distinct threads.  All work is stream-ordered on the
 * context's CUDA stream; only mrt_sync / mrt_readback / mrt_stats block.
 *
 * Ownership: the context owns all device memory.  Host pointers passed in are consumed before
 * the call returns.  Device pointers from mrt_buffer are BORROWED: valid until the next call
 * that (re)renders that buffer at a different size, or mrt_destroy.
 *
 * There is no CPU fallback: without a CUDA device mrt_create fails with MRT_ERR_CUDA.
 */
#ifndef MINOTERT_H
#define MINOTERT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRT_ABI_VERSION 1
#define MRT_MISS_ID 0xFFFFFFFFu /* primaryRay.comp:25 `-1u` */

typedef enum {
    MRT_OK = 0,
    MRT_ERR_INVALID = -1, /* bad argument / call order */
    MRT_ERR_CUDA = -2,    /* CUDA runtime error (context is poisoned) */
    MRT_ERR_OOM = -3,
    MRT_ERR_STATE = -4    /* required input (scene, LUTs, blue noise, G-buffer) missing */
} mrt_status;

typedef struct mrt_context mrt_context;

/* ---- PODs: the binary host<->device contract of the reference ---- */

/* column-major, m[col][row]  (src/stx/math.ixx:531-539) */
typedef struct { float m[4][4]; } mrt_mat4;

/* UBO of primaryRay.comp:14-21, filled at src/gfx/modules/pathtracer.ixx:94-104 */
typedef struct {
    mrt_mat4 view, projection, invView, invProjection, prevView;
    uint32_t frameCounter;
} mrt_primary_constants;

/* UBO of secondaryRays.comp:25-32, filled at src/gfx/modules/pathtracer.ixx:178-188 */
typedef struct {
    mrt_mat4 view, projection, invView, invProjection;
    float cameraPos[3];
    uint32_t frameCounter;
} mrt_secondary_constants;

/* src/gpu/intersect.glsl:9-13 */
typedef struct { float center[3]; float radius; float albedo[3]; } mrt_sphere;

/* Atmosphere::Params, std140 mirror, 144 bytes (src/gfx/modules/sky.ixx:28-56) */
typedef struct {
    float bottomRadius, topRadius, rayleighDensityExpScale, _pad0;
    float rayleighScattering[3]; float mieDensityExpScale;
    float mieScattering[3]; float _pad1;
    float mieExtinction[3]; float _pad2;
    float mieAbsorption[3]; float miePhaseG;
    float absorptionDensity0LayerWidth, absorptionDensity0ConstantTerm,
          absorptionDensity0LinearTerm, absorptionDensity1ConstantTerm;
    float absorptionDensity1LinearTerm, _pad3, _pad4, _pad5;
    float absorptionExtinction[3]; float _pad6;
    float groundAlbedo[3]; float _pad7;
} mrt_atmosphere_params;

/* Tonemapper::<op> selector, same order as Renderer_impl::tonemap (src/gfx/renderer.ixx:165-172) */
typedef enum {
    MRT_TONEMAP_LINEAR = 0, MRT_TONEMAP_REINHARD = 1, MRT_TONEMAP_HABLE = 2,
    MRT_TONEMAP_ACES = 3, MRT_TONEMAP_UCHIMURA = 4, MRT_TONEMAP_AMD = 5
} mrt_tonemap_mode;

typedef enum {
    MRT_BUF_VISIBILITY = 0, /* R32_UINT   4 B/px  (pathtracer.ixx:42-48)  */
    MRT_BUF_DEPTH = 1,      /* R16F       2 B/px  (pathtracer.ixx:49-55)  */
    MRT_BUF_NORMAL = 2,     /* RGBA16F    8 B/px  (pathtracer.ixx:56-62)  */
    MRT_BUF_MOTION = 3,     /* RG16F      4 B/px  (pathtracer.ixx:63-69)  */
    MRT_BUF_COLOR = 4,      /* RGBA16F    8 B/px  (pathtracer.ixx:139-145): frame average   */
    MRT_BUF_ACCUM = 5,      /* RGBA32F   16 B/px  progressive sum, .w = samples (row n7)    */
    MRT_BUF_LDR = 6,        /* RGBA8      4 B/px  (tonemapper.ixx:332) output framebuffer   */
    MRT_BUF_TRANSMITTANCE = 7,   /* RGBA16F 256x64 (sky.ixx:22-23)   */
    MRT_BUF_MULTISCATTERING = 8, /* RGBA16F 32x32  (sky.ixx:25-26)   */
    MRT_BUF_SKY_VIEW = 9,        /* B10G11R11 192x108 (sky.ixx:187-188) */
    MRT_BUF_HIT_T = 10,          /* R32F 4 B/px primary hit distance (triangle scenes) */
    MRT_BUF_DENOISED = 11,       /* RGBA8 4 B/px (denoiser.ixx:56) output of mrt_denoise_bilateral */
    MRT_BUF_BVH_NODES = 12,      /* the built wide BVH, for tests and tools: 80-byte nodes ... */
    MRT_BUF_BVH_TRIS = 13,       /* ... and its triangles in leaf order, 3 x float4 each (v0.w = primitive id) */
    MRT_BUF_TEMPORAL = 14,       /* RGBA32F 16 B/px (rgb, 1): output of mrt_temporal_accumulate = next frame's history */
    MRT_BUF_TEMPORAL_COUNT = 15, /* R32F 4 B/px: history length of each pixel after mrt_temporal_accumulate (1 = reset) */
    MRT_BUF_AERIAL = 16,         /* RGBA16F 32 x 32 x 32 (x fastest): aerial-perspective volume of mrt_sky_aerial_perspective */
    MRT_BUF_COUNT_
} mrt_buffer_id;

typedef enum {
    MRT_BUILD_FULL = 0, /* Morton sort + LBVH + wide-node collapse */
    MRT_BUILD_REFIT = 1 /* keep topology, recompute boxes after mrt_scene_update_positions */
} mrt_build_mode;

/* mrt_secondary_rays flags */
#define MRT_SECONDARY_ACCUMULATE 1u /* add to MRT_BUF_ACCUM instead of restarting it */
#define MRT_SECONDARY_SORT_RAYS 2u  /* reorder each bounce's ray queue before it is traced (image unchanged) with the
                                     * mode of option "sort_rays" (octant binning if that is 0) */
#define MRT_SECONDARY_NEE_SUN 8u    /* triangle scenes: the sun is a sampled light -- every hit vertex that
                                     * bounces sends a shadow ray into the sun's disc (two more rotated random numbers,
                                     * drawn before the bounce's; any-hit traversal) and adds throughput * E_sun * n.l *
                                     * limb * Omega / pi when it is free; escaping bounce rays then add the sky-view term
                                     * only.  Same expectation as the reference's estimator, without waiting for a bounce
                                     * ray to hit a 0.5 degree disc.  Shadow rays count as secondary rays. */
#define MRT_SECONDARY_SKY_AT_HIT 16u /* the sky is evaluated at the origin of the escaping ray (the shaded point), not at
                                     * the camera position as secondaryRays.comp:37 does for every vertex */
#define MRT_SECONDARY_AERIAL 32u    /* aerial perspective between the camera and the primary hit from the volume of
                                     * mrt_sky_aerial_perspective: each sample's throughput starts at 1 - AP.a and AP.rgb
                                     * is added once per sample (pixels whose primary ray escapes are left alone) */
#define MRT_SECONDARY_FRAME_SUM 4u  /* triangle scenes: this call's samples are summed into the context's own per-frame
                                     * buffer, starting from zero, and MRT_BUF_ACCUM is left alone; mrt_accum_commit
                                     * then adds the frame to an accumulator.  Frames rendered by different contexts
                                     * (frames in flight) and committed in frame order give the same bits as one
                                     * context rendering them one after the other in this mode. */

typedef struct {
    uint64_t primary_rays;   /* rays traced by the last mrt_primary_rays */
    uint64_t secondary_rays; /* rays traced by the last mrt_secondary_rays */
    float ms_primary;        /* device time of the last call of each kind (CUDA events) */
    float ms_secondary;
    float ms_trace;          /* bounce-wave traversal launches since mrt_stats_reset: summed device time */
    float ms_tonemap;
    float ms_build;
    float ms_sky;
    uint32_t kernel_launches; /* kernels launched by this context since mrt_stats_reset */
    uint32_t stack_overflows; /* traversal stack overflow events (must be 0) */
    uint32_t num_triangles;
    uint32_t num_wide_nodes;
    uint64_t bvh_bytes;       /* wide nodes + reordered triangles resident in HBM */
    uint64_t node_visits;     /* wide nodes fetched / triangles tested by the last counted trace */
    uint64_t tri_tests;       /*   (only when mrt_set_option("count_visits", 1))            */
    uint32_t trace_launches;  /* ... and their number (CUDA event pairs, at most 4096 per reset) */
    uint32_t _reserved;
    uint64_t total_rays;      /* primary + secondary rays traced since mrt_stats_reset (device-side running sum) */
    float sah_node_cost;      /* surface-area heuristic of the wide BVH: expected node steps ... */
    float sah_tri_cost;       /* ... and triangle tests of a random ray that hits the root box */
    float ms_denoise;         /* device time of the last mrt_denoise_bilateral */
    float ms_temporal;        /* device time of the last mrt_temporal_accumulate */
    uint64_t secondary_node_visits; /* the share of node_visits / tri_tests spent by the last mrt_secondary_rays   */
    uint64_t secondary_tri_tests;   /*   (bounce rays only: what the roofline of the bounce-wave kernel is quoted on) */
} mrt_stats;

/* ---- lifetime ---- */
int mrt_abi_version(void);
/* device: CUDA ordinal.  Replaces Vulkan::Provider + lazy pipeline creation
 * (src/sys/vulkan.ixx:106-115, pathtracer.ixx:30-39). */
int mrt_create(int device, mrt_context** out);
void mrt_destroy(mrt_context* ctx);
/* Sticky message of the last failure on this context (ctx may be NULL: last mrt_create error). */
const char* mrt_last_error(const mrt_context* ctx);
/* Tuning / instrumentation switches (unknown names fail with MRT_ERR_INVALID):
 *   "count_visits" 0/1        count node visits and triangle tests (mrt_stats.node_visits / tri_tests)
 *   "trace_timing" 0/1        CUDA event pair around every bounce-wave traversal launch (default 1)
 *   "sort_rays" 0/1/2         reorder every bounce wave's queue before it is traced: 1 stable binning by direction
 *                             octant, 2 radix sort by (Morton cell of the ray origin in a 128^3 grid, direction octant);
 *                             default 0 (both measured slower than the pixel order compaction leaves).  Same image.
 *   "primary_entry" 0/1       primary pass: the top of the BVH is walked once per 32x16-pixel tile against the tile's
 *                             frustum and every ray starts from the resulting entry list (1); default 0: every ray
 *                             walks from the root (measured faster).  Same G-buffer bit for bit.
 *   "bands" 1..8              wavefront: the image's pixels are cut into that many ranges, each rendered on its own
 *                             stream, so that the drain of one band's traversal launch overlaps another band's
 *                             kernels.  Same image bit for bit.
 *   "path_kernel" 0/1         triangle scenes: the whole secondary pass (all samples and bounces) as ONE persistent
 *                             launch in which a lane owns a pixel (1), instead of the wavefront of trace + shade
 *                             launches per bounce wave with compacted ray queues (0, default: measured faster).
 *                             Same image bit for bit.
 *   "fused_shade" 0/1         shade stage of a bounce wave inside its traversal launch (default 0; measured +-2 %)
 *   "persistent_primary" 0/1  ... and for primary rays (default 0: coherent per-lane loop)
 *   "trace_ctas_per_sm" n     cap the persistent traversal grid at n CTAs per SM (0 = as many as fit), for
 *                             contexts that share a GPU
 *   "build_device_loop" 0/1   PLOC rounds and collapse levels looped inside two cooperative kernels (default 1) or
 *                             driven from the host with a readback per round (0); same tree either way; invalidates the BVH
 *   "prepared_rays" 0/1       bounce rays are queued with 1/d, the shear constants of the triangle test and the octant, computed
 *                             by the shade stage (default 1), or the traversal kernel derives them when a lane takes
 *                             the ray (0); same image bit for bit
 *   "spheres_batched" 0/1/2   sphere path: the shader's nested loops (0); every lane runs its samples and bounces back to back
 *                             through a warp-synchronous state machine and escaped paths wait for a batched sky
 *                             evaluation (1, default); ... and a lane that has finished its pixel takes the next one,
 *                             persistent grid (2: measured slower); same image and ray count bit for bit
 *   "fused_sort" 0/1          Morton / ray radix sort: all 8-bit passes in one cooperative launch when every tile's CTA is
 *                             resident at once (<= 1.2 M keys; default 1), or 5 launches per pass (0); same permutation;
 *                             invalidates the BVH
 *   "wide_refit" 0/1          final node emission and MRT_BUILD_REFIT level by level on the wide tree, 8 lanes per node
 *                             (default 1), or through the binary tree's boxes with one thread per node (0); same nodes
 *                             and leaf triangles bit for bit
 *   "async_update" 0/1        animated scenes without host stalls (default 0).  With 1, mrt_scene_update_positions queues the
 *                             copy on an upload stream of its own and returns at once -- `positions` must be page-locked
 *                             and stay untouched until mrt_sync, a readback or mrt_stats_get of this context returns --
 *                             and mrt_scene_build(MRT_BUILD_REFIT) queues the refit without waiting for it: frames in
 *                             flight of contexts that borrow the scene (mrt_scene_share) are waited for on the GPU, the
 *                             borrowers stay valid (a refit rewrites nodes and triangles in place) and the frames they
 *                             record next wait for the refit.  MRT_BUILD_FULL still returns when the build is done, but
 *                             runs beside the frames in flight (into the second copy of the tree) instead of draining them.
 *   "builder" 0/1             hierarchy builder: 0 Karras LBVH, 1 PLOC (default); invalidates the BVH
 *   "ploc_radius" 1..32       PLOC search radius (default 6); invalidates the BVH */
int mrt_set_option(mrt_context* ctx, const char* name, int64_t value);

/* ---- inputs ---- */
/* RGBA8 blue-noise texture (Renderer_impl ctor, src/gfx/renderer.ixx:101-110) */
int mrt_upload_blue_noise(mrt_context* ctx, const uint8_t* rgba8, uint32_t w, uint32_t h);
/* The scene of src/gpu/scene.glsl:4-11 as data (n <= 16).  Selects the sphere path. */
int mrt_scene_set_spheres(mrt_context* ctx, const mrt_sphere* spheres, uint32_t n);
/* Indexed triangle mesh, primitive id = triangle index in upload order; albedo: 3 floats per
 * triangle.  Selects the triangle path (north_star rows n1-n7).  Call mrt_scene_build next. */
int mrt_scene_upload_mesh(mrt_context* ctx, const float* positions, uint32_t nverts,
                          const uint32_t* indices, uint32_t ntris, const float* albedo);
/* New vertex positions for the uploaded topology (animated scenes); then build FULL or REFIT. */
int mrt_scene_update_positions(mrt_context* ctx, const float* positions, uint32_t nverts);
int mrt_scene_build(mrt_context* ctx, int build_mode);
/* Frames in flight.  The reference records each frame into its own arena with 3 frames in flight
 * (src/gfx/renderer.ixx:36,42-43,97) while the scene resources are shared; here one context = one frame's
 * buffers + stream, and ctx renders the mesh scene (triangles, albedo, built BVH) that `owner` holds instead
 * of its own copy.  Borrowed, not copied.  Whenever the owner changes the scene (mrt_scene_upload_mesh,
 * mrt_scene_update_positions, mrt_scene_build) the library first waits for the borrowers' frames in flight and marks
 * them stale: their render calls fail with MRT_ERR_STATE until mrt_scene_share is called again (cheap; it also orders
 * ctx's stream behind the owner's queued build).  Destroying the owner leaves its borrowers without a scene.
 * mrt_scene_update_positions / mrt_scene_build on a borrowing context fail with MRT_ERR_STATE;
 * mrt_scene_upload_mesh ends the borrowing.  Owner and borrowers are driven from one host thread. */
int mrt_scene_share(mrt_context* ctx, mrt_context* owner);

/* ---- sky (src/gfx/modules/sky.ixx) ---- */
/* Atmosphere(allocator, params): transmittance + multi-scattering LUTs (sky.ixx:91-176) */
int mrt_atmosphere(mrt_context* ctx, const mrt_atmosphere_params* params);
/* Sky::createView(atmo, probePos) with the push constants of sky.ixx:239-250 */
int mrt_sky_view(mrt_context* ctx, const float probePos[3], const float sunDirection[3],
                 const float sunIlluminance[3]);
/* Aerial-perspective camera volume (the reference declares it -- Sky::AerialPerspectiveFormat / Size, sky.ixx:190-191,
 * AP_KM_PER_SLICE and the depth<->slice maps, skyAccess.glsl:9,119-125 -- and never builds it).  32^3
 * froxels over the view of the camera whose primary constants are given: luminance scattered towards the camera and
 * 1 - transmittance up to depth ((z + 0.5) / 32)^2 * 128 km, RGBA16F.  Needs mrt_atmosphere.  Applied by
 * mrt_secondary_rays(MRT_SECONDARY_AERIAL); readable as MRT_BUF_AERIAL. */
int mrt_sky_aerial_perspective(mrt_context* ctx, const mrt_primary_constants* c, const float cameraPos[3],
                               const float sunDirection[3], const float sunIlluminance[3]);

/* ---- partition (multi-GPU tile mode; default rank 0 of 1) ----
 * Image rows are grouped into slabs of slab_rows rows; slab j belongs to rank j % nranks.
 * A context renders only its own slabs, stored compactly (slab-major) in its buffers. */
int mrt_set_partition(mrt_context* ctx, uint32_t rank, uint32_t nranks, uint32_t slab_rows);
/* rows of the full image owned by this context, in local storage order */
int mrt_partition_rows(const mrt_context* ctx, uint32_t full_h, uint32_t* rows_out,
                       uint32_t* nrows_out);
/* same mapping without a context (host-side gather/scatter logic, usable without a GPU) */
int mrt_partition_rows_for(uint32_t rank, uint32_t nranks, uint32_t slab_rows, uint32_t full_h,
                           uint32_t* rows_out, uint32_t* nrows_out);

/* ---- per-frame render calls, in the order of Renderer_impl::draw (renderer.ixx:56-62) ---- */
/* Pathtracer::primaryRays(size, camera, prevCamera) -> GBuffer (pathtracer.ixx:29-116) */
int mrt_primary_rays(mrt_context* ctx, uint32_t w, uint32_t h, const mrt_primary_constants* c);
/* Pathtracer::secondaryRays(gbuffer, camera, atmo, skyView, blueNoise) (pathtracer.ixx:118-195).
 * The reference hard-codes spp = 8, bounces = 8 (secondaryRays.comp:128-129). */
int mrt_secondary_rays(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp,
                       uint32_t bounces, uint32_t flags);
/* Denoiser::bilateral(color, depth, normal, camera, params) (src/gfx/modules/denoiser.ixx:36-97,
 * src/gpu/denoise/bilateral.comp): filters MRT_BUF_COLOR guided by MRT_BUF_DEPTH / MRT_BUF_NORMAL into
 * MRT_BUF_DENOISED (RGBA8 unorm, i.e. clamped to [0,1] before the tonemapper, denoiser.ixx:56).
 * sigma, kSigma, threshold: BilateralParams (defaults 5, 2, 0.12, denoiser.ixx:27-33); nearPlane and
 * frameCounter: the other push constants (denoiser.ixx:85-91).  round(kSigma*sigma) <= 32.
 * Needs the whole image in one context (fails with MRT_ERR_STATE under mrt_set_partition with nranks > 1). */
int mrt_denoise_bilateral(mrt_context* ctx, float sigma, float kSigma, float threshold, float nearPlane,
                          uint32_t frameCounter);
/* Temporal accumulation by reprojection.  No reference counterpart: it is the consumer of
 * the motion image that src/gpu/primaryRay.comp:73-75 writes (pathtracer.ixx:63-69) and Renderer_impl::draw drops
 * (renderer.ixx:61).  Blends the frame's average radiance (MRT_BUF_ACCUM) into the history fetched at each pixel's
 * previous position (MRT_BUF_MOTION; history taps on another primitive are rejected through the previous frame's
 * MRT_BUF_VISIBILITY) -> MRT_BUF_TEMPORAL (+ MRT_BUF_TEMPORAL_COUNT), which is the next call's history.
 * maxHistory: cap of the history length (the blend weight of a new frame never falls below 1/(maxHistory+1)).
 * The first call, a call after the image size or the scene changed, or MRT_TEMPORAL_RESET start a new history.
 * Call after mrt_primary_rays (with prevView = the previous frame's view) + mrt_secondary_rays; whole image in
 * one context (MRT_ERR_STATE under a partition).  Exact contract: oracle/minote_oracle.h orc_temporal_accumulate. */
#define MRT_TEMPORAL_RESET 1u
int mrt_temporal_accumulate(mrt_context* ctx, float maxHistory, uint32_t flags);
/* Tonemapper::{linear,reinhard,hable,aces,uchimura,amd}(input, exposure, params)
 * (tonemapper.ixx:57-373).  source: MRT_BUF_COLOR (reference path), MRT_BUF_ACCUM
 * (progressive average), MRT_BUF_DENOISED (the denoiser's RGBA8 image, as Renderer_impl::draw
 * chains them, renderer.ixx:61-62) or MRT_BUF_TEMPORAL.  params: the push constants after `exposure`. */
int mrt_tonemap(mrt_context* ctx, int mode, float exposure, const float* params, uint32_t nparams,
                int source);

/* ---- outputs ---- */
int mrt_buffer(mrt_context* ctx, int buffer_id, void** device_ptr, size_t* bytes);
int mrt_readback(mrt_context* ctx, int buffer_id, void* host, size_t bytes);
/* Pipelined framebuffer readback (MRT_BUF_LDR only): the device->host copy of the frame just tonemapped runs
 * on a second stream while the caller issues the next frame, which tonemaps into the other half of the
 * double-buffered framebuffer -- the counterpart of the reference's frames in flight (renderer.ixx:36).
 * host must stay valid (ideally pinned) until mrt_readback_wait returns. */
int mrt_readback_async(mrt_context* ctx, int buffer_id, void* host, size_t bytes);
/* blocks until at most frames_in_flight (0 or 1) async readbacks are still outstanding */
int mrt_readback_wait(mrt_context* ctx, int frames_in_flight);
int mrt_sync(mrt_context* ctx);
int mrt_stats_get(mrt_context* ctx, mrt_stats* out);
int mrt_stats_reset(mrt_context* ctx);
/* raw CUDA stream (cudaStream_t) of the context, for callers that enqueue collectives */
int mrt_stream(mrt_context* ctx, void** stream_out);

/* ---- closest-hit query (tests / tools): traces n rays through the built scene ---- */
int mrt_trace_rays(mrt_context* ctx, const float* origins, const float* directions, uint32_t n,
                   uint32_t* prim_ids, float* t, int brute_force);

/* ---- checkpoint / resume of a progressive render: the only cross-frame state of the path is the fp32
 * accumulator (xyz radiance sums, w samples).  Dump it with mrt_readback(MRT_BUF_ACCUM); restore it here -- after an
 * mrt_primary_rays at the same size and partition -- and continue with MRT_SECONDARY_ACCUMULATE and the next frame
 * counter: the result is bit-identical to the uninterrupted render. */
int mrt_accum_restore(mrt_context* ctx, const float* rgba32f, size_t bytes);
/* Progressive accumulation with frames in flight: dst.MRT_BUF_ACCUM (+)= the frame src rendered with
 * MRT_SECONDARY_FRAME_SUM (flags & MRT_SECONDARY_ACCUMULATE: add to dst's accumulator, else restart it with this
 * frame).  dst may be src; otherwise both live on one device, dst takes src's image size and partition, the add runs on
 * dst's stream behind src's render, and src's next render waits for it.  After the call dst can be tonemapped
 * (source MRT_BUF_ACCUM) / read back / gathered like a context that rendered the frames itself. */
int mrt_accum_commit(mrt_context* dst, mrt_context* src, uint32_t flags);

/* ---- device-side unit probes (tests / tools): the per-path shading functions evaluated on the GPU ----
 * skyColor() of src/gpu/secondaryRays.comp:36-58 for n directions (directions, rgb_out: n x 3 floats) with the LUTs of
 * the last mrt_atmosphere / mrt_sky_view */
int mrt_eval_sky_color(mrt_context* ctx, const float cameraPos[3], const float* directions, uint32_t n, float* rgb_out);
/* the sample stream of one pixel (secondaryRays.comp:60-72,124-125): PCG state seeded with (frameCounter << 1) | 1,
 * Cranley-Patterson rotation = blue-noise texel of pixel (x, y), then n consecutive Lambert bounces off
 * (position, normal).  out9: n x 9 floats = r0, r1 (the two rotated randoms), ray origin xyz, ray direction xyz,
 * PCG state after the bounce (bit pattern) */
int mrt_eval_bounce_stream(mrt_context* ctx, uint32_t frameCounter, uint32_t x, uint32_t y, const float position[3],
                           const float normal[3], uint32_t n, float* out9);

/* ---- multi-GPU groups (no reference counterpart: the reference is single-GPU) ----
 * A group = nranks contexts, each holding the whole scene (replicated BVH: upload + build through the ordinary
 * calls on every mrt_group_context).  The render never communicates; the one exchange step is of the finished
 * image:  tile mode   mrt_group_set_tiles + mrt_group_render + [mrt_group_tonemap] + mrt_group_gather
 *         sample sets mrt_group_render(frame_stride != 0) + mrt_group_reduce, then tonemap on the root.
 * Transports: MRT_GROUP_NCCL (ncclSend/ncclRecv/ncclReduce over NVLink; libnccl.so.2 is bound with dlopen when the
 * first group is made) or MRT_GROUP_P2P (cudaMemcpyPeerAsync; one process only; several contexts may share a device).
 * Calls on a group are made from one host thread.  Errors: mrt_group_last_error (g may be NULL: creation errors). */
typedef struct mrt_group mrt_group;
#define MRT_GROUP_NCCL 0
#define MRT_GROUP_P2P 1
/* one process drives ndev devices: creates one context per entry of devices[] (rank i = entry i); ncclCommInitAll */
int mrt_group_create(const int* devices, uint32_t ndev, int transport, mrt_group** out);
/* one process per GPU: rank 0 calls mrt_group_unique_id and hands the 128 bytes to the others by any means
 * (bench.py: torch.distributed); every process then calls mrt_group_create_rank (ncclCommInitRank, collective) */
int mrt_group_unique_id(uint8_t id_out[128]);
int mrt_group_create_rank(int device, uint32_t rank, uint32_t nranks, const uint8_t unique_id[128], mrt_group** out);
void mrt_group_destroy(mrt_group* g);  /* destroys the group's contexts too */
const char* mrt_group_last_error(const mrt_group* g);
int mrt_group_size(const mrt_group* g, uint32_t* nranks, uint32_t* nlocal);
/* the local_index-th context this process drives (borrowed; owned by the group) and its rank */
int mrt_group_context(mrt_group* g, uint32_t local_index, mrt_context** ctx_out, uint32_t* rank_out);
/* tile mode: mrt_set_partition(rank, nranks, slab_rows) on every local context */
int mrt_group_set_tiles(mrt_group* g, uint32_t slab_rows);
/* mrt_primary_rays + mrt_secondary_rays on every local context; rank r renders frameCounter + r * frame_stride
 * (0 in tile mode: all ranks render the same frame; N in sample-set mode: disjoint seeds) */
int mrt_group_render(mrt_group* g, uint32_t w, uint32_t h, const mrt_primary_constants* pc, const mrt_secondary_constants* sc,
                     uint32_t spp, uint32_t bounces, uint32_t flags, uint32_t frame_stride);
/* Frames in flight for progressive tile mode (the reference keeps 3 frames in flight, renderer.ixx:36): every local
 * rank gets `frames` (1..MRT_GROUP_MAX_FRAMES) frame contexts on its device; mrt_group_render calls that carry MRT_SECONDARY_FRAME_SUM go
 * round-robin over them and are committed (mrt_accum_commit) to the rank's context in call order, so consecutive
 * frames overlap on the GPU -- the drain of one frame's traversal launches is filled by the next frame's kernels -- and
 * the accumulated image is bit-identical for every `frames` and every number of ranks.  The caller prepares each
 * frame context like the rank's own (blue noise, mrt_scene_share from the rank's context, atmosphere, sky view);
 * partitions and traversal grid sizes are set by the group.  frames = 1 (default): the rank's context renders. */
#define MRT_GROUP_MAX_FRAMES 8
int mrt_group_set_frames_in_flight(mrt_group* g, uint32_t frames);
int mrt_group_frame_context(mrt_group* g, uint32_t local_index, uint32_t slot, mrt_context** ctx_out);
int mrt_group_tonemap(mrt_group* g, int mode, float exposure, const float* params, uint32_t nparams, int source);
/* tile mode: every rank's slabs of a per-pixel buffer (MRT_BUF_LDR, _ACCUM, _VISIBILITY, ...) -> the full image in
 * row order on rank `root`.  Asynchronous (exchange streams); collective over all processes of the group. */
int mrt_group_gather(mrt_group* g, int buffer_id, uint32_t root);
/* sample sets: MRT_BUF_ACCUM of all ranks summed onto rank `root` in place (ncclReduce) */
int mrt_group_reduce(mrt_group* g, uint32_t root);
/* the image of the last gather on the root process: borrowed device pointer + the stream it is ordered on, or a
 * blocking copy to host memory */
int mrt_group_result(mrt_group* g, void** device_ptr, size_t* bytes, void** stream_out);
int mrt_group_readback(mrt_group* g, void* host, size_t bytes);
/* non-blocking variant for frames in flight: the copy into (pinned) host memory is queued behind the last gather;
 * mrt_group_readback_wait returns once at most keep_in_flight queued copies are outstanding (0: all landed) */
int mrt_group_readback_async(mrt_group* g, void* host, size_t bytes);
int mrt_group_readback_wait(mrt_group* g, uint32_t keep_in_flight);
int mrt_group_sync(mrt_group* g);

#ifdef __cplusplus
}
#endif
#endif /* MINOTERT_H */
