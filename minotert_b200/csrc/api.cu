// api.cu -- the extern "C" entry points declared in include/minotert.h.
#include <stdarg.h>

#include <vector>

#include "context.cuh"

static char g_create_error[512] = "";

int mrt_fail(mrt_context* ctx, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char* dst = ctx ? ctx->err : g_create_error;
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    if (ctx && code == MRT_ERR_CUDA) ctx->poisoned = true;
    return code;
}

int mrt_check_cuda(mrt_context* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MRT_OK;
    return mrt_fail(ctx, e == cudaErrorMemoryAllocation ? MRT_ERR_OOM : MRT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define MRT_ENTER(ctx)                                                                        \
    do {                                                                                      \
        if (!(ctx)) return MRT_ERR_INVALID;                                                   \
        if ((ctx)->poisoned) return MRT_ERR_CUDA;                                             \
        cudaError_t _e = cudaSetDevice((ctx)->device);                                        \
        if (_e != cudaSuccess) return mrt_check_cuda((ctx), _e, "cudaSetDevice");             \
    } while (0)

namespace {

__global__ void k_accum_to_color16(const float4* __restrict__ accum, uint2* __restrict__ color16, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = accum[i];
    float3 c = a.w > 0.0f ? f3(a.x / a.w, a.y / a.w, a.z / a.w) : f3s(0.0f);
    uint2 pk;
    pk.x = (uint32_t)f32_to_f16_bits(c.x) | ((uint32_t)f32_to_f16_bits(c.y) << 16);
    pk.y = (uint32_t)f32_to_f16_bits(c.z) | (0x3C00u << 16);
    color16[i] = pk;
}

int buffer_info(mrt_context* ctx, int id, void** p, size_t* bytes) {
    const size_t n = ctx->npix;
    switch (id) {
    case MRT_BUF_VISIBILITY: if (!ctx->have_gbuffer) break; *p = ctx->visibility.p; *bytes = n * 4; return MRT_OK;
    case MRT_BUF_DEPTH: if (!ctx->have_gbuffer) break; *p = ctx->depth.p; *bytes = n * 2; return MRT_OK;
    case MRT_BUF_NORMAL: if (!ctx->have_gbuffer) break; *p = ctx->normal.p; *bytes = n * 8; return MRT_OK;
    case MRT_BUF_MOTION: if (!ctx->have_gbuffer) break; *p = ctx->motion.p; *bytes = n * 4; return MRT_OK;
    case MRT_BUF_COLOR:
        if (!ctx->have_accum) break;
        if (!ctx->have_color) {  // triangle path: resolve the accumulator into the reference's RGBA16F image
            k_accum_to_color16<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->accum.p, reinterpret_cast<uint2*>(ctx->color16.p), n);
            MRT_LAUNCHED(ctx);
            ctx->have_color = true;
        }
        *p = ctx->color16.p; *bytes = n * 8; return MRT_OK;
    case MRT_BUF_ACCUM: if (!ctx->have_accum) break; *p = ctx->accum.p; *bytes = n * 16; return MRT_OK;
    case MRT_BUF_LDR: if (!ctx->have_ldr) break; *p = ctx->ldr_buf[ctx->ldr_cur].p; *bytes = n * 4; return MRT_OK;
    case MRT_BUF_TRANSMITTANCE: if (!ctx->have_atmo) break; *p = ctx->trans16.p; *bytes = (size_t)MRT_TRANS_W * MRT_TRANS_H * 8; return MRT_OK;
    case MRT_BUF_MULTISCATTERING: if (!ctx->have_atmo) break; *p = ctx->multi16.p; *bytes = (size_t)MRT_MULTI_W * MRT_MULTI_H * 8; return MRT_OK;
    case MRT_BUF_AERIAL: if (!ctx->have_aerial) break; *p = ctx->aerial16.p; *bytes = (size_t)MRT_AERIAL_SIZE * MRT_AERIAL_SIZE * MRT_AERIAL_SIZE * 8; return MRT_OK;
    case MRT_BUF_SKY_VIEW:
        if (!ctx->have_view) break;
        if (ctx->sky_pending) {  // still being generated on the side stream: order the main stream after it
            cudaStreamWaitEvent(ctx->stream, ctx->sky_ready, 0);
            ctx->sky_pending = false;
        }
        *p = ctx->view_packed.p; *bytes = (size_t)MRT_VIEW_W * MRT_VIEW_H * 4; return MRT_OK;
    case MRT_BUF_BVH_NODES: if (ctx->scene_kind != 2 || !ctx->bvh_valid) break; *p = ctx->nodes.p; *bytes = (size_t)ctx->num_nodes * sizeof(WideNode); return MRT_OK;
    case MRT_BUF_BVH_TRIS: if (ctx->scene_kind != 2 || !ctx->bvh_valid) break; *p = ctx->tris.p; *bytes = (size_t)ctx->num_leaf_tris * 48; return MRT_OK;
    case MRT_BUF_DENOISED: if (!ctx->have_denoised) break; *p = ctx->denoised.p; *bytes = n * 4; return MRT_OK;
    case MRT_BUF_TEMPORAL: if (!ctx->have_temporal) break; *p = ctx->tp_rgba[ctx->tp_cur].p; *bytes = n * 16; return MRT_OK;
    case MRT_BUF_TEMPORAL_COUNT: if (!ctx->have_temporal) break; *p = ctx->tp_count[ctx->tp_cur].p; *bytes = n * 4; return MRT_OK;
    case MRT_BUF_HIT_T: if (!ctx->have_gbuffer || ctx->scene_kind != 2) break; *p = ctx->hit_t.p; *bytes = n * 4; return MRT_OK;
    default: return mrt_fail(ctx, MRT_ERR_INVALID, "unknown buffer id %d", id);
    }
    return mrt_fail(ctx, MRT_ERR_STATE, "buffer %d has not been rendered yet", id);
}

}  // namespace

// a context that renders another context's scene (mrt_scene_share) forgets the borrowed arrays instead of freeing them
static void scene_unborrow(mrt_context* ctx) {
    if (!ctx->scene_borrowed) return;
    if (ctx->scene_owner) {
        auto& list = ctx->scene_owner->borrowers;
        for (size_t i = 0; i < list.size(); i++)
            if (list[i] == ctx) { list.erase(list.begin() + (long)i); break; }
    }
    ctx->scene_owner = nullptr;
    ctx->pos = {}; ctx->idx = {}; ctx->albedo = {}; ctx->nodes = {}; ctx->tris = {};
    ctx->nverts = ctx->ntris = ctx->num_nodes = ctx->num_leaf_tris = 0;
    ctx->scene_borrowed = false;
    ctx->bvh_valid = false;
    if (ctx->scene_kind == 2) ctx->scene_kind = 0;  // a borrower that has since switched to a sphere scene keeps it
}

// Called before the owner's scene arrays are rewritten, reallocated or freed: frames of the borrowing contexts that
// still read them are waited for, and the borrowers are marked stale (mrt_scene_share again) or, when the owner goes
// away, left without a scene.
static void scene_borrowers_stale(mrt_context* owner, bool owner_dies) {
    if (owner->borrowers.empty()) return;
    std::vector<mrt_context*> list = owner->borrowers;
    for (mrt_context* b : list) {
        cudaStreamSynchronize(b->stream);
        if (b->scene_kind == 2) {  // (a borrower rendering its own sphere scene meanwhile is not touched)
            b->bvh_valid = false;
            b->have_gbuffer = b->have_accum = b->have_color = b->have_ldr = b->have_denoised = b->have_temporal = false;
        }
        if (owner_dies) scene_unborrow(b);
    }
}

extern "C" {

int mrt_abi_version(void) { return MRT_ABI_VERSION; }

int mrt_create(int device, mrt_context** out) {
    if (!out) return MRT_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return mrt_fail(nullptr, MRT_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return mrt_fail(nullptr, MRT_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return mrt_fail(nullptr, MRT_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    mrt_context* ctx = new mrt_context();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        mrt_fail(nullptr, MRT_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return MRT_ERR_CUDA;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);  // pairs: sky, primary, secondary, tonemap, denoise
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->sky_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ldr_ready, cudaEventDisableTiming);
    for (auto& ev : ctx->copy_done) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    *out = ctx;
    return MRT_OK;
}

void mrt_destroy(mrt_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->upload_stream) { cudaStreamSynchronize(ctx->upload_stream); cudaStreamDestroy(ctx->upload_stream); }
    cudaStreamSynchronize(ctx->stream);
    if (ctx->build_results_host) cudaFreeHost(ctx->build_results_host);
    if (ctx->ev_pos_ready) cudaEventDestroy(ctx->ev_pos_ready);
    if (ctx->ev_refit_done) cudaEventDestroy(ctx->ev_refit_done);
    if (ctx->ev_frames_done) cudaEventDestroy(ctx->ev_frames_done);
    if (ctx->bn) cudaFree(ctx->bn);
    scene_borrowers_stale(ctx, true);
    scene_unborrow(ctx);
    dev_free(ctx->pos); dev_free(ctx->idx); dev_free(ctx->albedo);
    dev_free(ctx->prim_lo); dev_free(ctx->prim_hi); dev_free(ctx->keys); dev_free(ctx->keys_alt);
    dev_free(ctx->order); dev_free(ctx->order_alt); dev_free(ctx->hist); dev_free(ctx->scan_tmp);
    dev_free(ctx->bin_left); dev_free(ctx->bin_right); dev_free(ctx->bin_parent); dev_free(ctx->bin_count);
    for (int k = 0; k < 2; k++) { dev_free(ctx->ploc_c[k]); dev_free(ctx->ploc_flag[k]); dev_free(ctx->ploc_scan[k]); }
    dev_free(ctx->ploc_nn);
    dev_free(ctx->bin_lo); dev_free(ctx->bin_hi); dev_free(ctx->bin_flag);
    dev_free(ctx->scene_bounds); dev_free(ctx->work_a); dev_free(ctx->work_b); dev_free(ctx->slot_node);
    dev_free(ctx->node_nchild); dev_free(ctx->node_ntri); dev_free(ctx->node_child_base); dev_free(ctx->node_tri_base);
    dev_free(ctx->node_lo); dev_free(ctx->node_hi); dev_free(ctx->level_starts_dev); dev_free(ctx->bin_rec);
    if (!ctx->scene_borrowed) { dev_free(ctx->nodes_alt); dev_free(ctx->tris_alt); }
    dev_free(ctx->nodes); dev_free(ctx->tris); dev_free(ctx->counters); dev_free(ctx->loop_sums);
    dev_free(ctx->trans16); dev_free(ctx->multi16); dev_free(ctx->view_packed);
    dev_free(ctx->trans_f); dev_free(ctx->multi_f); dev_free(ctx->view_f);
    dev_free(ctx->visibility); dev_free(ctx->depth); dev_free(ctx->normal); dev_free(ctx->motion); dev_free(ctx->color16);
    dev_free(ctx->denoised); dev_free(ctx->dn_taps);
    for (int k = 0; k < 2; k++) { dev_free(ctx->tp_rgba[k]); dev_free(ctx->tp_count[k]); dev_free(ctx->tp_vis[k]); }
    dev_free(ctx->hit_t); dev_free(ctx->accum); dev_free(ctx->frame_sum); dev_free(ctx->sun_e); dev_free(ctx->aerial16); dev_free(ctx->aerial_f); for (auto& q : ctx->shadow_q) dev_free(q); dev_free(ctx->ldr_buf[0]); dev_free(ctx->ldr_buf[1]);
    dev_free(ctx->hit0_pos); dev_free(ctx->hit0_n); dev_free(ctx->path_state);
    for (int q = 0; q < 2; q++) { dev_free(ctx->ray_o[q]); dev_free(ctx->ray_d[q]); dev_free(ctx->ray_p[q]); dev_free(ctx->ray_s[q]); }
    dev_free(ctx->hits); dev_free(ctx->queue_counts); dev_free(ctx->sort_keys); dev_free(ctx->sort_keys_alt);
    dev_free(ctx->sort_vals); dev_free(ctx->sort_vals_alt); dev_free(ctx->visit_counters); dev_free(ctx->total_rays);
    dev_free(ctx->query_o); dev_free(ctx->query_d); dev_free(ctx->query_t); dev_free(ctx->query_ids);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->trace_ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    for (int b = 0; b < MRT_MAX_BANDS; b++) {
        if (ctx->band_stream[b]) { cudaStreamSynchronize(ctx->band_stream[b]); cudaStreamDestroy(ctx->band_stream[b]); }
        if (ctx->band_done[b]) cudaEventDestroy(ctx->band_done[b]);
    }
    if (ctx->band_fork) cudaEventDestroy(ctx->band_fork);
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->sky_ready) cudaEventDestroy(ctx->sky_ready);
    if (ctx->ldr_ready) cudaEventDestroy(ctx->ldr_ready);
    for (auto& ev : ctx->commit_ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->copy_done) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mrt_last_error(const mrt_context* ctx) { return ctx ? ctx->err : g_create_error; }

int mrt_set_option(mrt_context* ctx, const char* name, int64_t value) {
    MRT_ENTER(ctx);
    if (!name) return mrt_fail(ctx, MRT_ERR_INVALID, "option name is NULL");
    if (!strcmp(name, "count_visits")) ctx->opt_count_visits = value != 0;
    else if (!strcmp(name, "sort_rays")) ctx->opt_sort_rays = (int)(value < 0 ? 0 : (value > 2 ? 2 : value));
    else if (!strcmp(name, "persistent_primary")) ctx->opt_persistent_primary = value != 0;
    else if (!strcmp(name, "primary_batched")) ctx->opt_primary_batched = value != 0;
    else if (!strcmp(name, "shadow_coherent")) ctx->opt_shadow_coherent = value != 0;
    else if (!strcmp(name, "trace_carveout")) ctx->opt_trace_carveout = (int)value;
    else if (!strcmp(name, "ray_split")) ctx->opt_ray_split = (int)value;
    else if (!strcmp(name, "shade_tiles")) ctx->opt_shade_tiles = value != 0;
    else if (!strcmp(name, "trace_timing")) ctx->opt_trace_timing = value != 0;
    else if (!strcmp(name, "fused_shade")) ctx->opt_fused_shade = value != 0;
    else if (!strcmp(name, "primary_entry")) ctx->opt_primary_entry = value != 0;
    else if (!strcmp(name, "path_kernel")) ctx->opt_path_kernel = value != 0;
    else if (!strcmp(name, "bands")) ctx->opt_bands = (int)(value < 1 ? 1 : (value > MRT_MAX_BANDS ? MRT_MAX_BANDS : value));
    else if (!strcmp(name, "trace_ctas_per_sm")) ctx->opt_trace_ctas_per_sm = (int)(value > 32 ? 32 : value);
    else if (!strcmp(name, "wide_refit")) ctx->opt_wide_refit = value != 0;
    else if (!strcmp(name, "prepared_rays")) ctx->opt_prepared_rays = value != 0;
    else if (!strcmp(name, "spheres_batched")) ctx->opt_spheres_batched = value < 0 ? 0 : value > 2 ? 2 : (int)value;
    else if (!strcmp(name, "fused_sort")) { ctx->opt_fused_sort = value != 0; ctx->bvh_valid = false; }
    else if (!strcmp(name, "async_update")) ctx->opt_async_update = value != 0;
    else if (!strcmp(name, "build_device_loop")) { ctx->opt_build_device_loop = value != 0; ctx->bvh_valid = false; }
    else if (!strcmp(name, "builder")) { ctx->opt_builder = value != 0; ctx->bvh_valid = false; }
    else if (!strcmp(name, "ploc_radius")) { ctx->opt_ploc_radius = (int)(value < 1 ? 1 : (value > 32 ? 32 : value)); ctx->bvh_valid = false; }
    else return mrt_fail(ctx, MRT_ERR_INVALID, "unknown option '%s'", name);
    return MRT_OK;
}

int mrt_upload_blue_noise(mrt_context* ctx, const uint8_t* rgba8, uint32_t w, uint32_t h) {
    MRT_ENTER(ctx);
    if (!rgba8 || w == 0 || h == 0) return mrt_fail(ctx, MRT_ERR_INVALID, "blue noise: empty texture");
    if (ctx->bn) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->bn); ctx->bn = nullptr; }
    MRT_CUDA(ctx, cudaMalloc((void**)&ctx->bn, (size_t)w * h * 4));
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->bn, rgba8, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bnW = w;
    ctx->bnH = h;
    return MRT_OK;
}

int mrt_scene_set_spheres(mrt_context* ctx, const mrt_sphere* spheres, uint32_t n) {
    MRT_ENTER(ctx);
    if (n > MRT_MAX_SPHERES) return mrt_fail(ctx, MRT_ERR_INVALID, "at most %d spheres", MRT_MAX_SPHERES);
    if (n && !spheres) return mrt_fail(ctx, MRT_ERR_INVALID, "spheres is NULL");
    memset(&ctx->spheres, 0, sizeof ctx->spheres);
    for (uint32_t i = 0; i < n; i++) ctx->spheres.s[i] = spheres[i];
    ctx->spheres.n = n;
    ctx->scene_kind = 1;
    ctx->have_gbuffer = ctx->have_accum = ctx->have_color = ctx->have_ldr = ctx->have_denoised = ctx->have_temporal = false;  // a new scene restarts accumulation
    return MRT_OK;
}

int mrt_scene_upload_mesh(mrt_context* ctx, const float* positions, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                          const float* albedo) {
    MRT_ENTER(ctx);
    if (ntris && (!positions || !indices || !albedo || nverts == 0))
        return mrt_fail(ctx, MRT_ERR_INVALID, "mesh: NULL array with ntris = %u", ntris);
    for (size_t i = 0; i < 3 * (size_t)ntris; i++)
        if (indices[i] >= nverts) return mrt_fail(ctx, MRT_ERR_INVALID, "mesh: index %u out of range at %zu", indices[i], i);
    if (ctx->scene_borrowed) { cudaStreamSynchronize(ctx->stream); scene_unborrow(ctx); }
    scene_borrowers_stale(ctx, false);
    MRT_TRY(dev_reserve(ctx, ctx->pos, 3 * (size_t)nverts));
    MRT_TRY(dev_reserve(ctx, ctx->idx, 3 * (size_t)ntris));
    MRT_TRY(dev_reserve(ctx, ctx->albedo, ntris));
    if (ntris) {
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->pos.p, positions, sizeof(float) * 3 * (size_t)nverts, cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->idx.p, indices, sizeof(uint32_t) * 3 * (size_t)ntris, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<float4> al(ntris);
        for (uint32_t i = 0; i < ntris; i++) al[i] = make_float4(albedo[3 * (size_t)i], albedo[3 * (size_t)i + 1], albedo[3 * (size_t)i + 2], 0.0f);
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->albedo.p, al.data(), sizeof(float4) * (size_t)ntris, cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->nverts = nverts;
    ctx->ntris = ntris;
    ctx->scene_kind = 2;
    ctx->bvh_valid = false;
    ctx->have_gbuffer = ctx->have_accum = ctx->have_color = ctx->have_ldr = ctx->have_denoised = ctx->have_temporal = false;  // a new scene restarts accumulation
    return MRT_OK;
}

// lazily created: a context that never updates asynchronously pays nothing
static int scene_async_objects(mrt_context* ctx) {
    if (!ctx->upload_stream) MRT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
    if (!ctx->ev_pos_ready) MRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_pos_ready, cudaEventDisableTiming));
    if (!ctx->ev_refit_done) MRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_refit_done, cudaEventDisableTiming));
    return MRT_OK;
}

int mrt_scene_update_positions(mrt_context* ctx, const float* positions, uint32_t nverts) {
    MRT_ENTER(ctx);
    if (ctx->scene_kind != 2) return mrt_fail(ctx, MRT_ERR_STATE, "no mesh uploaded");
    if (ctx->scene_borrowed) return mrt_fail(ctx, MRT_ERR_STATE, "the scene is borrowed (mrt_scene_share): update it through its owner");
    if (nverts != ctx->nverts || !positions) return mrt_fail(ctx, MRT_ERR_INVALID, "vertex count mismatch (%u vs %u)", nverts, ctx->nverts);
    if (ctx->opt_async_update) {
        // The vertex array is read by builds and refits only (rendering reads the leaf triangles), so the copy runs on
        // the scene-update stream beside the frames in flight, behind the refit that may still read the old positions.
        MRT_TRY(scene_async_objects(ctx));
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->pos.p, positions, sizeof(float) * 3 * (size_t)nverts, cudaMemcpyHostToDevice, ctx->upload_stream));
        MRT_CUDA(ctx, cudaEventRecord(ctx->ev_pos_ready, ctx->upload_stream));
        ctx->pos_upload_pending = true;
        return MRT_OK;
    }
    scene_borrowers_stale(ctx, false);
    if (ctx->pos_upload_pending) {  // an asynchronous upload queued before the option was switched off
        MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pos_ready, 0));
        ctx->pos_upload_pending = false;
    }
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->pos.p, positions, sizeof(float) * 3 * (size_t)nverts, cudaMemcpyHostToDevice, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

int mrt_scene_build(mrt_context* ctx, int build_mode) {
    MRT_ENTER(ctx);
    if (ctx->scene_kind != 2) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_scene_build: no mesh uploaded");
    if (ctx->scene_borrowed) return mrt_fail(ctx, MRT_ERR_STATE, "the scene is borrowed (mrt_scene_share): build it through its owner");
    if (build_mode != MRT_BUILD_REFIT && build_mode != MRT_BUILD_FULL) return mrt_fail(ctx, MRT_ERR_INVALID, "unknown build mode %d", build_mode);
    if (build_mode == MRT_BUILD_REFIT && ctx->opt_async_update && bvh_refit_can_be_async(ctx)) {
        // The host never blocks, and neither do the frames in flight: the tree exists twice, the refit runs on the
        // scene-update stream (behind the upload) into the copy nobody reads, and the copies swap roles.  Two event
        // edges order it against rendering:
        //   * the copy being rewritten was current until the PREVIOUS asynchronous refit: its last readers are the
        //     frames recorded before that call, marked by the events recorded then (and re-recorded now);
        //   * frames recorded from now on wait for this refit and read the new copy (the contexts that borrow the
        //     scene stay valid: same topology, same sizes -- only their two pointers move).
        MRT_TRY(scene_async_objects(ctx));
        cudaStream_t ss = ctx->upload_stream;
        std::vector<mrt_context*> all = ctx->borrowers;
        all.push_back(ctx);
        for (mrt_context* c : all) {
            if (!c->ev_frames_done) MRT_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_frames_done, cudaEventDisableTiming));
            if (c->frames_marked) MRT_CUDA(ctx, cudaStreamWaitEvent(ss, c->ev_frames_done, 0));
            MRT_CUDA(ctx, cudaEventRecord(c->ev_frames_done, c->stream));
            c->frames_marked = true;
        }
        if (!ctx->alt_valid) {  // first refit after a full build: the second copy takes over the topology
            MRT_TRY(dev_reserve(ctx, ctx->nodes_alt, ctx->num_nodes));
            MRT_TRY(dev_reserve(ctx, ctx->tris_alt, 3 * (size_t)ctx->ntris));
            MRT_CUDA(ctx, cudaMemcpyAsync(ctx->nodes_alt.p, ctx->nodes.p, sizeof(WideNode) * (size_t)ctx->num_nodes, cudaMemcpyDeviceToDevice, ss));
            MRT_CUDA(ctx, cudaMemcpyAsync(ctx->tris_alt.p, ctx->tris.p, sizeof(float4) * 3 * (size_t)ctx->ntris, cudaMemcpyDeviceToDevice, ss));
            ctx->alt_valid = true;
        }
        std::swap(ctx->nodes, ctx->nodes_alt);
        std::swap(ctx->tris, ctx->tris_alt);
        ctx->pos_upload_pending = false;  // same stream: the upload comes first
        MRT_TRY(bvh_refit(ctx, false, ss));
        MRT_CUDA(ctx, cudaEventRecord(ctx->ev_refit_done, ss));
        ctx->refit_recorded = true;
        for (mrt_context* c : all) {
            MRT_CUDA(ctx, cudaStreamWaitEvent(c->stream, ctx->ev_refit_done, 0));
            if (c != ctx) { c->nodes = ctx->nodes; c->tris = ctx->tris; }
        }
        return MRT_OK;
    }
    if (build_mode == MRT_BUILD_FULL && ctx->opt_async_update && ctx->bvh_valid && ctx->num_nodes != 0 && ctx->opt_sort_rays == 0) {
        // A rebuild with frames in flight: it needs the host (two counters are read back), but not the GPU's
        // attention -- it runs on the scene-update stream into the copy of the tree nobody reads, while the frames
        // already recorded keep rendering the current copy; the host blocks on the scene-update stream only.  The same two
        // event edges as for the refit above; afterwards the copies swap roles and the borrowers take over the new sizes.
        // (Not with the ray sort on: it shares the radix-sort scratch with the builder.)
        MRT_TRY(scene_async_objects(ctx));
        cudaStream_t ss = ctx->upload_stream;
        std::vector<mrt_context*> all = ctx->borrowers;
        all.push_back(ctx);
        for (mrt_context* c : all) {
            if (!c->ev_frames_done) MRT_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_frames_done, cudaEventDisableTiming));
            if (c->frames_marked) MRT_CUDA(ctx, cudaStreamWaitEvent(ss, c->ev_frames_done, 0));
            MRT_CUDA(ctx, cudaEventRecord(c->ev_frames_done, c->stream));
            c->frames_marked = true;
        }
        std::swap(ctx->nodes, ctx->nodes_alt);
        std::swap(ctx->tris, ctx->tris_alt);
        ctx->pos_upload_pending = false;  // same stream: the upload comes first
        cudaStream_t main_stream = ctx->stream;
        ctx->stream = ss;
        const int s = bvh_build_full(ctx);  // leaves alt_valid false: the other copy holds the old topology
        ctx->stream = main_stream;
        if (s != MRT_OK) return s;
        MRT_CUDA(ctx, cudaEventRecord(ctx->ev_refit_done, ss));
        ctx->refit_recorded = true;
        for (mrt_context* c : all) {
            MRT_CUDA(ctx, cudaStreamWaitEvent(c->stream, ctx->ev_refit_done, 0));
            if (c == ctx) continue;
            c->nodes = ctx->nodes; c->tris = ctx->tris;
            c->num_nodes = ctx->num_nodes; c->num_leaf_tris = ctx->num_leaf_tris;
            c->stats.num_wide_nodes = ctx->stats.num_wide_nodes; c->stats.bvh_bytes = ctx->stats.bvh_bytes;
            c->stats.sah_node_cost = ctx->stats.sah_node_cost; c->stats.sah_tri_cost = ctx->stats.sah_tri_cost;
        }
        return MRT_OK;
    }
    scene_borrowers_stale(ctx, false);
    if (ctx->pos_upload_pending) {  // positions uploaded asynchronously: the build reads them on the main stream
        MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pos_ready, 0));
        ctx->pos_upload_pending = false;
    }
    // (an asynchronous refit still running has every main stream waiting for it already; builds are synchronous, so a
    // later asynchronous upload cannot overtake this one)
    return build_mode == MRT_BUILD_REFIT ? bvh_refit(ctx, true) : bvh_build_full(ctx);
}

int mrt_scene_share(mrt_context* ctx, mrt_context* owner) {
    MRT_ENTER(ctx);
    if (!owner || owner == ctx) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_scene_share: owner is NULL or the context itself");
    if (owner->device != ctx->device) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_scene_share: contexts live on devices %d and %d", ctx->device, owner->device);
    if (owner->scene_borrowed) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_scene_share: the owner borrows its scene itself");
    if (owner->scene_kind != 2 || !owner->bvh_valid) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_scene_share: the owner has no built mesh scene");
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // frames of this context that still read the old scene
    if (ctx->scene_borrowed) scene_unborrow(ctx);
    else {
        // ctx's own arrays are about to be freed: contexts that alias them lose their scene first
        scene_borrowers_stale(ctx, true);
        dev_free(ctx->pos); dev_free(ctx->idx); dev_free(ctx->albedo); dev_free(ctx->nodes); dev_free(ctx->tris);
    }
    ctx->pos = owner->pos; ctx->idx = owner->idx; ctx->albedo = owner->albedo; ctx->nodes = owner->nodes; ctx->tris = owner->tris;
    ctx->nverts = owner->nverts; ctx->ntris = owner->ntris;
    ctx->num_nodes = owner->num_nodes; ctx->num_leaf_tris = owner->num_leaf_tris;
    ctx->stats.num_triangles = owner->stats.num_triangles; ctx->stats.num_wide_nodes = owner->stats.num_wide_nodes;
    ctx->stats.bvh_bytes = owner->stats.bvh_bytes;
    ctx->stats.sah_node_cost = owner->stats.sah_node_cost; ctx->stats.sah_tri_cost = owner->stats.sah_tri_cost;
    ctx->scene_borrowed = true;
    ctx->scene_owner = owner;
    owner->borrowers.push_back(ctx);
    ctx->scene_kind = 2;
    ctx->bvh_valid = true;
    ctx->have_gbuffer = ctx->have_accum = ctx->have_color = ctx->have_ldr = ctx->have_denoised = ctx->have_temporal = false;
    // builds and refits already queued on the owner's stream come first
    cudaEvent_t built;
    MRT_CUDA(ctx, cudaEventCreateWithFlags(&built, cudaEventDisableTiming));
    cudaEventRecord(built, owner->stream);
    cudaStreamWaitEvent(ctx->stream, built, 0);
    cudaEventDestroy(built);
    return MRT_OK;
}

// the main stream waits for a sky view still being generated on the side stream
static void sky_join(mrt_context* ctx) {
    if (ctx->sky_pending) {
        cudaStreamWaitEvent(ctx->stream, ctx->sky_ready, 0);
        ctx->sky_pending = false;
    }
}

int mrt_atmosphere(mrt_context* ctx, const mrt_atmosphere_params* params) {
    MRT_ENTER(ctx);
    if (!params) return mrt_fail(ctx, MRT_ERR_INVALID, "atmosphere params is NULL");
    sky_join(ctx);
    ctx->atmo = *params;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    MRT_TRY(sky_gen_atmosphere(ctx));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->have_atmo = true;
    ctx->have_view = false;
    return MRT_OK;
}

int mrt_sky_view(mrt_context* ctx, const float probePos[3], const float sunDirection[3], const float sunIlluminance[3]) {
    MRT_ENTER(ctx);
    if (!ctx->have_atmo) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_sky_view before mrt_atmosphere");
    if (!probePos || !sunDirection || !sunIlluminance) return mrt_fail(ctx, MRT_ERR_INVALID, "sky view: NULL argument");
    // side stream: ordered after everything issued so far (atmosphere LUTs, the previous view's readers), then
    // free to overlap whatever the caller issues next on the main stream until a reader joins
    MRT_CUDA(ctx, cudaEventRecord(ctx->aux_fork, ctx->stream));
    MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
    MRT_TRY(sky_gen_view(ctx, probePos, sunDirection, sunIlluminance));
    MRT_CUDA(ctx, cudaEventRecord(ctx->sky_ready, ctx->aux_stream));
    ctx->sky_pending = true;
    ctx->have_view = true;
    return MRT_OK;
}

int mrt_sky_aerial_perspective(mrt_context* ctx, const mrt_primary_constants* c, const float cameraPos[3],
                               const float sunDirection[3], const float sunIlluminance[3]) {
    MRT_ENTER(ctx);
    if (!ctx->have_atmo) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_sky_aerial_perspective before mrt_atmosphere");
    if (!c || !cameraPos || !sunDirection || !sunIlluminance) return mrt_fail(ctx, MRT_ERR_INVALID, "aerial perspective: NULL argument");
    MRT_TRY(sky_gen_aerial(ctx, c, cameraPos, sunDirection, sunIlluminance));
    ctx->have_aerial = true;
    return MRT_OK;
}

int mrt_set_partition(mrt_context* ctx, uint32_t rank, uint32_t nranks, uint32_t slab_rows) {
    MRT_ENTER(ctx);
    if (nranks == 0 || rank >= nranks || slab_rows == 0) return mrt_fail(ctx, MRT_ERR_INVALID, "bad partition %u/%u slab %u", rank, nranks, slab_rows);
    ctx->part = Partition{rank, nranks, slab_rows};
    ctx->have_gbuffer = ctx->have_accum = ctx->have_color = ctx->have_ldr = ctx->have_denoised = ctx->have_temporal = false;
    return MRT_OK;
}

int mrt_partition_rows(const mrt_context* ctx, uint32_t full_h, uint32_t* rows_out, uint32_t* nrows_out) {
    if (!ctx || !nrows_out) return MRT_ERR_INVALID;
    uint32_t n = partition_local_rows(ctx->part, full_h);
    *nrows_out = n;
    if (rows_out)
        for (uint32_t lr = 0; lr < n; lr++) rows_out[lr] = partition_local_to_y(ctx->part, lr);
    return MRT_OK;
}

int mrt_partition_rows_for(uint32_t rank, uint32_t nranks, uint32_t slab_rows, uint32_t full_h, uint32_t* rows_out,
                           uint32_t* nrows_out) {
    if (!nrows_out || nranks == 0 || rank >= nranks || slab_rows == 0) return MRT_ERR_INVALID;
    Partition p{rank, nranks, slab_rows};
    uint32_t n = partition_local_rows(p, full_h);
    *nrows_out = n;
    if (rows_out)
        for (uint32_t lr = 0; lr < n; lr++) rows_out[lr] = partition_local_to_y(p, lr);
    return MRT_OK;
}

int mrt_primary_rays(mrt_context* ctx, uint32_t w, uint32_t h, const mrt_primary_constants* c) {
    MRT_ENTER(ctx);
    if (!c || w == 0 || h == 0) return mrt_fail(ctx, MRT_ERR_INVALID, "primary rays: bad size %ux%u or NULL constants", w, h);
    if (ctx->scene_kind == 0) return mrt_fail(ctx, MRT_ERR_STATE, "primary rays: no scene");
    if (ctx->scene_kind == 2 && !ctx->bvh_valid)
        return mrt_fail(ctx, MRT_ERR_STATE, ctx->scene_borrowed ? "primary rays: the borrowed scene changed (call mrt_scene_share again)"
                                                                 : "primary rays: mesh uploaded but not built");
    uint32_t rows = partition_local_rows(ctx->part, h);
    if (w != ctx->W || h != ctx->H || rows != ctx->local_rows)
        ctx->have_accum = ctx->have_color = ctx->have_ldr = ctx->have_denoised = ctx->have_temporal = false;
    ctx->W = w; ctx->H = h; ctx->local_rows = rows;
    ctx->npix = (size_t)w * rows;
    ctx->pc = *c;
    size_t n = ctx->npix;
    MRT_TRY(dev_reserve(ctx, ctx->visibility, n));
    MRT_TRY(dev_reserve(ctx, ctx->depth, n));
    MRT_TRY(dev_reserve(ctx, ctx->normal, 4 * n));
    MRT_TRY(dev_reserve(ctx, ctx->motion, 2 * n));
    MRT_TRY(dev_reserve(ctx, ctx->color16, 4 * n));
    MRT_TRY(dev_reserve(ctx, ctx->accum, n));
    cudaEventRecord(ctx->ev[2], ctx->stream);
    int s = n == 0 ? MRT_OK : (ctx->scene_kind == 1 ? spheres_primary(ctx) : mesh_primary(ctx));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    if (s == MRT_OK) ctx->have_gbuffer = true;
    return s;
}

int mrt_secondary_rays(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp, uint32_t bounces, uint32_t flags) {
    MRT_ENTER(ctx);
    if (!c || spp == 0) return mrt_fail(ctx, MRT_ERR_INVALID, "secondary rays: NULL constants or spp = 0");
    if (!ctx->have_gbuffer) return mrt_fail(ctx, MRT_ERR_STATE, "secondary rays before primary rays");
    if (!ctx->have_atmo || !ctx->have_view) return mrt_fail(ctx, MRT_ERR_STATE, "secondary rays: sky LUTs missing (mrt_atmosphere, mrt_sky_view)");
    if (!ctx->bn) return mrt_fail(ctx, MRT_ERR_STATE, "secondary rays: blue noise texture missing");
    if ((flags & MRT_SECONDARY_AERIAL) && !ctx->have_aerial)
        return mrt_fail(ctx, MRT_ERR_STATE, "secondary rays: MRT_SECONDARY_AERIAL before mrt_sky_aerial_perspective");
    if ((flags & (MRT_SECONDARY_FRAME_SUM | MRT_SECONDARY_NEE_SUN | MRT_SECONDARY_SKY_AT_HIT | MRT_SECONDARY_AERIAL)) && ctx->scene_kind != 2)
        return mrt_fail(ctx, MRT_ERR_INVALID, "secondary rays: flags 0x%x are for triangle scenes", flags);
    sky_join(ctx);
    cudaEventRecord(ctx->ev[4], ctx->stream);
    int s = MRT_OK;
    if (ctx->npix) s = ctx->scene_kind == 1 ? spheres_secondary(ctx, c, spp, bounces, flags) : mesh_secondary(ctx, c, spp, bounces, flags);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    if (s == MRT_OK) {
        if (flags & MRT_SECONDARY_FRAME_SUM) {
            ctx->have_frame_sum = true;  // MRT_BUF_ACCUM is untouched until mrt_accum_commit
        } else {
            ctx->have_accum = true;
            ctx->have_color = ctx->scene_kind == 1;
            ctx->have_denoised = false;
        }
        ctx->secondary_done = true;
        ctx->stats.secondary_rays = ~0ull;  // resolved lazily in mrt_stats_get
    }
    return s;
}

namespace {
// dst (+)= src, both float4 (xyz radiance sums, w samples); plain fp32 adds in a fixed order: frames committed in frame
// order give the same bits whichever context rendered them
__global__ void __launch_bounds__(256) k_accum_commit(float4* __restrict__ dst, const float4* __restrict__ src, size_t n, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 f = src[i];
        float4 a = accumulate ? dst[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        dst[i] = make_float4(a.x + f.x, a.y + f.y, a.z + f.z, a.w + f.w);
    }
}
}  // namespace

int mrt_accum_commit(mrt_context* dst, mrt_context* src, uint32_t flags) {
    MRT_ENTER(dst);
    if (!src) return mrt_fail(dst, MRT_ERR_INVALID, "mrt_accum_commit: src is NULL");
    if (src->poisoned) return mrt_fail(dst, MRT_ERR_CUDA, "mrt_accum_commit: the source context is poisoned: %s", src->err);
    if (!src->have_frame_sum) return mrt_fail(dst, MRT_ERR_STATE, "mrt_accum_commit: the source holds no frame (render it with MRT_SECONDARY_FRAME_SUM)");
    if (src->device != dst->device) return mrt_fail(dst, MRT_ERR_INVALID, "mrt_accum_commit: contexts of different devices (%d, %d)", dst->device, src->device);
    if (flags & ~MRT_SECONDARY_ACCUMULATE) return mrt_fail(dst, MRT_ERR_INVALID, "mrt_accum_commit: unknown flags 0x%x", flags);
    if (dst != src && (dst->W != src->W || dst->H != src->H || dst->local_rows != src->local_rows || dst->npix != src->npix ||
                       dst->part.rank != src->part.rank || dst->part.nranks != src->part.nranks || dst->part.slab_rows != src->part.slab_rows)) {
        // dst becomes a display context of src's image: same size and partition, nothing rendered yet
        dst->part = src->part;
        dst->W = src->W; dst->H = src->H; dst->local_rows = src->local_rows; dst->npix = src->npix;
        dst->have_gbuffer = dst->have_accum = dst->have_color = dst->have_ldr = dst->have_denoised = dst->have_temporal = false;
    }
    const size_t n = src->npix;
    MRT_TRY(dev_reserve(dst, dst->accum, n));
    const int accumulate = ((flags & MRT_SECONDARY_ACCUMULATE) && dst->have_accum) ? 1 : 0;
    if (dst != src) {
        for (mrt_context* c : {dst, src})
            for (int k = 0; k < 2; k++)
                if (!c->commit_ev[k]) MRT_CUDA(dst, cudaEventCreateWithFlags(&c->commit_ev[k], cudaEventDisableTiming));
        MRT_CUDA(dst, cudaEventRecord(src->commit_ev[0], src->stream));
        MRT_CUDA(dst, cudaStreamWaitEvent(dst->stream, src->commit_ev[0], 0));
    }
    if (n) {
        k_accum_commit<<<(unsigned)std::min<size_t>(div_up(n, 256), 148 * 16), 256, 0, dst->stream>>>(dst->accum.p, src->frame_sum.p, n, accumulate);
        MRT_LAUNCHED(dst);
        MRT_CUDA(dst, cudaGetLastError());
    }
    if (dst != src) {  // src may overwrite its frame buffer only after the add has read it
        MRT_CUDA(dst, cudaEventRecord(src->commit_ev[1], dst->stream));
        MRT_CUDA(dst, cudaStreamWaitEvent(src->stream, src->commit_ev[1], 0));
    }
    src->have_frame_sum = false;
    dst->have_accum = true;
    dst->have_color = false;
    dst->have_denoised = false;
    return MRT_OK;
}

int mrt_denoise_bilateral(mrt_context* ctx, float sigma, float kSigma, float threshold, float nearPlane, uint32_t frameCounter) {
    MRT_ENTER(ctx);
    if (!ctx->have_gbuffer || !ctx->have_accum) return mrt_fail(ctx, MRT_ERR_STATE, "denoise before primary + secondary rays");
    if (ctx->part.nranks > 1) return mrt_fail(ctx, MRT_ERR_STATE, "denoise needs the whole image in one context (partition %u of %u)", ctx->part.rank, ctx->part.nranks);
    if (!ctx->have_color) {  // triangle path: resolve the accumulator into the RGBA16F image the filter reads
        void* p; size_t b;
        MRT_TRY(buffer_info(ctx, MRT_BUF_COLOR, &p, &b));
    }
    cudaEventRecord(ctx->ev[8], ctx->stream);
    int s = denoise_bilateral(ctx, sigma, kSigma, threshold, nearPlane, frameCounter);
    cudaEventRecord(ctx->ev[9], ctx->stream);
    if (s == MRT_OK) ctx->have_denoised = true;
    return s;
}

int mrt_temporal_accumulate(mrt_context* ctx, float maxHistory, uint32_t flags) {
    MRT_ENTER(ctx);
    if (!ctx->have_gbuffer || !ctx->have_accum) return mrt_fail(ctx, MRT_ERR_STATE, "temporal accumulation before primary + secondary rays");
    if (ctx->part.nranks > 1) return mrt_fail(ctx, MRT_ERR_STATE, "temporal accumulation needs the whole image in one context (partition %u of %u)", ctx->part.rank, ctx->part.nranks);
    if (!(maxHistory >= 0.0f)) return mrt_fail(ctx, MRT_ERR_INVALID, "temporal accumulation: maxHistory %g", (double)maxHistory);
    if (flags & ~MRT_TEMPORAL_RESET) return mrt_fail(ctx, MRT_ERR_INVALID, "temporal accumulation: unknown flags 0x%x", flags);
    cudaEventRecord(ctx->ev[10], ctx->stream);
    int s = ctx->npix ? temporal_accumulate(ctx, maxHistory, (flags & MRT_TEMPORAL_RESET) != 0) : MRT_OK;
    cudaEventRecord(ctx->ev[11], ctx->stream);
    if (s == MRT_OK) ctx->have_temporal = true;
    return s;
}

int mrt_tonemap(mrt_context* ctx, int mode, float exposure, const float* params, uint32_t nparams, int source) {
    MRT_ENTER(ctx);
    if (mode < MRT_TONEMAP_LINEAR || mode > MRT_TONEMAP_AMD) return mrt_fail(ctx, MRT_ERR_INVALID, "unknown tonemap mode %d", mode);
    static const uint32_t need[6] = {0, 1, 0, 0, 6, 5};
    if (nparams < need[mode] || (need[mode] && !params)) return mrt_fail(ctx, MRT_ERR_INVALID, "tonemap mode %d needs %u params", mode, need[mode]);
    if (source != MRT_BUF_COLOR && source != MRT_BUF_ACCUM && source != MRT_BUF_DENOISED && source != MRT_BUF_TEMPORAL)
        return mrt_fail(ctx, MRT_ERR_INVALID, "tonemap source must be COLOR, ACCUM, DENOISED or TEMPORAL");
    if (!ctx->have_accum) return mrt_fail(ctx, MRT_ERR_STATE, "tonemap before secondary rays");
    if (source == MRT_BUF_DENOISED && !ctx->have_denoised) return mrt_fail(ctx, MRT_ERR_STATE, "tonemap of the denoised image before mrt_denoise_bilateral");
    if (source == MRT_BUF_TEMPORAL && !ctx->have_temporal) return mrt_fail(ctx, MRT_ERR_STATE, "tonemap of the temporal image before mrt_temporal_accumulate");
    // source COLOR on the triangle path while the RGBA16F image has not been asked for: tonemap_run reads the
    // accumulator and rounds through fp16 in registers -- same bits as resolving into MRT_BUF_COLOR first, one launch
    // and 24 B/px less (the image is still resolved lazily when mrt_buffer / the denoiser ask for it)
    cudaEventRecord(ctx->ev[6], ctx->stream);
    int s = tonemap_run(ctx, mode, exposure, params, nparams, source);
    cudaEventRecord(ctx->ev[7], ctx->stream);
    return s;
}

int mrt_buffer(mrt_context* ctx, int buffer_id, void** device_ptr, size_t* bytes) {
    MRT_ENTER(ctx);
    if (!device_ptr || !bytes) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_buffer: NULL out pointer");
    return buffer_info(ctx, buffer_id, device_ptr, bytes);
}

int mrt_readback(mrt_context* ctx, int buffer_id, void* host, size_t bytes) {
    MRT_ENTER(ctx);
    void* p = nullptr;
    size_t have = 0;
    MRT_TRY(buffer_info(ctx, buffer_id, &p, &have));
    if (!host || bytes > have) return mrt_fail(ctx, MRT_ERR_INVALID, "readback of %zu bytes from a %zu-byte buffer", bytes, have);
    MRT_CUDA(ctx, cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

// Asynchronous framebuffer readback: the copy runs on a second stream as soon as the tonemap that produced
// the current LDR buffer is done, so the host can already issue the next frame (which tonemaps into the
// other LDR buffer).  mrt_readback_wait blocks until every outstanding async readback has landed.
int mrt_readback_async(mrt_context* ctx, int buffer_id, void* host, size_t bytes) {
    MRT_ENTER(ctx);
    if (buffer_id != MRT_BUF_LDR) return mrt_fail(ctx, MRT_ERR_INVALID, "async readback is only available for MRT_BUF_LDR");
    void* p = nullptr;
    size_t have = 0;
    MRT_TRY(buffer_info(ctx, buffer_id, &p, &have));
    if (!host || bytes > have) return mrt_fail(ctx, MRT_ERR_INVALID, "readback of %zu bytes from a %zu-byte buffer", bytes, have);
    MRT_CUDA(ctx, cudaEventRecord(ctx->ldr_ready, ctx->stream));
    MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ldr_ready, 0));
    MRT_CUDA(ctx, cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    MRT_CUDA(ctx, cudaEventRecord(ctx->copy_done[ctx->ldr_cur], ctx->copy_stream));
    ctx->copy_pending[ctx->ldr_cur] = true;
    return MRT_OK;
}

int mrt_readback_wait(mrt_context* ctx, int frames_in_flight) {
    MRT_ENTER(ctx);
    const int newest = ctx->ldr_cur, older = ctx->ldr_cur ^ 1;
    if (ctx->copy_pending[older]) {
        MRT_CUDA(ctx, cudaEventSynchronize(ctx->copy_done[older]));
        ctx->copy_pending[older] = false;
    }
    if (frames_in_flight <= 0 && ctx->copy_pending[newest]) {
        MRT_CUDA(ctx, cudaEventSynchronize(ctx->copy_done[newest]));
        ctx->copy_pending[newest] = false;
    }
    return MRT_OK;
}

int mrt_sync(mrt_context* ctx) {
    MRT_ENTER(ctx);
    sky_join(ctx);
    if (ctx->upload_stream) MRT_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));  // option async_update: the caller's vertex array is free again
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

int mrt_stream(mrt_context* ctx, void** stream_out) {
    MRT_ENTER(ctx);
    if (!stream_out) return MRT_ERR_INVALID;
    *stream_out = (void*)ctx->stream;
    return MRT_OK;
}

int mrt_stats_get(mrt_context* ctx, mrt_stats* out) {
    MRT_ENTER(ctx);
    if (!out) return MRT_ERR_INVALID;
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms;
    if (ctx->have_atmo && cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->stats.ms_sky = ms;
    if (ctx->have_gbuffer && cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->stats.ms_primary = ms;
    if (ctx->secondary_done && cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]) == cudaSuccess) ctx->stats.ms_secondary = ms;
    if (ctx->have_ldr && cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]) == cudaSuccess) ctx->stats.ms_tonemap = ms;
    if (ctx->have_denoised && cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]) == cudaSuccess) ctx->stats.ms_denoise = ms;
    if (ctx->have_temporal && cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]) == cudaSuccess) ctx->stats.ms_temporal = ms;
    if (ctx->build_time_pending && cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]) == cudaSuccess) { ctx->stats.ms_build = ms; ctx->build_time_pending = false; }
    cudaGetLastError();
    if (ctx->secondary_done && ctx->stats.secondary_rays == ~0ull) {
        if (ctx->scene_kind == 1) {
            unsigned long long v = 0;
            MRT_CUDA(ctx, cudaMemcpy(&v, ctx->visit_counters.p + 4, sizeof v, cudaMemcpyDeviceToHost));
            ctx->stats.secondary_rays = v;
        } else if (ctx->secondary_was_path_kernel) {
            unsigned long long v = 0;
            MRT_CUDA(ctx, cudaMemcpy(&v, ctx->visit_counters.p + 4 + 3, sizeof v, cudaMemcpyDeviceToHost));
            ctx->stats.secondary_rays = v;
        } else {
            std::vector<uint32_t> counts(ctx->num_queue_counts);
            MRT_CUDA(ctx, cudaMemcpy(counts.data(), ctx->queue_counts.p, sizeof(uint32_t) * counts.size(), cudaMemcpyDeviceToHost));
            uint64_t total = 0;
            for (uint32_t v : counts) total += v;
            if (ctx->num_back_counts) {    // option ray_split: the queues' back ends
                counts.resize(ctx->num_back_counts);
                MRT_CUDA(ctx, cudaMemcpy(counts.data(), ctx->queue_counts.p + ctx->back_counts_at, sizeof(uint32_t) * counts.size(), cudaMemcpyDeviceToHost));
                for (uint32_t v : counts) total += v;
            }
            if (ctx->num_shadow_counts) {  // MRT_SECONDARY_NEE_SUN: shadow rays are traced rays too
                counts.resize(ctx->num_shadow_counts);
                MRT_CUDA(ctx, cudaMemcpy(counts.data(), ctx->queue_counts.p + ctx->shadow_counts_at, sizeof(uint32_t) * counts.size(), cudaMemcpyDeviceToHost));
                for (uint32_t v : counts) total += v;
            }
            ctx->stats.secondary_rays = total;
        }
    }
    ctx->stats.ms_trace = 0.0f;
    ctx->stats.trace_launches = ctx->scene_kind == 2 ? ctx->trace_ev_used : 0;
    if (ctx->scene_kind == 2 && ctx->secondary_done)
        for (uint32_t i = 0; i < ctx->trace_ev_used; i++)
            if (cudaEventElapsedTime(&ms, ctx->trace_ev[2 * i], ctx->trace_ev[2 * i + 1]) == cudaSuccess) ctx->stats.ms_trace += ms;
    cudaGetLastError();
    if (ctx->scene_kind == 2 && ctx->visit_counters.p) {
        unsigned long long vc[8];
        MRT_CUDA(ctx, cudaMemcpy(vc, ctx->visit_counters.p, sizeof vc, cudaMemcpyDeviceToHost));
        ctx->stats.node_visits = vc[0] + vc[4];
        ctx->stats.tri_tests = vc[1] + vc[5];
        ctx->stats.secondary_node_visits = vc[4];
        ctx->stats.secondary_tri_tests = vc[5];
        ctx->stats.stack_overflows = (uint32_t)(vc[2] + vc[6]);
    }
    if (ctx->total_rays.p) {
        unsigned long long t = 0;
        MRT_CUDA(ctx, cudaMemcpy(&t, ctx->total_rays.p, sizeof t, cudaMemcpyDeviceToHost));
        ctx->stats.total_rays = t;
    }
    *out = ctx->stats;
    return MRT_OK;
}

int mrt_stats_reset(mrt_context* ctx) {
    MRT_ENTER(ctx);
    ctx->stats.kernel_launches = 0;
    ctx->trace_ev_used = 0;
    MRT_TRY(dev_reserve(ctx, ctx->total_rays, 1));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->total_rays.p, 0, sizeof(unsigned long long), ctx->stream));
    return MRT_OK;
}

int mrt_trace_rays(mrt_context* ctx, const float* origins, const float* directions, uint32_t n, uint32_t* prim_ids, float* t,
                   int brute_force) {
    MRT_ENTER(ctx);
    if (ctx->scene_kind != 2) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_trace_rays: no mesh uploaded");
    if (!brute_force && !ctx->bvh_valid) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_trace_rays: BVH not built");
    if (n && (!origins || !directions || !prim_ids || !t)) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_trace_rays: NULL array");
    return mesh_trace_rays(ctx, origins, directions, n, prim_ids, t, brute_force);
}

int mrt_accum_restore(mrt_context* ctx, const float* rgba32f, size_t bytes) {
    MRT_ENTER(ctx);
    if (!ctx->have_gbuffer) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_accum_restore before mrt_primary_rays (the image size is not known yet)");
    if (!rgba32f || bytes != ctx->npix * sizeof(float4))
        return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_accum_restore: %zu bytes for a %zu-byte accumulator", bytes, ctx->npix * sizeof(float4));
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->accum.p, rgba32f, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffer is consumed before the call returns
    ctx->have_accum = true;
    ctx->have_color = false;
    return MRT_OK;
}

int mrt_eval_sky_color(mrt_context* ctx, const float cameraPos[3], const float* directions, uint32_t n, float* rgb_out) {
    MRT_ENTER(ctx);
    if (!ctx->have_atmo || !ctx->have_view) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_eval_sky_color: sky LUTs missing (mrt_atmosphere, mrt_sky_view)");
    if (!cameraPos || (n && (!directions || !rgb_out))) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_eval_sky_color: NULL array");
    sky_join(ctx);
    return probe_sky_color(ctx, cameraPos, directions, n, rgb_out);
}

int mrt_eval_bounce_stream(mrt_context* ctx, uint32_t frameCounter, uint32_t x, uint32_t y, const float position[3],
                           const float normal[3], uint32_t n, float* out9) {
    MRT_ENTER(ctx);
    if (!ctx->bn) return mrt_fail(ctx, MRT_ERR_STATE, "mrt_eval_bounce_stream: blue noise texture missing");
    if (!position || !normal || (n && !out9)) return mrt_fail(ctx, MRT_ERR_INVALID, "mrt_eval_bounce_stream: NULL array");
    return probe_bounce_stream(ctx, frameCounter, x, y, position, normal, n, out9);
}

}  // extern "C"
