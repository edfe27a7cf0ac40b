// spheres.cu -- the reference's own scene (src/gpu/scene.glsl) through the reference's two-pass
// structure, "faithful" mode: primaryRay.comp -> fp16 G-buffer -> secondaryRays.comp -> RGBA16F.
// This path is ALU/SFU-bound (5 analytic spheres in kernel parameters, no scene memory traffic):
// one thread per pixel, 8x8 sample/bounce loop kept in registers exactly as the reference does.
#include "shading.cuh"

namespace {

// primaryRay.comp:23-76
__global__ void __launch_bounds__(256)
k_spheres_primary(RayGen g, Mat4 PV, Mat4 PVprev, Spheres sp, Partition part, uint32_t local_rows, uint32_t* vis,
                  uint16_t* depth, uint16_t* normal, uint16_t* motion) {
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, lr = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || lr >= local_rows) return;
    uint32_t y = partition_local_to_y(part, lr);
    float3 o, d;
    ray_gen(g, x, y, o, d);
    uint32_t id = MRT_MISS_ID;
    float best = 0.0f;
    for (uint32_t i = 0; i < sp.n; i++) {
        float t = ray_sphere(o, d, sp.s[i]);
        if (t >= 0.0f && (id == MRT_MISS_ID || t < best)) { best = t; id = i; }
    }
    // miss: depth 0 (infinitely far, inverted-Z), motion 0, normal = ray direction (:69-70)
    float dep = 0.0f;
    float2 mo = make_float2(0.0f, 0.0f);
    float3 n = d;
    if (id != MRT_MISS_ID) {
        float3 pos = o + d * best;
        n = normalize3(pos - f3(sp.s[id].center[0], sp.s[id].center[1], sp.s[id].center[2]));
        project_hit(PV, PVprev, pos, g.W, g.H, dep, mo);
    }
    store_gbuffer(vis, depth, normal, motion, (size_t)lr * g.W + x, id, dep, n, mo);
}

struct SecondaryParams {
    Mat4 invView, invProj;
    float3 cameraPos;
    uint32_t frameCounter;
    uint32_t W, H, local_rows, spp, bounces;
    uint32_t bnW, bnH;
};

// secondaryRays.comp:64-135 with Samples/Bounces as parameters.
// BATCHED = false: the shader's loops as they stand -- for each sample, for each bounce -- one pixel per lane.  Paths end
// at very different bounces, and a lane whose path has ended waits for the warp's longest one, sample after sample
// (ncu: 16.8 of 32 lanes active); and the sky colour of a path that escapes (~300 instructions of LUT arithmetic) is
// evaluated inside the bounce loop, i.e. in almost every iteration for the few lanes that escaped in it.
// BATCHED = true (default, option "spheres_batched"): the same arithmetic per lane in the same order -- samples one after the
// other, bounces one after the other, the running sum added to in sample order, the RNG state handed on -- but the warp
// steps a small state machine instead of the nested loops: a lane whose path ends starts its NEXT sample at once, so lanes
// only idle at the very end of the pixel (the sum of 8 path lengths varies far less than one of them); and a lane whose
// path escaped waits with (throughput, direction) until SKY_MIN lanes want the sky colour (or no lane can bounce), so that
// the LUT code runs with a dozen lanes instead of two; and a pixel that shows the sky evaluates its colour once, not once per
// sample (same direction, same value).  Bit-identical image and ray count (test).
#ifndef SPHERES_SKY_MIN
#define SPHERES_SKY_MIN 8   // measured 1 / 3 / 5 / 8 / 12 / 20 / 28: 0.303 / 0.289 / 0.277 / 0.270 / 0.277 / 0.342 / 0.409 ms per frame
#endif
template <bool BATCHED>
__global__ void __launch_bounds__(128)
k_spheres_secondary(SecondaryParams P, Spheres sp, Partition part, mrt_atmosphere_params A, SkyLuts luts,
                    const uchar4* __restrict__ bn, const uint32_t* __restrict__ vis, const uint16_t* __restrict__ depth,
                    const uint16_t* __restrict__ normal, uint16_t* __restrict__ color16, float4* __restrict__ accum,
                    int accumulate, unsigned long long* __restrict__ ray_counter) {
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, lr = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long rays = 0;
    // the lanes of this warp that own a pixel: the votes of the BATCHED state machine are taken among them
    const unsigned warp_mask = __ballot_sync(0xFFFFFFFFu, x < P.W && lr < P.local_rows);
    if (x < P.W && lr < P.local_rows) {
        uint32_t y = partition_local_to_y(part, lr);
        size_t p = (size_t)lr * P.W + x;
        float pitchx = 1.0f / (float)P.W, pitchy = 1.0f / (float)P.H;
        float u = ((float)x + 0.5f) * pitchx;
        float v = ((float)y + 0.5f) * pitchy;
        v = 1.0f - v;
        // reconstruct the primary hit from the fp16 G-buffer (:114-123)
        uint32_t pid = vis[p];
        uint2 npk = reinterpret_cast<const uint2*>(normal)[p];
        float3 pn = f3(f16_bits_to_f32((uint16_t)(npk.x & 0xFFFF)), f16_bits_to_f32((uint16_t)(npk.x >> 16)),
                       f16_bits_to_f32((uint16_t)(npk.y & 0xFFFF)));
        float dep = f16_bits_to_f32(depth[p]);
        float4 vp = mat_vec(P.invProj, u * 2.0f - 1.0f, v * 2.0f - 1.0f, dep, 1.0f);
        vp.x /= vp.w; vp.y /= vp.w; vp.z /= vp.w;
        float4 wp = mat_vec(P.invView, vp.x, vp.y, vp.z, 1.0f);
        float3 ppos = f3(wp.x, wp.y, wp.z);
        uint32_t rng = (P.frameCounter << 1u) | 1u;
        float2 rot = blue_noise_rotation(bn, P.bnW, P.bnH, x, y);

        float3 color = f3s(0.0f);
        if (BATCHED) {
            // per-lane path state; sky_wait: the path escaped along hn with throughput thr, its sky colour is still to be added
            uint32_t s = 0, i = 0, hid = pid;
            float3 thr = f3s(1.0f), hpos = ppos, hn = pn;
            bool sky_wait = false;
            // A pixel that shows the sky asks for the same sky colour (direction pn, throughput 1) once per sample: the
            // first answer is kept and the others add it again -- same value, same sum
            bool sky_known = false;
            float3 sky_pn = f3s(0.0f);
            // vertex 0 of sample s is the primary hit from the G-buffer: no ray, no random numbers
            auto begin_samples = [&]() {
                while (s < P.spp) {
                    hid = pid; hpos = ppos; hn = pn;
                    thr = f3s(1.0f);
                    if (hid == MRT_MISS_ID) {
                        if (sky_known) { color = color + thr * sky_pn; s++; continue; }
                        sky_wait = true;
                        return;
                    }
                    thr = thr * f3(sp.s[hid].albedo[0], sp.s[hid].albedo[1], sp.s[hid].albedo[2]);
                    i = 1;
                    if (i < P.bounces + 1u) return;
                    s++;  // no bounces asked for: the sample adds nothing
                }
            };
            begin_samples();
            for (;;) {
                const bool want_sky = s < P.spp && sky_wait, want_bounce = s < P.spp && !sky_wait;
                const unsigned sky_mask = __ballot_sync(warp_mask, want_sky), bounce_mask = __ballot_sync(warp_mask, want_bounce);
                if (!(sky_mask | bounce_mask)) break;
                if (sky_mask && (bounce_mask == 0u || __popc(sky_mask) >= SPHERES_SKY_MIN)) {
                    if (want_sky) {
                        const float3 sky = sky_color(A, luts, P.cameraPos, hn);
                        if (pid == MRT_MISS_ID) { sky_pn = sky; sky_known = true; }
                        color = color + thr * sky;
                        sky_wait = false;
                        s++;
                        begin_samples();
                    }
                } else if (want_bounce) {
                    float3 ro, rd;
                    lambert_bounce(hpos, hn, rng, rot.x, rot.y, ro, rd);
                    rays++;
                    hid = MRT_MISS_ID;
                    float ht = -1.0f;
                    for (uint32_t k = 0; k < sp.n; k++) {
                        float t = ray_sphere(ro, rd, sp.s[k]);
                        if (t >= 0.0f && (t < ht || ht < 0.0f)) { hid = k; ht = t; }
                    }
                    if (hid != MRT_MISS_ID) {
                        hpos = ro + rd * ht;
                        hn = normalize3(hpos - f3(sp.s[hid].center[0], sp.s[hid].center[1], sp.s[hid].center[2]));
                        thr = thr * f3(sp.s[hid].albedo[0], sp.s[hid].albedo[1], sp.s[hid].albedo[2]);
                        i++;
                        if (i == P.bounces + 1u) {  // the last bounce hit a sphere: the sample adds nothing (contrib = 0)
                            s++;
                            begin_samples();
                        }
                    } else {
                        hn = rd;
                        sky_wait = true;
                    }
                }
            }
        } else
        for (uint32_t s = 0; s < P.spp; s++) {
            float3 thr = f3s(1.0f);
            uint32_t hid = pid;
            float3 hpos = ppos, hn = pn;
            float3 contrib = f3s(0.0f);
            for (uint32_t i = 0; i < P.bounces + 1u; i++) {
                if (i > 0) {
                    float3 ro, rd;
                    lambert_bounce(hpos, hn, rng, rot.x, rot.y, ro, rd);
                    rays++;
                    hid = MRT_MISS_ID;
                    float ht = -1.0f;
                    for (uint32_t k = 0; k < sp.n; k++) {
                        float t = ray_sphere(ro, rd, sp.s[k]);
                        if (t >= 0.0f && (t < ht || ht < 0.0f)) { hid = k; ht = t; }
                    }
                    if (hid != MRT_MISS_ID) {
                        hpos = ro + rd * ht;
                        hn = normalize3(hpos - f3(sp.s[hid].center[0], sp.s[hid].center[1], sp.s[hid].center[2]));
                    } else {
                        hn = rd;
                    }
                }
                if (hid != MRT_MISS_ID) {
                    thr = thr * f3(sp.s[hid].albedo[0], sp.s[hid].albedo[1], sp.s[hid].albedo[2]);
                } else {
                    contrib = thr * sky_color(A, luts, P.cameraPos, hn);
                    break;
                }
            }
            color = color + contrib;
        }
        // progressive sum (row n7) before the reference's own average (:133)
        float4 a = accumulate ? accum[p] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        accum[p] = make_float4(a.x + color.x, a.y + color.y, a.z + color.z, a.w + (float)P.spp);
        color = color / (float)P.spp;
        uint2 pk;
        pk.x = (uint32_t)f32_to_f16_bits(color.x) | ((uint32_t)f32_to_f16_bits(color.y) << 16);
        pk.y = (uint32_t)f32_to_f16_bits(color.z) | (0x3C00u << 16);
        reinterpret_cast<uint2*>(color16)[p] = pk;
    }
    // one atomic per warp for the ray statistics
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(0xFFFFFFFFu, rays, off);
    if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == 0 && rays) atomicAdd(ray_counter, rays);
}

// option "spheres_batched" 2 (A/B: 0.396 vs 0.271 ms per frame -- 80 registers, and every pixel pays a thin refill: kept as an
// option, off): the state machine of BATCHED above, plus the traversal kernel's refill -- a
// persistent grid whose lanes take the NEXT PIXEL (in 16x2-tile order, one global atomic per tile) as soon as six of a
// warp's lanes have finished theirs.  What BATCHED leaves idle is the end of a pixel: the sum of eight path lengths still
// varies by +-17 % across a warp, and warps over the sky / ground boundary mix cheap and expensive pixels.  Same per-pixel
// arithmetic in the same order; pixels are independent, so the image does not depend on which lane renders which.
#ifndef SPHERES_REFILL_MIN
#define SPHERES_REFILL_MIN 6
#endif
__global__ void __launch_bounds__(128)
k_spheres_secondary_persistent(SecondaryParams P, Spheres sp, Partition part, mrt_atmosphere_params A, SkyLuts luts,
                               const uchar4* __restrict__ bn, const uint32_t* __restrict__ vis, const uint16_t* __restrict__ depth,
                               const uint16_t* __restrict__ normal, uint16_t* __restrict__ color16, float4* __restrict__ accum,
                               int accumulate, unsigned long long* __restrict__ ray_counter, uint32_t* __restrict__ work_counter) {
    const unsigned FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    const uint32_t tiles_x = (P.W + 15u) / 16u, tiles_y = (P.local_rows + 1u) / 2u, total = tiles_x * tiles_y * 32u;
    const float pitchx = 1.0f / (float)P.W, pitchy = 1.0f / (float)P.H;
    unsigned long long rays = 0;
    // the pixel this lane renders and its path state (see k_spheres_secondary<true>)
    bool have_pixel = false, sky_wait = false, sky_known = false;
    size_t p = 0;
    uint32_t pid = MRT_MISS_ID, rng = 0, s = 0, i = 0, hid = MRT_MISS_ID;
    float3 pn = f3s(0.0f), ppos = f3s(0.0f), color = f3s(0.0f), thr = f3s(1.0f), hpos = f3s(0.0f), hn = f3s(0.0f), sky_pn = f3s(0.0f);
    float2 rot = make_float2(0.0f, 0.0f);
    auto begin_samples = [&]() {
        while (s < P.spp) {
            hid = pid; hpos = ppos; hn = pn;
            thr = f3s(1.0f);
            if (hid == MRT_MISS_ID) {
                if (sky_known) { color = color + thr * sky_pn; s++; continue; }
                sky_wait = true;
                return;
            }
            thr = thr * f3(sp.s[hid].albedo[0], sp.s[hid].albedo[1], sp.s[hid].albedo[2]);
            i = 1;
            if (i < P.bounces + 1u) return;
            s++;
        }
    };
    uint32_t pool_next = 0, pool_end = 0;
    bool exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !have_pixel);
        if (idle && !exhausted && (idle == FULL || __popc(idle) >= SPHERES_REFILL_MIN)) {
            if (pool_next >= pool_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, 32u);
                base = __shfl_sync(FULL, base, 0);
                pool_next = base;
                pool_end = min(base + 32u, total);
                if (base >= total) { exhausted = true; pool_next = pool_end = 0; }
            }
            if (!exhausted) {
                const uint32_t mine = pool_next + __popc(idle & lt_mask);
                if (!have_pixel && mine < pool_end) {
                    const uint32_t tile = mine >> 5, j = mine & 31u;
                    const uint32_t x = (tile % tiles_x) * 16u + (j & 15u), lr = (tile / tiles_x) * 2u + (j >> 4);
                    if (x < P.W && lr < P.local_rows) {
                        const uint32_t y = partition_local_to_y(part, lr);
                        p = (size_t)lr * P.W + x;
                        float u = ((float)x + 0.5f) * pitchx;
                        float v = ((float)y + 0.5f) * pitchy;
                        v = 1.0f - v;
                        // reconstruct the primary hit from the fp16 G-buffer (:114-123)
                        pid = vis[p];
                        const uint2 npk = reinterpret_cast<const uint2*>(normal)[p];
                        pn = f3(f16_bits_to_f32((uint16_t)(npk.x & 0xFFFF)), f16_bits_to_f32((uint16_t)(npk.x >> 16)),
                                f16_bits_to_f32((uint16_t)(npk.y & 0xFFFF)));
                        const float dep = f16_bits_to_f32(depth[p]);
                        float4 vp = mat_vec(P.invProj, u * 2.0f - 1.0f, v * 2.0f - 1.0f, dep, 1.0f);
                        vp.x /= vp.w; vp.y /= vp.w; vp.z /= vp.w;
                        const float4 wp = mat_vec(P.invView, vp.x, vp.y, vp.z, 1.0f);
                        ppos = f3(wp.x, wp.y, wp.z);
                        rng = (P.frameCounter << 1u) | 1u;
                        rot = blue_noise_rotation(bn, P.bnW, P.bnH, x, y);
                        color = f3s(0.0f);
                        s = 0;
                        sky_wait = false;
                        sky_known = false;
                        have_pixel = true;
                        begin_samples();
                    }
                }
                pool_next = min(pool_next + (uint32_t)__popc(idle), pool_end);
            }
        }
        if (exhausted && __ballot_sync(FULL, have_pixel) == 0u) break;
        if (have_pixel && s >= P.spp) {
            // progressive sum (row n7) before the reference's own average (:133)
            const float4 a = accumulate ? accum[p] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            accum[p] = make_float4(a.x + color.x, a.y + color.y, a.z + color.z, a.w + (float)P.spp);
            const float3 avg = color / (float)P.spp;
            uint2 pk;
            pk.x = (uint32_t)f32_to_f16_bits(avg.x) | ((uint32_t)f32_to_f16_bits(avg.y) << 16);
            pk.y = (uint32_t)f32_to_f16_bits(avg.z) | (0x3C00u << 16);
            reinterpret_cast<uint2*>(color16)[p] = pk;
            have_pixel = false;
        }
        const bool want_sky = have_pixel && sky_wait, want_bounce = have_pixel && !sky_wait;
        const unsigned sky_mask = __ballot_sync(FULL, want_sky), bounce_mask = __ballot_sync(FULL, want_bounce);
        if (sky_mask && (bounce_mask == 0u || __popc(sky_mask) >= SPHERES_SKY_MIN)) {
            if (want_sky) {
                const float3 sky = sky_color(A, luts, P.cameraPos, hn);
                if (pid == MRT_MISS_ID) { sky_pn = sky; sky_known = true; }
                color = color + thr * sky;
                sky_wait = false;
                s++;
                begin_samples();
            }
        } else if (want_bounce) {
            float3 ro, rd;
            lambert_bounce(hpos, hn, rng, rot.x, rot.y, ro, rd);
            rays++;
            hid = MRT_MISS_ID;
            float ht = -1.0f;
            for (uint32_t k = 0; k < sp.n; k++) {
                float t = ray_sphere(ro, rd, sp.s[k]);
                if (t >= 0.0f && (t < ht || ht < 0.0f)) { hid = k; ht = t; }
            }
            if (hid != MRT_MISS_ID) {
                hpos = ro + rd * ht;
                hn = normalize3(hpos - f3(sp.s[hid].center[0], sp.s[hid].center[1], sp.s[hid].center[2]));
                thr = thr * f3(sp.s[hid].albedo[0], sp.s[hid].albedo[1], sp.s[hid].albedo[2]);
                i++;
                if (i == P.bounces + 1u) {  // the last bounce hit a sphere: the sample adds nothing (contrib = 0)
                    s++;
                    begin_samples();
                }
            } else {
                hn = rd;
                sky_wait = true;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(FULL, rays, off);
    if (lane == 0 && rays) atomicAdd(ray_counter, rays);
}

}  // namespace

int spheres_primary(mrt_context* ctx) {
    RayGen g;
    memcpy(&g.invView, &ctx->pc.invView, sizeof(Mat4));
    memcpy(&g.invProj, &ctx->pc.invProjection, sizeof(Mat4));
    g.W = ctx->W;
    g.H = ctx->H;
    Mat4 P, V, Vp;
    memcpy(&P, &ctx->pc.projection, sizeof(Mat4));
    memcpy(&V, &ctx->pc.view, sizeof(Mat4));
    memcpy(&Vp, &ctx->pc.prevView, sizeof(Mat4));
    Mat4 PV = mat_mul(P, V), PVprev = mat_mul(P, Vp);
    dim3 b(32, 8), grid(div_up(ctx->W, 32), div_up(ctx->local_rows, 8));
    k_spheres_primary<<<grid, b, 0, ctx->stream>>>(g, PV, PVprev, ctx->spheres, ctx->part, ctx->local_rows,
                                                   ctx->visibility.p, ctx->depth.p, ctx->normal.p, ctx->motion.p);
    MRT_LAUNCHED(ctx);
    ctx->stats.primary_rays = ctx->npix;
    return mrt_check_cuda(ctx, cudaGetLastError(), "spheres_primary");
}

int spheres_secondary(mrt_context* ctx, const mrt_secondary_constants* c, uint32_t spp, uint32_t bounces, uint32_t flags) {
    SecondaryParams P;
    memcpy(&P.invView, &c->invView, sizeof(Mat4));
    memcpy(&P.invProj, &c->invProjection, sizeof(Mat4));
    P.cameraPos = f3(c->cameraPos[0], c->cameraPos[1], c->cameraPos[2]);
    P.frameCounter = c->frameCounter;
    P.W = ctx->W; P.H = ctx->H; P.local_rows = ctx->local_rows; P.spp = spp; P.bounces = bounces;
    P.bnW = ctx->bnW; P.bnH = ctx->bnH;
    SkyLuts luts{ctx->trans_f.p, nullptr, ctx->view_f.p};
    MRT_TRY(dev_reserve(ctx, ctx->visit_counters, 8));
    MRT_CUDA(ctx, cudaMemsetAsync(ctx->visit_counters.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
    // 16x8 tiles: warps cover 16x2 pixel footprints, which keeps the divergent bounce loops of
    // neighbouring pixels (same sphere, similar path length) in one warp.
    dim3 b(16, 8), grid(div_up(ctx->W, 16), div_up(ctx->local_rows, 8));
    const int acc = (flags & MRT_SECONDARY_ACCUMULATE) && ctx->have_accum ? 1 : 0;
    if (ctx->opt_spheres_batched >= 2) {
        // persistent grid: as many CTAs as are resident at once (or fewer when the image is small)
        if (ctx->spheres_grid <= 0) {
            int sms = 148, per_sm = 4;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spheres_secondary_persistent, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
            ctx->spheres_grid = sms * per_sm;
        }
        const unsigned warps = div_up(ctx->W, 16u) * div_up(ctx->local_rows, 2u);
        const unsigned grid1 = max(1u, min((unsigned)ctx->spheres_grid, div_up(warps, 4u)));
        k_spheres_secondary_persistent<<<grid1, 128, 0, ctx->stream>>>(P, ctx->spheres, ctx->part, ctx->atmo, luts, ctx->bn, ctx->visibility.p,
                                                                       ctx->depth.p, ctx->normal.p, ctx->color16.p, ctx->accum.p, acc,
                                                                       ctx->visit_counters.p + 4,
                                                                       reinterpret_cast<uint32_t*>(ctx->visit_counters.p + 5));
        MRT_LAUNCHED(ctx);
        return mrt_check_cuda(ctx, cudaGetLastError(), "spheres_secondary");
    }
    auto* const kernel = ctx->opt_spheres_batched ? k_spheres_secondary<true> : k_spheres_secondary<false>;
    kernel<<<grid, b, 0, ctx->stream>>>(P, ctx->spheres, ctx->part, ctx->atmo, luts, ctx->bn, ctx->visibility.p, ctx->depth.p,
                                        ctx->normal.p, ctx->color16.p, ctx->accum.p, acc, ctx->visit_counters.p + 4);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "spheres_secondary");
}
