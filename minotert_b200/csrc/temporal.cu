// temporal.cu -- temporal accumulation by reprojection.
//
// The reference writes the screen-space motion of every primary hit between prevView and view into an RG16F image
// (src/gpu/primaryRay.comp:73-75, src/gfx/modules/pathtracer.ixx:63-69) and nothing reads it (renderer.ixx:61).
// This is the consumer: the frame's radiance (fp32 accumulator average, row n7) is blended into a history image
// fetched at the pixel's previous position, so a moving camera keeps its accumulated samples instead of
// restarting.  Contract = oracle/minote_oracle.h (orc_temporal_accumulate): previous position (x + 0.5 - m.x / 2,
// y + 0.5 + m.y / 2); bilinear over the 4 nearest history texels, taps outside the image or on another primitive
// (visibility id of the previous frame != this pixel's) dropped and the weights renormalised; exponential moving
// average with the history length capped at maxHistory; primary misses (the noise-free sky) pass through.
// Built with -fmad=false and written in the oracle's expression order: bit-identical results (tested).
//
// HBM-streaming: per pixel 16 (accumulator) + 4 (visibility) + 4 (motion) read, 16 + 4 + 4 written, plus the
// gathered history taps (4 x 24 B, neighbouring pixels share them through L1/L2): about 70 algorithmic B/px.
#include "context.cuh"

namespace {

__global__ void __launch_bounds__(256) k_temporal(uint32_t W, uint32_t H, const float4* __restrict__ accum,
                                                  const uint32_t* __restrict__ vis, const uint32_t* __restrict__ motion,
                                                  int have_history, const float4* __restrict__ hist, const float* __restrict__ hist_count,
                                                  const uint32_t* __restrict__ hist_vis, float maxHistory,
                                                  float4* __restrict__ out, float* __restrict__ out_count, uint32_t* __restrict__ out_vis) {
    const uint32_t x = blockIdx.x * 32u + (threadIdx.x & 31u), y = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t p = (size_t)y * W + x;
    const float4 a = __ldcs(&accum[p]);
    const float3 cur = a.w > 0.0f ? f3(a.x / a.w, a.y / a.w, a.z / a.w) : f3s(0.0f);
    const uint32_t id = vis[p];
    float3 res = cur;
    float count = 1.0f;
    if (have_history && id != MRT_MISS_ID) {
        const uint32_t m = motion[p];
        const float mx = f16_bits_to_f32((uint16_t)(m & 0xFFFFu)), my = f16_bits_to_f32((uint16_t)(m >> 16));
        const float gx = ((float)x + 0.5f - mx * 0.5f) - 0.5f, gy = ((float)y + 0.5f + my * 0.5f) - 0.5f;
        if (gx > -2.0f && gy > -2.0f && gx < (float)W + 1.0f && gy < (float)H + 1.0f) {
            const float fx0 = floorf(gx), fy0 = floorf(gy);
            const float wx = gx - fx0, wy = gy - fy0;
            const int x0 = (int)fx0, y0 = (int)fy0;
            float3 sum = f3s(0.0f);
            float nsum = 0.0f, wsum = 0.0f;
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int tx = x0 + i, ty = y0 + j;
                    if (tx < 0 || ty < 0 || tx >= (int)W || ty >= (int)H) continue;
                    const size_t q = (size_t)ty * W + (size_t)tx;
                    if (__ldg(&hist_vis[q]) != id) continue;
                    const float wt = (i ? wx : 1.0f - wx) * (j ? wy : 1.0f - wy);
                    const float4 hq = __ldg(&hist[q]);
                    sum = f3(sum.x + wt * hq.x, sum.y + wt * hq.y, sum.z + wt * hq.z);
                    nsum = nsum + wt * __ldg(&hist_count[q]);
                    wsum = wsum + wt;
                }
            if (wsum > 0.00390625f) {
                float n = nsum / wsum;
                n = n < maxHistory ? n : maxHistory;
                const float k = 1.0f / (n + 1.0f);
                const float3 hc = f3(sum.x / wsum, sum.y / wsum, sum.z / wsum);
                res = f3(hc.x + (cur.x - hc.x) * k, hc.y + (cur.y - hc.y) * k, hc.z + (cur.z - hc.z) * k);
                count = n + 1.0f;
            }
        }
    }
    out[p] = make_float4(res.x, res.y, res.z, 1.0f);
    out_count[p] = count;
    out_vis[p] = id;
}

}  // namespace

int temporal_accumulate(mrt_context* ctx, float maxHistory, bool reset) {
    const size_t n = ctx->npix;
    const bool same_size = ctx->tp_w == ctx->W && ctx->tp_h == ctx->H;
    for (int k = 0; k < 2; k++) {
        MRT_TRY(dev_reserve(ctx, ctx->tp_rgba[k], n));
        MRT_TRY(dev_reserve(ctx, ctx->tp_count[k], n));
        MRT_TRY(dev_reserve(ctx, ctx->tp_vis[k], n));
    }
    const int have_history = ctx->have_temporal && same_size && !reset ? 1 : 0;
    const int prev = ctx->tp_cur, cur = prev ^ 1;
    dim3 grid(div_up(ctx->W, 32), div_up(ctx->H, 8));
    k_temporal<<<grid, 256, 0, ctx->stream>>>(ctx->W, ctx->H, ctx->accum.p, ctx->visibility.p,
                                              reinterpret_cast<const uint32_t*>(ctx->motion.p), have_history, ctx->tp_rgba[prev].p,
                                              ctx->tp_count[prev].p, ctx->tp_vis[prev].p, maxHistory, ctx->tp_rgba[cur].p,
                                              ctx->tp_count[cur].p, ctx->tp_vis[cur].p);
    MRT_LAUNCHED(ctx);
    ctx->tp_cur = cur;
    ctx->tp_w = ctx->W;
    ctx->tp_h = ctx->H;
    return mrt_check_cuda(ctx, cudaGetLastError(), "temporal_accumulate");
}
