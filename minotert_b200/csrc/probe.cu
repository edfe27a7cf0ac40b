// probe.cu -- device-side unit probes (tests / tools): the per-path shading functions of the hot path evaluated on
// the GPU for caller-supplied inputs, so that they can be compared call by call with the reference's own code
// (oracle/_ref) instead of only through whole images.
//   mrt_eval_sky_color     : skyColor() of secondaryRays.comp:36-58 (sky.cuh) for n directions
//   mrt_eval_bounce_stream : the per-pixel sample stream of secondaryRays.comp:60-72,124-125 (shading.cuh): PCG state
//                            seeded with (frameCounter << 1) | 1, blue-noise rotation of pixel (x, y), n consecutive
//                            Lambert bounces off (pos, normal)
#include "shading.cuh"

namespace {

__global__ void __launch_bounds__(128) k_eval_sky(mrt_atmosphere_params A, SkyLuts luts, float3 cameraPos,
                                                  const float* __restrict__ dirs, uint32_t n, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = sky_color(A, luts, cameraPos, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

__global__ void k_eval_bounces(const uchar4* bn, uint32_t bnW, uint32_t bnH, uint32_t seed, uint32_t x, uint32_t y,
                               float3 pos, float3 nrm, uint32_t n, float* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t rng = seed;
    const float2 rot = blue_noise_rotation(bn, bnW, bnH, x, y);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t probe = rng;  // the two rotated randoms the bounce is about to draw, reported separately
        const float r0 = rotated_random(probe, rot.x), r1 = rotated_random(probe, rot.y);
        float3 ro, rd;
        lambert_bounce(pos, nrm, rng, rot.x, rot.y, ro, rd);
        float* o = out + 9 * (size_t)i;
        o[0] = r0; o[1] = r1; o[2] = ro.x; o[3] = ro.y; o[4] = ro.z; o[5] = rd.x; o[6] = rd.y; o[7] = rd.z;
        o[8] = __uint_as_float(rng);
    }
}

}  // namespace

int probe_sky_color(mrt_context* ctx, const float cameraPos[3], const float* dirs, uint32_t n, float* out) {
    if (n == 0) return MRT_OK;
    MRT_TRY(dev_reserve(ctx, ctx->query_d, 3 * (size_t)n));
    MRT_TRY(dev_reserve(ctx, ctx->query_o, 3 * (size_t)n));
    MRT_CUDA(ctx, cudaMemcpyAsync(ctx->query_d.p, dirs, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    SkyLuts luts{ctx->trans_f.p, nullptr, ctx->view_f.p};
    k_eval_sky<<<div_up(n, 128), 128, 0, ctx->stream>>>(ctx->atmo, luts, f3(cameraPos[0], cameraPos[1], cameraPos[2]), ctx->query_d.p,
                                                       n, ctx->query_o.p);
    MRT_LAUNCHED(ctx);
    MRT_CUDA(ctx, cudaGetLastError());
    MRT_CUDA(ctx, cudaMemcpyAsync(out, ctx->query_o.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

int probe_bounce_stream(mrt_context* ctx, uint32_t frameCounter, uint32_t x, uint32_t y, const float pos[3], const float normal[3],
                        uint32_t n, float* out9) {
    if (n == 0) return MRT_OK;
    MRT_TRY(dev_reserve(ctx, ctx->query_o, 9 * (size_t)n));
    k_eval_bounces<<<1, 32, 0, ctx->stream>>>(ctx->bn, ctx->bnW, ctx->bnH, (frameCounter << 1u) | 1u, x, y, f3(pos[0], pos[1], pos[2]),
                                             f3(normal[0], normal[1], normal[2]), n, ctx->query_o.p);
    MRT_LAUNCHED(ctx);
    MRT_CUDA(ctx, cudaGetLastError());
    MRT_CUDA(ctx, cudaMemcpyAsync(out9, ctx->query_o.p, sizeof(float) * 9 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}
