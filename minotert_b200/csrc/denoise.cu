// denoise.cu -- Denoiser::bilateral (src/gfx/modules/denoiser.ixx:36-97): the bilateral filter of
// src/gpu/denoise/bilateral.comp:23-76 over the path tracer's RGBA16F colour image, guided by the G-buffer's
// R16F depth and RGBA16F normal; output RGBA8 unorm (denoiser.ixx:56), which the tonemapper then reads.
// the stage right after the hot path in Renderer_impl::draw (renderer.ixx:61).
//
// Shape of the work: every pixel visits the taps of a disc of radius round(kSigma*sigma) (325 taps at the
// defaults 5 / 2): d.x runs over integers, d.y starts at -sqrt(r^2 - d.x^2) and advances by 1, so most columns
// sample BETWEEN rows and the LinearClamp sampler really interpolates.  Per tap: 3 texture reads, a division, an
// exp and 5 multiply-adds -- an issue-bound stencil (about 25 000 instructions per pixel), not a bandwidth-bound
// one: the 18 B/px of input are read once per CTA tile into shared memory.
//
// Kernel: one thread per pixel, 32x8-pixel CTA, shared tile with a halo of radius+1 texels holding
// {r, g, b, depth, nx, ny, nz} as the stored halfs (one LDS.128 per texel); image-edge clamping is applied when
// the tile is filled, so the tap loop needs none.  The tap list (d/size and the spatial Gaussian, which depend
// only on the push constants and the image size) is evaluated once on the host with the same fp32 operations as
// the shader's loops and read through uniform loads.  Taps are visited in the shader's order.  Sampler rule
// (the oracle's, oracle/minote_oracle.c:texn_bilinear): bilinear weights carry 8 fractional bits and zero-weight
// texels are not read.  d.x is integral, so in x every tap is a texel centre (the fp32 residue of uv + d/size is
// <= 2^-12 texel and rounds to weight 0): one column, no x-lerp; rows are mixed with the k/256 weight in fp32,
// operation for operation as the oracle does.  The one deliberate deviation: exp and the depth division use the
// SFU approximations (ex2.approx, rcp.approx), as GLSL exp() and '/' do on the reference's GPU path; the bar in
// tests/test_gpu_denoise.py (RGBA8: <= 1 code value on >= 99.9 % of pixels) absorbs it.
#include <math.h>

#include <vector>

#include "context.cuh"
#include "shading.cuh"

namespace {

#ifndef DN_ROWS
#define DN_ROWS 8     // CTA = 32 x DN_ROWS pixels
#endif
#ifndef DN_UNROLL
#define DN_UNROLL 2   // taps per loop trip
#endif
constexpr int DN_BX = 32, DN_BY = DN_ROWS, DN_UNROLL_N = DN_UNROLL;

struct DenoiseArgs {
    uint32_t W, H;
    float sizeX, sizeY;
    float invThresholdSqx2, invThresholdSqrt2PI;
    float nearPlane;
    uint32_t frameCounter;
    int halo;   // radius + 1
    int ntaps;
};

struct Texel { float r, g, b, z, nx, ny, nz; };

MRT_D Texel unpack_texel(uint4 t) {
    float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    float2 bz = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    float2 nxy = __half22float2(*reinterpret_cast<const __half2*>(&t.z));
    float2 nzw = __half22float2(*reinterpret_cast<const __half2*>(&t.w));
    return Texel{rg.x, rg.y, bz.x, bz.y, nxy.x, nxy.y, nzw.x};
}

// bilateral.comp:23-76
__global__ void __launch_bounds__(DN_BX* DN_BY)
    k_denoise_bilateral(DenoiseArgs A, const uint2* __restrict__ color16, const uint16_t* __restrict__ depth16,
                        const uint2* __restrict__ normal16, const float4* __restrict__ taps, uchar4* __restrict__ out) {
    extern __shared__ uint4 tile[];
    const int TW = DN_BX + 2 * A.halo, TH = DN_BY + 2 * A.halo;
    const int x0 = (int)blockIdx.x * DN_BX - A.halo, y0 = (int)blockIdx.y * DN_BY - A.halo;
    const int tid = threadIdx.y * DN_BX + threadIdx.x;
    for (int i = tid; i < TW * TH; i += DN_BX * DN_BY) {
        int ty = i / TW, tx = i - ty * TW;
        int gx = min(max(x0 + tx, 0), (int)A.W - 1), gy = min(max(y0 + ty, 0), (int)A.H - 1);  // ClampToEdge
        size_t g = (size_t)gy * A.W + gx;
        uint2 c = __ldg(&color16[g]);
        uint2 n = __ldg(&normal16[g]);
        uint32_t d = __ldg(&depth16[g]);
        tile[i] = make_uint4(c.x, (c.y & 0xFFFFu) | (d << 16), n.x, n.y);
    }
    __syncthreads();
    const uint32_t px = blockIdx.x * DN_BX + threadIdx.x, py = blockIdx.y * DN_BY + threadIdx.y;
    if (px >= A.W || py >= A.H) return;

    const float uvx = ((float)px + 0.5f) / A.sizeX, uvy = ((float)py + 0.5f) / A.sizeY;
    const Texel centre = unpack_texel(tile[((int)py - y0) * TW + ((int)px - x0)]);
    float3 filtered;
    if (centre.z < 0.0f) {  // bilateral.comp:36
        filtered = f3(centre.r, centre.g, centre.b);
    } else {
        const float centreDist = __fdividef(A.nearPlane, centre.z);
        float zBuff = 0.0f;
        float3 aBuff = f3s(0.0f);
#pragma unroll DN_UNROLL_N
        for (int i = 0; i < A.ntaps; i++) {
            const float4 tp = __ldg(&taps[i]);  // d.x/size.x, d.y/size.y, blurFactor, d.y integral?
            const float x = (uvx + tp.x) * A.sizeX - 0.5f;
            const float y = (uvy + tp.y) * A.sizeY - 0.5f;
            const int col = __float2int_rn(x) - x0;
            Texel w;
            if (tp.w != 0.0f) {  // d.y integral (uniform across the CTA): the tap is a texel centre in y too
                w = unpack_texel(tile[(__float2int_rn(y) - y0) * TW + col]);
            } else {
                const float fy0 = floorf(y);
                const float fy = rintf((y - fy0) * 256.0f) * (1.0f / 256.0f), gy = 1.0f - fy;  // k/256 row weight
                const int row = (int)fy0 - y0;
                const Texel a = unpack_texel(tile[row * TW + col]);
                const Texel b = unpack_texel(tile[(row + 1) * TW + col]);
                if (fy == 0.0f) {  // zero-weight texels are not read (they may hold inf): rare, per-lane
                    w = a;
                } else if (gy == 0.0f) {
                    w = b;
                } else {
                    w.r = a.r * gy + b.r * fy;
                    w.g = a.g * gy + b.g * fy;
                    w.b = a.b * gy + b.b * fy;
                    w.z = a.z * gy + b.z * fy;
                    w.nx = a.nx * gy + b.nx * fy;
                    w.ny = a.ny * gy + b.ny * fy;
                    w.nz = a.nz * gy + b.nz * fy;
                }
            }
            float dZ = __fdividef(A.nearPlane, w.z) - centreDist;
            dZ *= 100.0f;
            const float dN = w.nx * centre.nx + w.ny * centre.ny + w.nz * centre.nz;
            const float deltaFactor =
                __expf(clampf(dN - dZ * dZ, 0.0f, 1.0f) * A.invThresholdSqx2) * A.invThresholdSqrt2PI * tp.z;
            zBuff += deltaFactor;
            aBuff.x += deltaFactor * w.r;
            aBuff.y += deltaFactor * w.g;
            aBuff.z += deltaFactor * w.b;
        }
        filtered = f3(aBuff.x / zBuff, aBuff.y / zBuff, aBuff.z / zBuff);
    }
    // bilateral.comp:71-73: one PCG draw per pixel, the same value on r, g and b
    uint32_t seed = px * 709u + py * 1153u + A.frameCounter * 1361u;
    const float noise = (float)(pcg(seed) & 0xFFFFFFu) / 16777216.0f * 0.005f;
    // alpha: sum(w * 1) / sum(w) = 1 up to rounding -> 255
    __stcs(&out[(size_t)py * A.W + px], make_uchar4((unsigned char)unorm8(filtered.x + noise), (unsigned char)unorm8(filtered.y + noise),
                                                    (unsigned char)unorm8(filtered.z + noise), 255));
}

}  // namespace

// The loops of smartDeNoise (bilateral.comp:43-47) in the shader's own fp32 arithmetic: which taps exist, their
// d/size offsets and the spatial Gaussian depend only on (sigma, kSigma, image size).
static void build_taps(float sigma, float kSigma, float sizeX, float sizeY, std::vector<float4>& taps) {
    const float INV_PI = 0.31830988618379067153776752674503f;
    const float radius = roundf(kSigma * sigma);
    const float radQ = radius * radius;
    const float invSigmaQx2 = .5f / (sigma * sigma);
    const float invSigmaQx2PI = INV_PI * invSigmaQx2;
    taps.clear();
    for (float dx = -radius; dx <= radius; dx++) {
        const float pt = sqrtf(radQ - dx * dx);
        for (float dy = -pt; dy <= pt; dy++) {
            const float blurFactor = expf(-(dx * dx + dy * dy) * invSigmaQx2) * invSigmaQx2PI;
            taps.push_back(make_float4(dx / sizeX, dy / sizeY, blurFactor, dy == rintf(dy) ? 1.0f : 0.0f));
        }
    }
}

int denoise_bilateral(mrt_context* ctx, float sigma, float kSigma, float threshold, float nearPlane, uint32_t frameCounter) {
    const float radius = roundf(kSigma * sigma);
    if (!(sigma > 0.0f) || !(threshold > 0.0f) || !(radius >= 0.0f) || radius > 32.0f)
        return mrt_fail(ctx, MRT_ERR_INVALID, "bilateral denoiser: sigma %g, kSigma %g, threshold %g (need sigma, threshold > 0 and round(kSigma*sigma) <= 32)",
                        sigma, kSigma, threshold);
    const uint32_t W = ctx->W, H = ctx->local_rows;
    const size_t n = ctx->npix;
    MRT_TRY(dev_reserve(ctx, ctx->denoised, n));
    if (n == 0) return MRT_OK;

    if (ctx->dn_key_sigma != sigma || ctx->dn_key_ksigma != kSigma || ctx->dn_key_w != W || ctx->dn_key_h != H) {
        std::vector<float4> taps;
        build_taps(sigma, kSigma, (float)W, (float)H, taps);
        MRT_TRY(dev_reserve(ctx, ctx->dn_taps, taps.size()));
        // pageable source: the copy is staged before the call returns
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->dn_taps.p, taps.data(), taps.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dn_ntaps = (int)taps.size();
        ctx->dn_key_sigma = sigma; ctx->dn_key_ksigma = kSigma; ctx->dn_key_w = W; ctx->dn_key_h = H;
    }

    DenoiseArgs A;
    A.W = W; A.H = H;
    A.sizeX = (float)W; A.sizeY = (float)H;
    A.invThresholdSqx2 = .5f / (threshold * threshold);
    A.invThresholdSqrt2PI = 0.39894228040143267793994605993439f / threshold;
    A.nearPlane = nearPlane;
    A.frameCounter = frameCounter;
    A.halo = (int)radius + 1;
    A.ntaps = ctx->dn_ntaps;
    const size_t smem = (size_t)(DN_BX + 2 * A.halo) * (DN_BY + 2 * A.halo) * sizeof(uint4);
    if (smem > 48 * 1024)
        MRT_CUDA(ctx, cudaFuncSetAttribute(k_denoise_bilateral, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(div_up(W, DN_BX), div_up(H, DN_BY)), block(DN_BX, DN_BY);
    k_denoise_bilateral<<<grid, block, smem, ctx->stream>>>(A, reinterpret_cast<const uint2*>(ctx->color16.p), ctx->depth.p,
                                                            reinterpret_cast<const uint2*>(ctx->normal.p), ctx->dn_taps.p,
                                                            ctx->denoised.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "denoise_bilateral");
}
