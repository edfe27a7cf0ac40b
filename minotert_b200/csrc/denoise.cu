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
// the tile is filled, so the tap loop needs none.  The tap list (which taps exist, d.y/size.y, the spatial
// Gaussian, the tile offset of d.x) depends only on the push constants and the image size: it is evaluated once on
// the host with the shader's own fp32 loops and read through uniform loads.  Taps are visited in the shader's order.
// Sampler rule (the oracle's, oracle/minote_oracle.c:texn_bilinear): bilinear weights carry 8 fractional bits and
// zero-weight texels are not read.  d.x is integral, so in x every tap is the texel centre px + d.x (the fp32
// residue of uv + d/size is < 2^-10 texel for images up to 4096 wide and rounds to weight 0): one column, no
// x-lerp, no per-tap x arithmetic; the same holds for the taps whose d.y is integral.  For the others the row
// position and its k/256 weight are computed per pixel with the oracle's operations, so both sides pick the same
// weight.  Deliberate deviations, all at the 1e-6 relative level and absorbed by the bar in
// tests/test_gpu_denoise.py (RGBA8: <= 1 code value on >= 99.9 % of pixels; measured: 8e-6 of the pixels differ,
// by 1): exp and the depth division run on the SFU (ex2.approx, rcp.approx), as GLSL exp() and '/' do on the
// reference's GPU path, and the row lerp and the accumulation are contracted to FMAs (this file is built with
// -fmad=false like the rest, the fmaf() calls are explicit).
#include <math.h>

#include <vector>

#include "context.cuh"
#include "shading.cuh"

namespace {

#ifndef DN_ROWS
#define DN_ROWS 8     // CTA = 32 x DN_ROWS pixels
#endif
#ifndef DN_UNROLL
#define DN_UNROLL 4   // taps per loop trip (1: 1.37 ms, 2: 1.34, 4: 1.29 at 1080p)
#endif
constexpr int DN_BX = 32, DN_BY = DN_ROWS, DN_UNROLL_N = DN_UNROLL;

struct DenoiseArgs {
    uint32_t W, H;
    float sizeX, sizeY;
    float invThresholdSqx2Log2e, invThresholdSqrt2PI;
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

// single-instruction SFU forms (MUFU.RCP / MUFU.EX2): arguments here are never fp32-denormal (depth comes from
// fp16, the exponent is >= 0), so the flush-to-zero variants lose nothing
MRT_D float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
MRT_D float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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

    const float uvy = ((float)py + 0.5f) / A.sizeY;
    const int tix = (int)px - x0, tiy = (int)py - y0;
    const int centre_idx = tiy * TW + tix;
    const Texel centre = unpack_texel(tile[centre_idx]);
    float3 filtered;
    if (centre.z < 0.0f) {  // bilateral.comp:36
        filtered = f3(centre.r, centre.g, centre.b);
    } else {
        const float centreDist = A.nearPlane * rcp_approx(centre.z);
        float zBuff = 0.0f;
        float3 aBuff = f3s(0.0f);
#pragma unroll DN_UNROLL_N
        for (int i = 0; i < A.ntaps; i++) {
            const float4 tp = __ldg(&taps[i]);  // d.y/size.y, blurFactor, tile offset (int), d.y integral? (int)
            const int off = __float_as_int(tp.z);
            Texel w;
            if (__float_as_int(tp.w) != 0) {  // d.x and d.y integral (uniform across the CTA): a texel centre
                w = unpack_texel(tile[centre_idx + off]);
            } else {
                // row position and its k/256 weight exactly as the oracle's sampler computes them
                const float y = (uvy + tp.x) * A.sizeY - 0.5f;
                const float fy0 = floorf(y);
                const float fy = rintf((y - fy0) * 256.0f) * (1.0f / 256.0f), gy = 1.0f - fy;
                const int idx = ((int)fy0 - y0) * TW + tix + off;
                const Texel a = unpack_texel(tile[idx]);
                const Texel b = unpack_texel(tile[idx + TW]);
                // zero-weight texels are not read: only the colour can hold inf (sun disc), depth and normal are finite
                const bool only_a = fy == 0.0f, only_b = gy == 0.0f;
                w.r = only_a ? a.r : (only_b ? b.r : fmaf(b.r, fy, a.r * gy));
                w.g = only_a ? a.g : (only_b ? b.g : fmaf(b.g, fy, a.g * gy));
                w.b = only_a ? a.b : (only_b ? b.b : fmaf(b.b, fy, a.b * gy));
                w.z = fmaf(b.z, fy, a.z * gy);
                w.nx = fmaf(b.nx, fy, a.nx * gy);
                w.ny = fmaf(b.ny, fy, a.ny * gy);
                w.nz = fmaf(b.nz, fy, a.nz * gy);
            }
            float dZ = A.nearPlane * rcp_approx(w.z) - centreDist;
            dZ *= 100.0f;
            const float dN = fmaf(w.nz, centre.nz, fmaf(w.ny, centre.ny, w.nx * centre.nx));
            // exp(c * invThresholdSqx2) = 2^(c * invThresholdSqx2 * log2 e)
            const float deltaFactor = ex2_approx(clampf(dN - dZ * dZ, 0.0f, 1.0f) * A.invThresholdSqx2Log2e) * A.invThresholdSqrt2PI * tp.y;
            zBuff += deltaFactor;
            aBuff.x = fmaf(deltaFactor, w.r, aBuff.x);
            aBuff.y = fmaf(deltaFactor, w.g, aBuff.y);
            aBuff.z = fmaf(deltaFactor, w.b, aBuff.z);
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
static void build_taps(float sigma, float kSigma, float sizeY, int tileW, std::vector<float4>& taps) {
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
            // In x the tap is the texel centre px + d.x: (uv.x + d.x/W) * W - 0.5 differs from it by the fp32 rounding of
            // uv.x and d.x/W, at most 2^-23 * W * 2 < 2^-10 texel for W <= 4096, which the 8-bit weight rounds to 0.
            // The same holds in y when d.y is integral; otherwise the kernel evaluates the row position per pixel.
            const bool whole = dy == rintf(dy);
            const int off = whole ? (int)dy * tileW + (int)dx : (int)dx;
            float4 t;
            t.x = dy / sizeY;
            t.y = blurFactor;
            memcpy(&t.z, &off, 4);
            const int flag = whole ? 1 : 0;
            memcpy(&t.w, &flag, 4);
            taps.push_back(t);
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
    if (W > 4096 || H > 4096)  // the integral taps are addressed as texel centres: exact up to this size (build_taps)
        return mrt_fail(ctx, MRT_ERR_INVALID, "bilateral denoiser: image %ux%u exceeds 4096x4096", W, H);
    MRT_TRY(dev_reserve(ctx, ctx->denoised, n));
    if (n == 0) return MRT_OK;

    if (ctx->dn_key_sigma != sigma || ctx->dn_key_ksigma != kSigma || ctx->dn_key_w != W || ctx->dn_key_h != H) {
        std::vector<float4> taps;
        build_taps(sigma, kSigma, (float)H, DN_BX + 2 * ((int)radius + 1), taps);
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
    A.invThresholdSqx2Log2e = .5f / (threshold * threshold) * 1.4426950408889634f;
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
