// tonemap.cu -- fused resolve (accumulator average) + tonemap + sRGB encode + RGBA8 store (rows a14, n7).
// Operators restate src/gpu/tonemap/{linear,reinhard,hable,aces,uchimura,amd}.comp and
// srgbEncode (src/gpu/util.glsl:13-18); host side mirrors Tonemapper::* (src/gfx/modules/tonemapper.ixx).
// HBM-streaming kernel: 16 B (accumulator) or 8 B (RGBA16F) in, 4 B out per pixel.
#include "context.cuh"

namespace {

struct TonemapParams { int mode; float exposure; float p[6]; float amd_b, amd_c; };

// x^e for x >= 0 as ex2(e * lg2(x)) on the SFU -- what GLSL pow() compiles to on the reference's GPU path
// (Vulkan precision: inherited from exp2/log2).  The general powf costs ~15x the instructions and made this
// kernel ALU-bound at 88 us per 1080p frame; the 8-bit result differs from the libm oracle by at most one
// code value (tests/test_gpu_spheres.py::test_tonemap_operators).  0^e = 0 for e > 0, 1^e = 1.
MRT_D float pw(float x, float e) { return __powf(x, e); }

MRT_D float srgb1(float c) { return c < 0.0031308f ? 12.92f * c : 1.055f * pw(c, 1.0f / 2.4f) - 0.055f; }
MRT_D float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }

// amd.comp:22-34 -- depends only on the push constants: evaluated once per call on the host
float col_tone_b(float hdrMax, float contrast, float shoulder, float midIn, float midOut) {
    return -((-powf(midIn, contrast) +
              (midOut * (powf(hdrMax, contrast * shoulder) * powf(midIn, contrast) -
                         powf(hdrMax, contrast) * powf(midIn, contrast * shoulder) * midOut)) /
                  (powf(hdrMax, contrast * shoulder) * midOut - powf(midIn, contrast * shoulder) * midOut)) /
             (powf(midIn, contrast * shoulder) * midOut));
}
float col_tone_c(float hdrMax, float contrast, float shoulder, float midIn, float midOut) {
    return (powf(hdrMax, contrast * shoulder) * powf(midIn, contrast) -
            powf(hdrMax, contrast) * powf(midIn, contrast * shoulder) * midOut) /
           (powf(hdrMax, contrast * shoulder) * midOut - powf(midIn, contrast * shoulder) * midOut);
}

// amd.comp:36-69
MRT_D float3 tm_amd(float3 color, const TonemapParams& T) {
    float contrast = T.p[1], shoulder = T.p[2];
    float peak = fmaxf(color.x, fmaxf(color.y, color.z));
    peak = fmaxf(1e-6f, peak);
    float3 ratio = color / peak;
    float z = pw(peak, contrast);
    peak = z / (pw(z, shoulder) * T.amd_b + T.amd_c);
    const float crosstalk = 4.0f;
    float saturation = contrast;
    float crossSaturation = contrast * 16.0f;
    float e0 = saturation / crossSaturation;
    ratio = f3(pw(fabsf(ratio.x), e0), pw(fabsf(ratio.y), e0), pw(fabsf(ratio.z), e0));
    float a = pw(peak, crosstalk);
    ratio = f3(mixf(ratio.x, 1.0f, a), mixf(ratio.y, 1.0f, a), mixf(ratio.z, 1.0f, a));
    ratio = f3(pw(fabsf(ratio.x), crossSaturation), pw(fabsf(ratio.y), crossSaturation),
               pw(fabsf(ratio.z), crossSaturation));
    return ratio * peak;
}

// reinhard.comp:16-19
MRT_D float3 tm_reinhard(float3 v, float maxWhite) {
    float3 mw = f3s(maxWhite * maxWhite);
    float3 num = v * (f3s(1.0f) + v / mw);
    return num / (f3s(1.0f) + v);
}

// hable.comp:14-28
MRT_D float hable1(float x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}

// aces.comp:15-43 (row-vector times matrix: result[i] = dot(v, column i as written))
MRT_D float3 tm_aces(float3 c) {
    float3 a = f3(c.x * 0.59719f + c.y * 0.35458f + c.z * 0.04823f, c.x * 0.07600f + c.y * 0.90834f + c.z * 0.01566f,
                  c.x * 0.02840f + c.y * 0.13383f + c.z * 0.83777f);
    float3 n = a * (a + f3s(0.0245786f)) - f3s(0.000090537f);
    float3 d = a * (a * 0.983729f + f3s(0.4329510f)) + f3s(0.238081f);
    a = n / d;
    float3 o = f3(a.x * 1.60475f + a.y * -0.53108f + a.z * -0.07367f, a.x * -0.10208f + a.y * 1.10813f + a.z * -0.00605f,
                  a.x * -0.00327f + a.y * -0.07276f + a.z * 1.07602f);
    return f3(clampf(o.x, 0.0f, 1.0f), clampf(o.y, 0.0f, 1.0f), clampf(o.z, 0.0f, 1.0f));
}

// uchimura.comp:20-38
MRT_D float uchimura1(float x, float P, float a, float m, float l, float c, float b) {
    float l0 = ((P - m) * l) / a;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = (a * P) / (P - S1);
    float CP = -C2 / P;
    float t = clampf((x - 0.0f) / (m - 0.0f), 0.0f, 1.0f);
    float w0 = 1.0f - t * t * (3.0f - 2.0f * t);
    float w2 = (x < m + l0) ? 0.0f : 1.0f;
    float w1 = 1.0f - w0 - w2;
    float T = m * powf(x / m, c) + b;
    float S = P - (P - S1) * expf(CP * (x - S0));
    float L = m + a * (x - m);
    return T * w0 + L * w1 + S * w2;
}

MRT_D float3 tonemap_pixel(float3 src, const TonemapParams& T) {
    src = src * T.exposure;
    switch (T.mode) {
    case MRT_TONEMAP_LINEAR: return src;
    case MRT_TONEMAP_REINHARD: return tm_reinhard(src, T.p[0]);
    case MRT_TONEMAP_HABLE: {
        float d = hable1(11.2f);
        return f3(hable1(2.0f * src.x) / d, hable1(2.0f * src.y) / d, hable1(2.0f * src.z) / d);
    }
    case MRT_TONEMAP_ACES: return tm_aces(src);
    case MRT_TONEMAP_UCHIMURA:
        return f3(uchimura1(src.x, T.p[0], T.p[1], T.p[2], T.p[3], T.p[4], T.p[5]),
                  uchimura1(src.y, T.p[0], T.p[1], T.p[2], T.p[3], T.p[4], T.p[5]),
                  uchimura1(src.z, T.p[0], T.p[1], T.p[2], T.p[3], T.p[4], T.p[5]));
    default: return tm_amd(src, T);
    }
}

MRT_D uchar4 encode_ldr(float3 mapped) {
    return make_uchar4((unsigned char)unorm8(srgb1(mapped.x)), (unsigned char)unorm8(srgb1(mapped.y)),
                       (unsigned char)unorm8(srgb1(mapped.z)), 255);
}

// SRC 1: read the fp32 accumulator and divide by its sample count (row n7); SRC 2: the denoiser's RGBA8 unorm
// image (denoiser.ixx:56, texel = k/255); SRC 0: the reference's RGBA16F colour image (row a12); SRC 3: that image
// computed on the fly from the accumulator (average rounded to fp16 exactly as k_accum_to_color16 stores it).
template <int SRC>
__global__ void __launch_bounds__(256) k_tonemap(TonemapParams T, const float4* __restrict__ accum,
                                                 const uint2* __restrict__ color16, const uchar4* __restrict__ rgba8,
                                                 uchar4* __restrict__ ldr, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float3 src;
        if (SRC == 1 || SRC == 3) {
            float4 a = __ldcs(&accum[i]);
            src = a.w > 0.0f ? f3(a.x / a.w, a.y / a.w, a.z / a.w) : f3s(0.0f);
            if (SRC == 3)
                src = f3(f16_bits_to_f32(f32_to_f16_bits(src.x)), f16_bits_to_f32(f32_to_f16_bits(src.y)), f16_bits_to_f32(f32_to_f16_bits(src.z)));
        } else if (SRC == 2) {
            uchar4 c = __ldcs(&rgba8[i]);
            src = f3((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f);
        } else {
            uint2 pk = __ldcs(&color16[i]);
            src = f3(f16_bits_to_f32((uint16_t)(pk.x & 0xFFFF)), f16_bits_to_f32((uint16_t)(pk.x >> 16)),
                     f16_bits_to_f32((uint16_t)(pk.y & 0xFFFF)));
        }
        __stcs(&ldr[i], encode_ldr(tonemap_pixel(src, T)));
    }
}

}  // namespace

int tonemap_run(mrt_context* ctx, int mode, float exposure, const float* params, uint32_t nparams, int source) {
    TonemapParams T;
    memset(&T, 0, sizeof T);
    T.mode = mode;
    T.exposure = exposure;
    for (uint32_t i = 0; i < nparams && i < 6; i++) T.p[i] = params[i];
    if (mode == MRT_TONEMAP_AMD) {
        T.amd_b = col_tone_b(T.p[0], T.p[1], T.p[2], T.p[3], T.p[4]);
        T.amd_c = col_tone_c(T.p[0], T.p[1], T.p[2], T.p[3], T.p[4]);
    }
    ctx->ldr_cur ^= 1;
    DevArray<uchar4>& ldr = ctx->ldr_buf[ctx->ldr_cur];
    if (ctx->copy_pending[ctx->ldr_cur]) {  // an async readback may still be draining this buffer
        MRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done[ctx->ldr_cur], 0));
        ctx->copy_pending[ctx->ldr_cur] = false;
    }
    MRT_TRY(dev_reserve(ctx, ldr, ctx->npix));
    size_t n = ctx->npix;
    // grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    unsigned grid = (unsigned)sms * 8u;
    if ((size_t)grid * 256 > n) grid = div_up(n, 256);
    if (grid == 0) grid = 1;
    if (source == MRT_BUF_ACCUM)
        k_tonemap<1><<<grid, 256, 0, ctx->stream>>>(T, ctx->accum.p, nullptr, nullptr, ldr.p, n);
    else if (source == MRT_BUF_TEMPORAL)  // (rgb, 1): the accumulator path with a sample count of one
        k_tonemap<1><<<grid, 256, 0, ctx->stream>>>(T, ctx->tp_rgba[ctx->tp_cur].p, nullptr, nullptr, ldr.p, n);
    else if (source == MRT_BUF_DENOISED)
        k_tonemap<2><<<grid, 256, 0, ctx->stream>>>(T, nullptr, nullptr, ctx->denoised.p, ldr.p, n);
    else if (!ctx->have_color)  // triangle path, RGBA16F image not materialised: same values straight from the accumulator
        k_tonemap<3><<<grid, 256, 0, ctx->stream>>>(T, ctx->accum.p, nullptr, nullptr, ldr.p, n);
    else
        k_tonemap<0><<<grid, 256, 0, ctx->stream>>>(T, nullptr, reinterpret_cast<const uint2*>(ctx->color16.p), nullptr,
                                                    ldr.p, n);
    MRT_LAUNCHED(ctx);
    ctx->have_ldr = true;
    return mrt_check_cuda(ctx, cudaGetLastError(), "tonemap");
}
