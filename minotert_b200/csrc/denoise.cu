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
// Kernel: one thread per pixel, 32x8-pixel CTA (a warp = one image row of the tile), shared tile with a halo of
// radius+1 texels holding {r, g, b, depth} and {nx, ny, nz} as fp32 (converted ONCE per CTA when the tile is filled; two
// conflict-free LDS.128 per texel); image-edge clamping is applied when the tile is filled, so the tap loop needs none.
// Everything about a tap that does not depend on the pixel's column is evaluated once on the host, with the shader's own
// fp32 operations: which taps exist, the spatial Gaussian, and -- the part the first version of this kernel recomputed per
// pixel and per tap, a third of its instructions -- the row the LinearClamp sampler reads and its k/256 weight.  The taps
// of one d.x (a "column") share the fractional part of d.y and step down one row at a time; the sampler's fp32
// evaluation agrees with that on all but a handful of (image row, column) pairs.  So the table holds one 16-byte descriptor
// per (image row, column) -- first texel, tap count, the two row weights, a class (DN_COL_*) -- which a warp reads with one
// broadcast load per column, plus one 16-byte record per (image row, tap) for the columns whose pattern breaks on that
// row.  Taps are visited in the shader's order (same sums, bit for bit, as the first version).
// Sampler rule (the oracle's, oracle/minote_oracle.c:texn_bilinear): bilinear weights carry 8 fractional bits and
// zero-weight texels are not read.  d.x is integral, so in x every tap is the texel centre px + d.x (the fp32
// residue of uv + d/size is < 2^-10 texel for images up to 4096 wide and rounds to weight 0): one column, no
// x-lerp.  In y the tap reads rows floor(y) and floor(y) + 1 with weights (1 - k/256, k/256); k = 0 and k = 256 (and every
// tap whose d.y is integral) are single texels.  Consecutive taps of a column step down one row, so the lower texel of
// one tap is the upper texel of the next: the loop keeps the texel in registers (two register sets that swap roles every
// tap) -- one shared-memory texel per tap instead of two.
// Deliberate deviations, all at the 1e-6 relative level and absorbed by the bar in
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
constexpr int DN_BX = 32, DN_BY = DN_ROWS;

struct DenoiseArgs {
    uint32_t W, H;
    float invThresholdSqx2Log2e, invThresholdSqrt2PI;
    float nearPlane;
    uint32_t frameCounter;
    int halo;   // radius + 1
    int ntaps;
    int ncols;  // tap columns (2 radius + 1)
};

// record of one (image row, tap): .x = (tile offset of the upper texel relative to the pixel) * 16 | flags, .w = spatial Gaussian
// descriptor of one (image row, tap column): .x = (tile offset of the column's first texel) * 16 | class, .y = taps in the column;
// class DN_COL_LERP: .z / .w = the two row weights shared by all its taps; DN_COL_GENERIC: .z = index of its first tap record
enum : int { DN_COL_GENERIC = 0,  // walk the per-tap records (rows whose rounding breaks the pattern)
             DN_COL_LERP = 1,     // every tap interpolates rows r, r + 1 with the same weights, r stepping down by one
             DN_COL_SINGLE = 2 }; // every tap is a single texel (d.y integral), rows stepping down by one
enum : int { DN_LERP = 1,    // the tap interpolates rows (.y = weight of the lower row, .z = 1 - .y); else a single texel
             DN_REUSE = 2 }; // the upper texel is the previous tap's lower texel (still in registers)

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

struct DnSums { float z, r, g, b; };
struct DnConsts { float centreDist, cnx, cny, cnz, nearPlane, k1, k2; };

MRT_D float lds1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
MRT_D float4 lds4(uint32_t addr) {  // 32-bit shared-memory address
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// the weight of one tap and its accumulation (bilateral.comp:52-61)
MRT_D void dn_accumulate(DnSums& sum, const DnConsts& K, float blur, float r, float g, float b, float z, float nx, float ny, float nz) {
    float dZ = K.nearPlane * rcp_approx(z) - K.centreDist;
    dZ *= 100.0f;
    const float dN = fmaf(nz, K.cnz, fmaf(ny, K.cny, nx * K.cnx));
    // exp(c * invThresholdSqx2) = 2^(c * invThresholdSqx2 * log2 e)
    const float deltaFactor = ex2_approx(clampf(dN - dZ * dZ, 0.0f, 1.0f) * K.k1) * K.k2 * blur;
    sum.z += deltaFactor;
    sum.r = fmaf(deltaFactor, r, sum.r);
    sum.g = fmaf(deltaFactor, g, sum.g);
    sum.b = fmaf(deltaFactor, b, sum.b);
}

// One tap (bilateral.comp:49-62).  P = register set of the upper texel, Q = of the lower one.  Both branches are uniform
// across the warp (the record belongs to the image row) and each carries its own copy of the accumulation, so that
// neither pays register moves to meet the other.
#define DN_TAP(P, Q, RC)                                                                                               \
    {                                                                                                                  \
        const float4 rc = RC;                                                                                          \
        const int code = __float_as_int(rc.x);                                                                         \
        const uint32_t at = centre_addr + (uint32_t)(code & ~3);                                                       \
        if (!(code & DN_REUSE)) { P##c = lds4(at); P##n = lds4(at + normal_off); }                                     \
        if (code & DN_LERP) {                                                                                          \
            Q##c = lds4(at + row_bytes); Q##n = lds4(at + row_bytes + normal_off);                                     \
            const float fy = rc.y, gy = rc.z;                                                                          \
            dn_accumulate(sum, K, rc.w, fmaf(Q##c.x, fy, P##c.x * gy), fmaf(Q##c.y, fy, P##c.y * gy),                  \
                          fmaf(Q##c.z, fy, P##c.z * gy), fmaf(Q##c.w, fy, P##c.w * gy), fmaf(Q##n.x, fy, P##n.x * gy), \
                          fmaf(Q##n.y, fy, P##n.y * gy), fmaf(Q##n.z, fy, P##n.z * gy));                               \
        } else {                                                                                                       \
            dn_accumulate(sum, K, rc.w, P##c.x, P##c.y, P##c.z, P##c.w, P##n.x, P##n.y, P##n.z);                       \
        }                                                                                                              \
    }

// bilateral.comp:23-76.  BLUR_SMEM: the per-tap spatial Gaussian is staged in shared memory (always, except for the largest
// radii, whose tile leaves no room for it).
template <bool BLUR_SMEM>
__global__ void __launch_bounds__(DN_BX* DN_BY)
    k_denoise_bilateral(DenoiseArgs A, const uint2* __restrict__ color16, const uint16_t* __restrict__ depth16,
                        const uint2* __restrict__ normal16, const float4* __restrict__ recs, const float4* __restrict__ cols,
                        const float* __restrict__ blur, uchar4* __restrict__ out) {
    extern __shared__ float4 tile[];
    const int TW = DN_BX + 2 * A.halo, TH = DN_BY + 2 * A.halo;
    float4* const tileC = tile;            // r, g, b, depth
    float4* const tileN = tile + TW * TH;  // nx, ny, nz, -
    float* const sblur = reinterpret_cast<float*>(tile + 2 * TW * TH);  // the spatial Gaussian of every tap
    const int x0 = (int)blockIdx.x * DN_BX - A.halo, y0 = (int)blockIdx.y * DN_BY - A.halo;
    const int tid = threadIdx.y * DN_BX + threadIdx.x;
    if (BLUR_SMEM)
        for (int i = tid; i < A.ntaps; i += DN_BX * DN_BY) sblur[i] = __ldg(&blur[i]);
    for (int i = tid; i < TW * TH; i += DN_BX * DN_BY) {
        int ty = i / TW, tx = i - ty * TW;
        int gx = min(max(x0 + tx, 0), (int)A.W - 1), gy = min(max(y0 + ty, 0), (int)A.H - 1);  // ClampToEdge
        size_t g = (size_t)gy * A.W + gx;
        const uint2 c = __ldg(&color16[g]);
        const uint2 n = __ldg(&normal16[g]);
        const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&c.x));
        const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&c.y));
        const float2 nxy = __half22float2(*reinterpret_cast<const __half2*>(&n.x));
        const float2 nzw = __half22float2(*reinterpret_cast<const __half2*>(&n.y));
        tileC[i] = make_float4(rg.x, rg.y, ba.x, __half2float(__ushort_as_half(__ldg(&depth16[g]))));
        tileN[i] = make_float4(nxy.x, nxy.y, nzw.x, 0.0f);
    }
    __syncthreads();
    const uint32_t px = blockIdx.x * DN_BX + threadIdx.x, py = blockIdx.y * DN_BY + threadIdx.y;
    if (px >= A.W || py >= A.H) return;

    const int centre_idx = ((int)py - y0) * TW + ((int)px - x0);
    const float4 cc = tileC[centre_idx], cn = tileN[centre_idx];
    float3 filtered;
    if (cc.w < 0.0f) {  // bilateral.comp:36
        filtered = f3(cc.x, cc.y, cc.z);
    } else {
        DnConsts K;
        K.centreDist = A.nearPlane * rcp_approx(cc.w);
        K.cnx = cn.x; K.cny = cn.y; K.cnz = cn.z;
        K.nearPlane = A.nearPlane; K.k1 = A.invThresholdSqx2Log2e; K.k2 = A.invThresholdSqrt2PI;
        asm volatile("" : "+f"(K.nearPlane), "+f"(K.k1), "+f"(K.k2));  // keep the kernel parameters in registers
        DnSums sum = {0.0f, 0.0f, 0.0f, 0.0f};
        // byte addresses: record offsets are tile offsets * 16, flags in the two low bits
        const uint32_t centre_addr = (uint32_t)__cvta_generic_to_shared(tileC + centre_idx);
        const uint32_t normal_off = (uint32_t)(TW * TH) * 16u, row_bytes = (uint32_t)TW * 16u;
        const float4* rp = recs + (size_t)py * A.ntaps;
        const float4* cp = cols + (size_t)py * A.ncols;
        const uint32_t blur_base = (uint32_t)__cvta_generic_to_shared(sblur);
        uint32_t blur_addr = blur_base;  // of the current column's first tap
#define DN_BLUR(ba) (BLUR_SMEM ? lds1(ba) : __ldg(blur + (((ba) - blur_base) >> 2)))
        float4 Xc = cc, Xn = cn, Yc = cc, Yn = cn;
        float4 cd = __ldg(cp);
        for (int col = 0; col < A.ncols; col++) {
            const int code = __float_as_int(cd.x), n = __float_as_int(cd.y);
            const float fy = cd.z, gy = cd.w;
            const int first_rec = __float_as_int(cd.z);
            if (col + 1 < A.ncols) cd = __ldg(cp + col + 1);  // the next column's descriptor travels while this one is summed
            uint32_t at = centre_addr + (uint32_t)(code & ~3);
            uint32_t ba = blur_addr;
            blur_addr += 4u * (uint32_t)n;
            const int cls = code & 3;
            if (cls == DN_COL_LERP) {
                // rows r, r + 1, ... : the lower texel of one tap is the upper texel of the next (X and Y swap roles)
                Xc = lds4(at); Xn = lds4(at + normal_off);
                int i = 0;
#define DN_LERP_TAP(P, Q)                                                                                              \
    {                                                                                                                  \
        at += row_bytes;                                                                                               \
        Q##c = lds4(at); Q##n = lds4(at + normal_off);                                                                 \
        dn_accumulate(sum, K, DN_BLUR(ba), fmaf(Q##c.x, fy, P##c.x * gy), fmaf(Q##c.y, fy, P##c.y * gy),                  \
                      fmaf(Q##c.z, fy, P##c.z * gy), fmaf(Q##c.w, fy, P##c.w * gy), fmaf(Q##n.x, fy, P##n.x * gy),     \
                      fmaf(Q##n.y, fy, P##n.y * gy), fmaf(Q##n.z, fy, P##n.z * gy));                                   \
        ba += 4u;                                                                                                      \
    }
#pragma unroll 2
                for (; i + 1 < n; i += 2) {
                    DN_LERP_TAP(X, Y)
                    DN_LERP_TAP(Y, X)
                }
                if (i < n) DN_LERP_TAP(X, Y)
#undef DN_LERP_TAP
            } else if (cls == DN_COL_SINGLE) {
#pragma unroll 2
                for (int i = 0; i < n; i++) {
                    const float4 tc = lds4(at), tn = lds4(at + normal_off);
                    dn_accumulate(sum, K, DN_BLUR(ba), tc.x, tc.y, tc.z, tc.w, tn.x, tn.y, tn.z);
                    at += row_bytes;
                    ba += 4u;
                }
            } else {
                // per-tap records: the sampler's row / weight of this image row do not follow the column's pattern
                const float4* r = rp + first_rec;
                int i = 0;
                for (; i + 1 < n; i += 2) {
                    DN_TAP(X, Y, __ldg(r + i))
                    DN_TAP(Y, X, __ldg(r + i + 1))
                }
                if (i < n) DN_TAP(X, Y, __ldg(r + i))
            }
        }
        filtered = f3(sum.r / sum.z, sum.g / sum.z, sum.b / sum.z);
    }
    // bilateral.comp:71-73: one PCG draw per pixel, the same value on r, g and b
    uint32_t seed = px * 709u + py * 1153u + A.frameCounter * 1361u;
    const float noise = (float)(pcg(seed) & 0xFFFFFFu) / 16777216.0f * 0.005f;
    // alpha: sum(w * 1) / sum(w) = 1 up to rounding -> 255
    __stcs(&out[(size_t)py * A.W + px], make_uchar4((unsigned char)unorm8(filtered.x + noise), (unsigned char)unorm8(filtered.y + noise),
                                                    (unsigned char)unorm8(filtered.z + noise), 255));
}

}  // namespace

// The loops of smartDeNoise (bilateral.comp:43-47) in the shader's own fp32 arithmetic, once per image row: which taps
// exist, the spatial Gaussian, and where the sampler reads -- all of it depends only on (sigma, kSigma, image height, row).
// recs[row * ntaps + tap]: one record per tap; cols[row * ncols + column]: one descriptor per tap column (the taps of one
// d.x); blur[tap].  Returns ntaps.
static int build_tap_records(float sigma, float kSigma, uint32_t H, int tileW, std::vector<float4>& recs, std::vector<float4>& cols,
                             std::vector<float>& blur, int* ncols_out) {
    const float INV_PI = 0.31830988618379067153776752674503f;
    const float radius = roundf(kSigma * sigma);
    const float radQ = radius * radius;
    const float invSigmaQx2 = .5f / (sigma * sigma);
    const float invSigmaQx2PI = INV_PI * invSigmaQx2;
    const float sizeY = (float)H;
    struct Tap { float dx, dy, blur; bool whole; };
    std::vector<Tap> taps;
    std::vector<int> col_first;  // first tap of each column, then ntaps
    for (float dx = -radius; dx <= radius; dx++) {
        const float pt = sqrtf(radQ - dx * dx);
        col_first.push_back((int)taps.size());
        for (float dy = -pt; dy <= pt; dy++)
            taps.push_back(Tap{dx, dy, expf(-(dx * dx + dy * dy) * invSigmaQx2) * invSigmaQx2PI, dy == rintf(dy)});
    }
    const int ntaps = (int)taps.size(), ncols = (int)col_first.size();
    col_first.push_back(ntaps);
    blur.resize(ntaps);
    for (int k = 0; k < ntaps; k++) blur[k] = taps[k].blur;
    recs.resize((size_t)H * ntaps);
    cols.resize((size_t)H * ncols);
    std::vector<int> offs(ntaps);
    std::vector<float> fys(ntaps);
    for (uint32_t py = 0; py < H; py++) {
        const float uvy = ((float)py + 0.5f) / sizeY;
        bool prev_lerp = false;
        int prev_lower = 0;  // tile offset of the previous tap's lower texel
        for (int k = 0; k < ntaps; k++) {
            const Tap& t = taps[k];
            // In x the tap is the texel centre px + d.x: (uv.x + d.x/W) * W - 0.5 differs from it by the fp32 rounding of
            // uv.x and d.x/W, at most 2^-23 * W * 2 < 2^-10 texel for W <= 4096, which the 8-bit weight rounds to 0.
            // The same holds in y when d.y is integral; otherwise the row position and its k/256 weight are evaluated
            // exactly as the oracle's sampler computes them.
            int row;
            float fy = 0.0f;
            if (t.whole) {
                row = (int)t.dy;
            } else {
                const float y = (uvy + t.dy / sizeY) * sizeY - 0.5f;
                const float fy0 = floorf(y);
                fy = rintf((y - fy0) * 256.0f) * (1.0f / 256.0f);
                row = (int)fy0 - (int)py;
                if (1.0f - fy == 0.0f) { row += 1; fy = 0.0f; }  // weight 256/256: the lower texel alone
            }
            const int off = row * tileW + (int)t.dx;
            offs[k] = off;
            fys[k] = fy;
            int flags = 0;
            if (fy != 0.0f) flags |= DN_LERP;
            if (prev_lerp && prev_lower == off) flags |= DN_REUSE;
            prev_lerp = fy != 0.0f;
            prev_lower = off + tileW;
            const int code = off * 16 + flags;  // byte offset in the tile of 16-byte texels; flags in the low bits
            float4 r;
            memcpy(&r.x, &code, 4);
            r.y = fy;
            r.z = 1.0f - fy;
            r.w = t.blur;
            recs[(size_t)py * ntaps + k] = r;
        }
        // Column descriptors.  In exact arithmetic the taps of a column share the fractional part of d.y and step down one
        // row at a time; the sampler's fp32 evaluation agrees with that on all but a handful of (row, column) pairs, which
        // fall back to the per-tap records.
        for (int c = 0; c < ncols; c++) {
            const int k0 = col_first[c], n = col_first[c + 1] - k0;
            bool steps = true, lerp = fys[k0] != 0.0f, same = true;
            for (int k = k0 + 1; k < k0 + n; k++) {
                steps = steps && offs[k] == offs[k - 1] + tileW;
                same = same && fys[k] == fys[k0];
            }
            int cls = DN_COL_GENERIC;
            if (steps && same) cls = lerp ? DN_COL_LERP : DN_COL_SINGLE;
            const int code = cls == DN_COL_GENERIC ? 0 : offs[k0] * 16 + cls;
            float4 d;
            memcpy(&d.x, &code, 4);
            memcpy(&d.y, &n, 4);
            if (cls == DN_COL_GENERIC) { memcpy(&d.z, &k0, 4); d.w = 0.0f; }
            else { d.z = fys[k0]; d.w = 1.0f - fys[k0]; }
            cols[(size_t)py * ncols + c] = d;
        }
    }
    *ncols_out = ncols;
    return ntaps;
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
        std::vector<float4> recs, cols;
        std::vector<float> blur;
        int ncols = 0;
        const int ntaps = build_tap_records(sigma, kSigma, H, DN_BX + 2 * ((int)radius + 1), recs, cols, blur, &ncols);
        // one allocation: [records | column descriptors | blur]
        const size_t total = recs.size() + cols.size() + (blur.size() + 3) / 4;
        MRT_TRY(dev_reserve(ctx, ctx->dn_taps, total));
        // pageable sources: the copies are staged before the calls return
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->dn_taps.p, recs.data(), recs.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->dn_taps.p + recs.size(), cols.data(), cols.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaMemcpyAsync(ctx->dn_taps.p + recs.size() + cols.size(), blur.data(), blur.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dn_ncols = ncols;
        ctx->dn_ntaps = ntaps;
        ctx->dn_key_sigma = sigma; ctx->dn_key_ksigma = kSigma; ctx->dn_key_w = W; ctx->dn_key_h = H;
    }

    DenoiseArgs A;
    A.W = W; A.H = H;
    A.invThresholdSqx2Log2e = .5f / (threshold * threshold) * 1.4426950408889634f;
    A.invThresholdSqrt2PI = 0.39894228040143267793994605993439f / threshold;
    A.nearPlane = nearPlane;
    A.frameCounter = frameCounter;
    A.halo = (int)radius + 1;
    A.ntaps = ctx->dn_ntaps;
    A.ncols = ctx->dn_ncols;
    const float4* const recs = ctx->dn_taps.p;
    const float4* const cols = recs + (size_t)H * A.ntaps;
    const float* const blur = reinterpret_cast<const float*>(cols + (size_t)H * A.ncols);
    const size_t tile_bytes = (size_t)(DN_BX + 2 * A.halo) * (DN_BY + 2 * A.halo) * 2 * sizeof(float4);
    const bool blur_smem = tile_bytes + (size_t)A.ntaps * sizeof(float) <= 227u * 1024u;
    const size_t smem = tile_bytes + (blur_smem ? (size_t)A.ntaps * sizeof(float) : 0u);
    auto* const kernel = blur_smem ? k_denoise_bilateral<true> : k_denoise_bilateral<false>;
    if (smem > 48 * 1024) MRT_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(div_up(W, DN_BX), div_up(H, DN_BY)), block(DN_BX, DN_BY);
    kernel<<<grid, block, smem, ctx->stream>>>(A, reinterpret_cast<const uint2*>(ctx->color16.p), ctx->depth.p,
                                               reinterpret_cast<const uint2*>(ctx->normal.p), recs, cols, blur, ctx->denoised.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "denoise_bilateral");
}
