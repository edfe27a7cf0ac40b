// vec.cuh -- small fp32 vector helpers + storage-format conversions for the device code.
// All csrc files are compiled with -fmad=false: an FMA is issued only where fmaf() is written,
// so results follow the same fp32 operation order as the GLSL restated in DESIGN.md.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MRT_HD __host__ __device__ __forceinline__
#define MRT_D __device__ __forceinline__

MRT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
MRT_HD float3 f3s(float s) { return make_float3(s, s, s); }
MRT_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
MRT_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
MRT_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
MRT_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
MRT_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
MRT_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
MRT_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
MRT_HD float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MRT_HD float length3(float3 a) { return sqrtf(dot3(a, a)); }
// GLSL normalize(): v / sqrt(dot(v,v)) per component (fixed in DESIGN.md)
MRT_HD float3 normalize3(float3 a) { return a / length3(a); }
MRT_HD float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
MRT_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
MRT_HD float3 exp3(float3 a) { return f3(expf(a.x), expf(a.y), expf(a.z)); }
MRT_HD float comp3(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

struct Mat4 { float m[4][4]; };  // column-major m[col][row], same bytes as mrt_mat4

// GLSL mat4 * vec4: columns scaled by components, summed left to right
MRT_HD float4 mat_vec(const Mat4& M, float x, float y, float z, float w) {
    float4 r;
    r.x = M.m[0][0] * x + M.m[1][0] * y + M.m[2][0] * z + M.m[3][0] * w;
    r.y = M.m[0][1] * x + M.m[1][1] * y + M.m[2][1] * z + M.m[3][1] * w;
    r.z = M.m[0][2] * x + M.m[1][2] * y + M.m[2][2] * z + M.m[3][2] * w;
    r.w = M.m[0][3] * x + M.m[1][3] * y + M.m[2][3] * z + M.m[3][3] * w;
    return r;
}

MRT_HD Mat4 mat_mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; c++) {
        float4 col = mat_vec(a, b.m[c][0], b.m[c][1], b.m[c][2], b.m[c][3]);
        r.m[c][0] = col.x; r.m[c][1] = col.y; r.m[c][2] = col.z; r.m[c][3] = col.w;
    }
    return r;
}

// ---- storage formats (rounding fixed in DESIGN.md "implementation-defined choices") ----

MRT_D uint16_t f32_to_f16_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
MRT_D float f16_bits_to_f32(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// unsigned small float, 5-bit exponent (bias 15), mb mantissa bits, RNE, saturating
MRT_D uint32_t ufloat_pack(float f, int mb) {
    const uint32_t maxv = (30u << mb) | ((1u << mb) - 1u);
    if (!(f > 0.0f)) return 0u;
    uint32_t x = __float_as_uint(f);
    if (x >= 0x7F800000u) return maxv;
    if (x < 0x38800000u) return (uint32_t)rintf(f * (float)(1u << (14 + mb)));
    const int drop = 23 - mb;
    uint32_t r = x - 0x38000000u;
    uint32_t v = r >> drop;
    uint32_t rem = r & ((1u << drop) - 1u);
    uint32_t halfway = 1u << (drop - 1);
    if (rem > halfway || (rem == halfway && (v & 1u))) v++;
    return v > maxv ? maxv : v;
}
MRT_D float ufloat_unpack(uint32_t v, int mb) {
    uint32_t e = v >> mb, m = v & ((1u << mb) - 1u);
    if (e == 0) return (float)m * (6.103515625e-05f / (float)(1u << mb));
    if (e == 31) return m ? __int_as_float(0x7FC00000) : __int_as_float(0x7F800000);
    return __uint_as_float(((e + 112u) << 23) | (m << (23 - mb)));
}
MRT_D uint32_t pack_b10g11r11(float3 c) {
    return ufloat_pack(c.x, 6) | (ufloat_pack(c.y, 6) << 11) | (ufloat_pack(c.z, 5) << 22);
}
MRT_D float3 unpack_b10g11r11(uint32_t p) {
    return f3(ufloat_unpack(p & 0x7FFu, 6), ufloat_unpack((p >> 11) & 0x7FFu, 6), ufloat_unpack(p >> 22, 5));
}
MRT_D uint32_t unorm8(float f) {
    if (!(f == f)) return 0u;
    return (uint32_t)rintf(clampf(f, 0.0f, 1.0f) * 255.0f);
}

// bilinear fetch from a decoded float4 LUT (Vulkan unnormalized-coordinate rule)
MRT_D float3 lut_bilinear(const float4* __restrict__ lut, int w, int h, float u, float v, bool repeat) {
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = x - fx0, fy = y - fy0;
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (repeat) {
        x0 %= w; if (x0 < 0) x0 += w;
        x1 %= w; if (x1 < 0) x1 += w;
        y0 %= h; if (y0 < 0) y0 += h;
        y1 %= h; if (y1 < 0) y1 += h;
    } else {
        x0 = min(max(x0, 0), w - 1); x1 = min(max(x1, 0), w - 1);
        y0 = min(max(y0, 0), h - 1); y1 = min(max(y1, 0), h - 1);
    }
    float4 p00 = __ldg(&lut[y0 * w + x0]), p10 = __ldg(&lut[y0 * w + x1]);
    float4 p01 = __ldg(&lut[y1 * w + x0]), p11 = __ldg(&lut[y1 * w + x1]);
    float gx = 1.0f - fx, gy = 1.0f - fy;
    float3 top = f3(p00.x * gx + p10.x * fx, p00.y * gx + p10.y * fx, p00.z * gx + p10.z * fx);
    float3 bot = f3(p01.x * gx + p11.x * fx, p01.y * gx + p11.y * fx, p01.z * gx + p11.z * fx);
    return f3(top.x * gy + bot.x * fy, top.y * gy + bot.y * fy, top.z * gy + bot.z * fy);
}
