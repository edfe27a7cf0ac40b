// shading.cuh -- per-path device functions shared by the sphere path and the triangle wavefront:
// ray generation (primaryRay.comp:40-56), PCG + blue-noise rotation (random.glsl:22-31,
// secondaryRays.comp:60-62,124-125), Lambertian bounce (random.glsl:10-19, secondaryRays.comp:70-72).
#pragma once
#include "context.cuh"

#define SHADE_PI 3.14159274101257324f  // util.glsl:4 (3.14159265359) rounded to fp32
#define RAY_OFFSET 0.000001f           // secondaryRays.comp:71 (1 mm in km units)

struct RayGen { Mat4 invView, invProj; uint32_t W, H; };

MRT_D void ray_gen(const RayGen& g, uint32_t x, uint32_t y, float3& origin, float3& dir) {
    float pitchx = 1.0f / (float)g.W, pitchy = 1.0f / (float)g.H;
    float u = ((float)x + 0.5f) * pitchx;
    float v = ((float)y + 0.5f) * pitchy;
    v = 1.0f - v;
    float4 o = mat_vec(g.invView, 0.0f, 0.0f, 0.0f, 1.0f);
    float4 vd = mat_vec(g.invProj, u * 2.0f - 1.0f, v * 2.0f - 1.0f, 1.0f, 1.0f);
    float4 wd = mat_vec(g.invView, vd.x, vd.y, vd.z, 0.0f);
    origin = f3(o.x, o.y, o.z);
    dir = normalize3(f3(wd.x, wd.y, wd.z));
}

// random.glsl:22-27
MRT_D uint32_t pcg(uint32_t& v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    v = (word >> 22u) ^ word;
    return v;
}
// random.glsl:29-31 + secondaryRays.comp:60-62
MRT_D float rotated_random(uint32_t& state, float rotation) {
    float x = (float)(pcg(state) & 0xFFFFFFu) / 16777216.0f + rotation;
    return x - floorf(x);
}
// random.glsl:10-19
MRT_D float3 random_sphere_point(float rx, float ry) {
    float ang1 = (rx + 1.0f) * SHADE_PI;
    float u = ry;
    float s = sqrtf(1.0f - u * u);
    float sn, cs;
    sincosf(ang1, &sn, &cs);
    return f3(s * cs, s * sn, u);
}
// secondaryRays.comp:70-72 -- r0 is drawn before r1 (GLSL argument order)
MRT_D void lambert_bounce(float3 pos, float3 n, uint32_t& rng, float rotx, float roty, float3& ro, float3& rd) {
    ro = pos + n * RAY_OFFSET;
    float r0 = rotated_random(rng, rotx);
    float r1 = rotated_random(rng, roty);
    rd = normalize3(n + random_sphere_point(r0 * 2.0f - 1.0f, r1 * 2.0f - 1.0f));
}

MRT_D float2 blue_noise_rotation(const uchar4* bn, uint32_t bnW, uint32_t bnH, uint32_t x, uint32_t y) {
    uchar4 t = __ldg(&bn[(size_t)(y % bnH) * bnW + (x % bnW)]);
    return make_float2((float)t.x / 255.0f, (float)t.y / 255.0f);
}

// intersect.glsl:26-37
MRT_D float ray_sphere(float3 o, float3 d, const mrt_sphere& s) {
    float3 oc = o - f3(s.center[0], s.center[1], s.center[2]);
    float a = dot3(d, d);
    float half_b = dot3(oc, d);
    float c = dot3(oc, oc) - s.radius * s.radius;
    float disc = half_b * half_b - a * c;
    if (disc < 0.0f) return -1.0f;
    return (-half_b - sqrtf(disc)) / a;
}

// G-buffer depth + motion of a hit position (primaryRay.comp:62-64,73-75)
MRT_D void project_hit(const Mat4& PV, const Mat4& PVprev, float3 pos, uint32_t W, uint32_t H, float& depth, float2& motion) {
    float4 ph = mat_vec(PV, pos.x, pos.y, pos.z, 1.0f);
    ph.x /= ph.w; ph.y /= ph.w; ph.z /= ph.w;
    depth = ph.z;
    float4 pp = mat_vec(PVprev, pos.x, pos.y, pos.z, 1.0f);
    pp.x /= pp.w; pp.y /= pp.w;
    float pitchx = 1.0f / (float)W, pitchy = 1.0f / (float)H;
    motion = make_float2((ph.x - pp.x) / pitchx, (ph.y - pp.y) / pitchy);
}

MRT_D void store_gbuffer(uint32_t* vis, uint16_t* depth, uint16_t* normal, uint16_t* motion, size_t p, uint32_t id,
                         float dep, float3 n, float2 mo) {
    vis[p] = id;
    depth[p] = f32_to_f16_bits(dep);
    // RGBA16F, w unused (pathtracer.ixx:58): one 8-byte store
    uint2 pk;
    pk.x = (uint32_t)f32_to_f16_bits(n.x) | ((uint32_t)f32_to_f16_bits(n.y) << 16);
    pk.y = (uint32_t)f32_to_f16_bits(n.z);
    reinterpret_cast<uint2*>(normal)[p] = pk;
    reinterpret_cast<uint32_t*>(motion)[p] = (uint32_t)f32_to_f16_bits(mo.x) | ((uint32_t)f32_to_f16_bits(mo.y) << 16);
}
