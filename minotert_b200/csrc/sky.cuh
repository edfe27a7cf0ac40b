// sky.cuh -- device side of the atmosphere model that lights every path (miss shader).
// Restates src/gpu/sky/{skyAccess,sky}.glsl and secondaryRays.comp:36-58 of the reference in
// CUDA; LUT taps are explicit fp32 bilinear fetches from decoded float4 copies of the LUTs.
#pragma once
#include "../../include/minotert.h"
#include "vec.cuh"

#define MRT_TRANS_W 256
#define MRT_TRANS_H 64
#define MRT_MULTI_W 32
#define MRT_MULTI_H 32
#define MRT_VIEW_W 192
#define MRT_VIEW_H 108
#define MRT_AERIAL_SIZE 32              // Sky::AerialPerspectiveSize, sky.ixx:191
#define MRT_AERIAL_KM_PER_SLICE 4.0f    // AP_KM_PER_SLICE, skyAccess.glsl:9

#define SKY_PI 3.14159274101257324f          // constants.glsl:4 rounded to fp32
#define SKY_PLANET_RADIUS_OFFSET 0.01f       // sky.glsl:13

struct SkyLuts {
    const float4* trans;  // 256x64 decoded RGBA16F
    const float4* multi;  // 32x32  decoded RGBA16F
    const float4* view;   // 192x108 decoded B10G11R11
};

// skyAccess.glsl:11-17
MRT_D float sky_unit_to_sub_uv(float u, float res) { return (u + 0.5f / res) * (res / (res + 1.0f)); }
MRT_D float sky_sub_uv_to_unit(float u, float res) { return (u - 0.5f / res) * (res / (res - 1.0f)); }

// skyAccess.glsl:36-52
MRT_D float2 sky_trans_params_to_uv(float viewHeight, float cosZen, float bottom, float top) {
    float H = sqrtf(fmaxf(0.0f, top * top - bottom * bottom));
    float rho = sqrtf(fmaxf(0.0f, viewHeight * viewHeight - bottom * bottom));
    float disc = viewHeight * viewHeight * (cosZen * cosZen - 1.0f) + top * top;
    float d = fmaxf(0.0f, (-viewHeight * cosZen + sqrtf(disc)));
    float d_min = top - viewHeight;
    float d_max = rho + H;
    return make_float2((d - d_min) / (d_max - d_min), rho / H);
}

// sky.glsl:58-77
MRT_D float sky_ray_sphere_nearest(float3 r0, float3 rd, float3 s0, float sR) {
    float a = dot3(rd, rd);
    float3 s0_r0 = r0 - s0;
    float b = 2.0f * dot3(rd, s0_r0);
    float c = dot3(s0_r0, s0_r0) - (sR * sR);
    float delta = b * b - 4.0f * a * c;
    if (delta < 0.0f || a == 0.0f) return -1.0f;
    float sq = sqrtf(delta);
    float sol0 = (-b - sq) / (2.0f * a);
    float sol1 = (-b + sq) / (2.0f * a);
    if (sol0 < 0.0f && sol1 < 0.0f) return -1.0f;
    if (sol0 < 0.0f) return fmaxf(0.0f, sol1);
    else if (sol1 < 0.0f) return fmaxf(0.0f, sol0);
    return fmaxf(0.0f, fminf(sol0, sol1));
}

// sky.glsl:129-155
MRT_D float3 sky_sun_luminance(const mrt_atmosphere_params& A, const SkyLuts& L, float3 worldPos, float3 worldDir,
                               float3 sunDir, float3 sunIll) {
    const float SunRadius = 0.5f * 0.505f * 3.14159f / 180.0f;
    float cosAngle = dot3(worldDir, sunDir);
    if (cosAngle > cosf(SunRadius)) {
        float t = sky_ray_sphere_nearest(worldPos, worldDir, f3s(0.0f), A.bottomRadius);
        if (t < 0.0f) {
            float2 uvUp = sky_trans_params_to_uv(A.bottomRadius, 1.0f, A.bottomRadius, A.topRadius);
            float pHeight = length3(worldPos);
            float3 up = worldPos / pHeight;
            float sunZen = dot3(sunDir, up);
            float2 uvSun = sky_trans_params_to_uv(pHeight, sunZen, A.bottomRadius, A.topRadius);
            float angle = acosf(clampf(cosAngle, -1.0f, 1.0f));
            float radiusRatio = angle / SunRadius;
            float limb = sqrtf(clampf(1.0f - radiusRatio * radiusRatio, 0.0001f, 1.0f));
            float3 inSpace = sunIll / lut_bilinear(L.trans, MRT_TRANS_W, MRT_TRANS_H, uvUp.x, uvUp.y, false);
            return (inSpace * lut_bilinear(L.trans, MRT_TRANS_W, MRT_TRANS_H, uvSun.x, uvSun.y, false)) * limb;
        }
    }
    return f3s(0.0f);
}

// secondaryRays.comp:36-58 (+ skyAccess.glsl:87-117 inlined).  with_sun = false: the sky-view term only (bounce rays of
// a path whose sun light comes from shadow rays, MRT_SECONDARY_NEE_SUN)
MRT_D float3 sky_color(const mrt_atmosphere_params& A, const SkyLuts& L, float3 cameraPos, float3 dir, bool with_sun = true) {
    float3 worldPos = cameraPos + f3(0.0f, 0.0f, A.bottomRadius);
    float3 up = normalize3(worldPos);
    float cosZen = dot3(dir, up);
    float viewHeight = length3(worldPos);
    const float3 sunDir = f3(-0.435286462f, 0.818654716f, 0.374606609f);
    const float3 sunIll = f3s(8.0f);
    float3 side = normalize3(cross3(up, dir));
    float3 fwd = normalize3(cross3(side, up));
    float lx = dot3(sunDir, fwd), ly = dot3(sunDir, side);
    float lightViewCos = lx / sqrtf(lx * lx + ly * ly);
    bool ground = sky_ray_sphere_nearest(worldPos, dir, f3s(0.0f), A.bottomRadius) >= 0.0f;

    float vHorizon = sqrtf(viewHeight * viewHeight - A.bottomRadius * A.bottomRadius);
    float cosBeta = vHorizon / viewHeight;
    float beta = acosf(cosBeta);
    float zenithHorizonAngle = SKY_PI - beta;
    float v;
    if (!ground) {
        float coord = acosf(cosZen) / zenithHorizonAngle;
        coord = 1.0f - coord;
        coord = sqrtf(coord);
        coord = 1.0f - coord;
        v = coord * 0.5f;
    } else {
        float coord = (acosf(cosZen) - zenithHorizonAngle) / beta;
        coord = sqrtf(coord);
        v = coord * 0.5f + 0.5f;
    }
    float u = sqrtf(-lightViewCos * 0.5f + 0.5f);
    u = sky_unit_to_sub_uv(u, (float)MRT_VIEW_W);
    v = sky_unit_to_sub_uv(v, (float)MRT_VIEW_H);
    float3 skyView = lut_bilinear(L.view, MRT_VIEW_W, MRT_VIEW_H, u, v, true);
    if (!with_sun) return skyView;
    float3 sun = sky_sun_luminance(A, L, worldPos, dir, sunDir, sunIll) * (f3s(120000.0f) / sunIll);
    return skyView + sun;
}

// ---- the sun as a sampled light (contract: oracle/minote_oracle.c sun_centre_radiance, nee_sun_sample) ----
// radiance of the sun's centre seen from pos: sky_sun_luminance without the disc test and the limb factor, times the
// 120000 / illuminance scale of sky_color; zero when the centre direction meets the ground sphere
MRT_D float3 sky_sun_centre_radiance(const mrt_atmosphere_params& A, const SkyLuts& L, float3 pos) {
    const float3 sunDir = f3(-0.435286462f, 0.818654716f, 0.374606609f);
    const float3 sunIll = f3s(8.0f);
    float3 worldPos = pos + f3(0.0f, 0.0f, A.bottomRadius);
    if (sky_ray_sphere_nearest(worldPos, sunDir, f3s(0.0f), A.bottomRadius) >= 0.0f) return f3s(0.0f);
    float2 uvUp = sky_trans_params_to_uv(A.bottomRadius, 1.0f, A.bottomRadius, A.topRadius);
    float pHeight = length3(worldPos);
    float3 up = worldPos / pHeight;
    float sunZen = dot3(sunDir, up);
    float2 uvSun = sky_trans_params_to_uv(pHeight, sunZen, A.bottomRadius, A.topRadius);
    float3 inSpace = sunIll / lut_bilinear(L.trans, MRT_TRANS_W, MRT_TRANS_H, uvUp.x, uvUp.y, false);
    return (inSpace * lut_bilinear(L.trans, MRT_TRANS_W, MRT_TRANS_H, uvSun.x, uvSun.y, false)) * (f3s(120000.0f) / sunIll);
}
// one sun sample: direction uniform in solid angle inside the disc, weight = limb darkening * Omega / pi
MRT_D void sky_nee_sun_sample(float u0, float u1, float3& l, float& weight) {
    const float3 sun = f3(-0.435286462f, 0.818654716f, 0.374606609f);
    const float SunRadius = 0.5f * 0.505f * 3.14159f / 180.0f;
    const float oneMinusCos = 1.0f - cosf(SunRadius);
    float cosT = 1.0f - u0 * oneMinusCos;
    float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
    float phi = (u1 * 2.0f) * SKY_PI;
    float3 t = normalize3(cross3(f3(0.0f, 0.0f, 1.0f), sun));
    float3 b = cross3(sun, t);
    float sn, cs;
    sincosf(phi, &sn, &cs);
    l = normalize3((t * (cs * sinT) + b * (sn * sinT)) + sun * cosT);
    weight = sqrtf(clampf(1.0f - u0, 0.0001f, 1.0f)) * (2.0f * oneMinusCos);
}

// aerial perspective of a surface at distance t (km) seen through image position (u, v): weight * trilinear(volume) at
// (u, v, sqrt(slice / 32)), clamp to edge; .xyz luminance scattered towards the camera, .w = 1 - transmittance
MRT_D float4 sky_aerial_lookup(const float4* __restrict__ vol, float u, float v, float t) {
    float slice = t * (1.0f / MRT_AERIAL_KM_PER_SLICE);  // aerialPerspectiveDepthToSlice, skyAccess.glsl:119-121
    float weight = 1.0f;
    if (slice < 0.5f) {
        weight = clampf(slice * 2.0f, 0.0f, 1.0f);
        slice = 0.5f;
    }
    const float w = sqrtf(slice / (float)MRT_AERIAL_SIZE);
    const float c[3] = {u * (float)MRT_AERIAL_SIZE - 0.5f, v * (float)MRT_AERIAL_SIZE - 0.5f, w * (float)MRT_AERIAL_SIZE - 0.5f};
    int i0[3];
    float f[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float fl = floorf(c[a]);
        i0[a] = (int)fl;
        f[a] = c[a] - fl;
    }
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const int x = min(max(i0[0] + dx, 0), MRT_AERIAL_SIZE - 1), y = min(max(i0[1] + dy, 0), MRT_AERIAL_SIZE - 1),
                          z = min(max(i0[2] + dz, 0), MRT_AERIAL_SIZE - 1);
                const float4 tx = __ldg(&vol[((size_t)z * MRT_AERIAL_SIZE + y) * MRT_AERIAL_SIZE + x]);
                const float wt = ((dx ? f[0] : 1.0f - f[0]) * (dy ? f[1] : 1.0f - f[1])) * (dz ? f[2] : 1.0f - f[2]);
                acc = make_float4(acc.x + tx.x * wt, acc.y + tx.y * wt, acc.z + tx.z * wt, acc.w + tx.w * wt);
            }
    return make_float4(acc.x * weight, acc.y * weight, acc.z * weight, acc.w * weight);
}
