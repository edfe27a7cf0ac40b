// sky.cu -- atmosphere LUT generation (row a13): transmittance 256x64, multi-scattering 32x32,
// sky-view 192x108.  CUDA restatement of src/gpu/sky/gen{Transmittance,MultiScattering,View}.comp
// and integrateScatteredLuminance (src/gpu/sky/sky.glsl:176-343); host side mirrors
// Atmosphere::Atmosphere / Sky::createView (src/gfx/modules/sky.ixx:91-176,199-262).
#include "context.cuh"
#include "shading.cuh"

namespace {

struct Medium { float3 scattering, extinction, scatteringMie, scatteringRay; };

// sky.glsl:96-125
MRT_D Medium sample_medium(const mrt_atmosphere_params& A, float3 worldPos) {
    float viewHeight = length3(worldPos) - A.bottomRadius;
    float densityMie = expf(A.mieDensityExpScale * viewHeight);
    float densityRay = expf(A.rayleighDensityExpScale * viewHeight);
    float densityOzo = clampf(viewHeight < A.absorptionDensity0LayerWidth
                                  ? A.absorptionDensity0LinearTerm * viewHeight + A.absorptionDensity0ConstantTerm
                                  : A.absorptionDensity1LinearTerm * viewHeight + A.absorptionDensity1ConstantTerm,
                              0.0f, 1.0f);
    Medium m;
    m.scatteringMie = f3(A.mieScattering[0], A.mieScattering[1], A.mieScattering[2]) * densityMie;
    float3 extinctionMie = f3(A.mieExtinction[0], A.mieExtinction[1], A.mieExtinction[2]) * densityMie;
    m.scatteringRay = f3(A.rayleighScattering[0], A.rayleighScattering[1], A.rayleighScattering[2]) * densityRay;
    float3 extinctionRay = m.scatteringRay + f3s(0.0f);
    float3 scatteringOzo = f3s(0.0f);
    float3 extinctionOzo =
        scatteringOzo + f3(A.absorptionExtinction[0], A.absorptionExtinction[1], A.absorptionExtinction[2]) * densityOzo;
    m.scattering = (m.scatteringMie + m.scatteringRay) + scatteringOzo;
    m.extinction = (extinctionMie + extinctionRay) + extinctionOzo;
    return m;
}

// sky.glsl:42-50
MRT_D float mie_phase(float g, float cosTheta) {
    float k = 3.0f / (8.0f * SKY_PI) * (1.0f - g * g) / (2.0f + g * g);
    return k * (1.0f + cosTheta * cosTheta) / powf(1.0f + g * g - 2.0f * g * -cosTheta, 1.5f);
}
MRT_D float rayleigh_phase(float cosTheta) {
    float factor = 3.0f / (16.0f * SKY_PI);
    return factor * (1.0f + cosTheta * cosTheta);
}

struct Scatter { float3 L, opticalDepth, transmittance, multiScatAs1; };

// sky.glsl:176-343.  HAS_TRANS / HAS_MULTI stand for the S_TRANSMITTANCE / S_MULTISCATTERING macros.
template <bool HAS_TRANS, bool HAS_MULTI>
MRT_D Scatter integrate_scattered(const mrt_atmosphere_params& A, const SkyLuts& luts, float3 worldPos, float3 worldDir,
                                  float3 sunDir, bool ground, float sampleCountIni, bool variableSampleCount,
                                  bool mieRayPhase, float tMaxMax, float3 sunIll) {
    Scatter result;
    result.L = result.opticalDepth = result.transmittance = result.multiScatAs1 = f3s(0.0f);
    float3 earthO = f3s(0.0f);
    float tBottom = sky_ray_sphere_nearest(worldPos, worldDir, earthO, A.bottomRadius);
    float tTop = sky_ray_sphere_nearest(worldPos, worldDir, earthO, A.topRadius);
    float tMax = 0.0f;
    if (tBottom < 0.0f) {
        if (tTop < 0.0f) return result;
        tMax = tTop;
    } else if (tTop > 0.0f) {
        tMax = fminf(tTop, tBottom);
    }
    tMax = fminf(tMax, tMaxMax);

    float sampleCount = sampleCountIni, sampleCountFloor = sampleCountIni, tMaxFloor = tMax;
    if (variableSampleCount) {
        float a = clampf(tMax * 0.01f, 0.0f, 1.0f);
        sampleCount = 4.0f * (1.0f - a) + 14.0f * a;  // mix(RAYMARCH_MIN_SPP, RAYMARCH_MAX_SPP, a)
        sampleCountFloor = floorf(sampleCount);
        tMaxFloor = tMax * sampleCountFloor / sampleCount;
    }
    float dt = tMax / sampleCount;

    float uniformPhase = 1.0f / (4.0f * SKY_PI);
    float cosTheta = dot3(sunDir, worldDir);
    float miePhaseValue = mie_phase(A.miePhaseG, -cosTheta);
    float rayleighPhaseValue = rayleigh_phase(cosTheta);

    float3 L = f3s(0.0f), throughput = f3s(1.0f), opticalDepth = f3s(0.0f);
    float t = 0.0f;
    const float sampleSegmentT = 0.3f;
    for (float s = 0.0f; s < sampleCount; s += 1.0f) {
        if (variableSampleCount) {
            float t0 = s / sampleCountFloor;
            float t1 = (s + 1.0f) / sampleCountFloor;
            t0 = t0 * t0;
            t1 = t1 * t1;
            t0 = tMaxFloor * t0;
            if (t1 > 1.0f) t1 = tMax;
            else t1 = tMaxFloor * t1;
            t = t0 + (t1 - t0) * sampleSegmentT;
            dt = t1 - t0;
        } else {
            float newT = tMax * (s + sampleSegmentT) / sampleCount;
            dt = newT - t;
            t = newT;
        }
        float3 P = worldPos + worldDir * t;
        Medium medium = sample_medium(A, P);
        float3 sampleOpticalDepth = medium.extinction * dt;
        float3 sampleTransmittance = exp3(-sampleOpticalDepth);
        opticalDepth = opticalDepth + sampleOpticalDepth;

        float pHeight = length3(P);
        float3 up = P / pHeight;
        float sunZen = dot3(sunDir, up);
        float3 transToSun = f3s(0.0f);
        if (HAS_TRANS) {
            float2 uv = sky_trans_params_to_uv(pHeight, sunZen, A.bottomRadius, A.topRadius);
            transToSun = lut_bilinear(luts.trans, MRT_TRANS_W, MRT_TRANS_H, uv.x, uv.y, false);
        }
        float3 phaseTimesScattering = mieRayPhase
                                          ? medium.scatteringMie * miePhaseValue + medium.scatteringRay * rayleighPhaseValue
                                          : medium.scattering * uniformPhase;
        float tEarth = sky_ray_sphere_nearest(P, sunDir, earthO + up * SKY_PLANET_RADIUS_OFFSET, A.bottomRadius);
        float earthShadow = tEarth >= 0.0f ? 0.0f : 1.0f;

        float3 msL = f3s(0.0f);
        if (HAS_MULTI) {  // sky.glsl:161-172
            float u = clampf(sunZen * 0.5f + 0.5f, 0.0f, 1.0f);
            float v = clampf((length3(P) - A.bottomRadius) / (A.topRadius - A.bottomRadius), 0.0f, 1.0f);
            u = sky_unit_to_sub_uv(u, (float)MRT_MULTI_W);
            v = sky_unit_to_sub_uv(v, (float)MRT_MULTI_H);
            msL = lut_bilinear(luts.multi, MRT_MULTI_W, MRT_MULTI_H, u, v, false);
        }
        float3 S = sunIll * ((transToSun * earthShadow) * phaseTimesScattering + msL * medium.scattering);

        float3 MS = medium.scattering * 1.0f;
        float3 MSint = (MS - MS * sampleTransmittance) / medium.extinction;
        result.multiScatAs1 = result.multiScatAs1 + throughput * MSint;

        float3 Sint = (S - S * sampleTransmittance) / medium.extinction;
        L = L + throughput * Sint;
        throughput = throughput * sampleTransmittance;
    }

    if (ground && tMax == tBottom && tBottom > 0.0f) {
        float3 P = worldPos + worldDir * tBottom;
        float pHeight = length3(P);
        float3 up = P / pHeight;
        float sunZen = dot3(sunDir, up);
        float3 transToSun = f3s(0.0f);
        if (HAS_TRANS) {
            float2 uv = sky_trans_params_to_uv(pHeight, sunZen, A.bottomRadius, A.topRadius);
            transToSun = lut_bilinear(luts.trans, MRT_TRANS_W, MRT_TRANS_H, uv.x, uv.y, false);
        }
        float NdotL = clampf(dot3(normalize3(up), normalize3(sunDir)), 0.0f, 1.0f);
        float3 term = (sunIll * transToSun) * throughput;
        term = term * NdotL;
        term = term * f3(A.groundAlbedo[0], A.groundAlbedo[1], A.groundAlbedo[2]);
        term = term / SKY_PI;
        L = L + term;
    }
    result.L = L;
    result.opticalDepth = opticalDepth;
    result.transmittance = throughput;
    return result;
}

MRT_D void store_rgba16f(uint16_t* dst16, float4* dstf, int i, float3 c) {
    uint16_t r = f32_to_f16_bits(c.x), g = f32_to_f16_bits(c.y), b = f32_to_f16_bits(c.z);
    dst16[4 * i + 0] = r; dst16[4 * i + 1] = g; dst16[4 * i + 2] = b; dst16[4 * i + 3] = 0x3C00;
    dstf[i] = make_float4(f16_bits_to_f32(r), f16_bits_to_f32(g), f16_bits_to_f32(b), 1.0f);
}

// genTransmittance.comp:21-44
__global__ void k_gen_transmittance(mrt_atmosphere_params A, uint16_t* out16, float4* outf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= MRT_TRANS_W || y >= MRT_TRANS_H) return;
    float u = ((float)x + 0.5f) / (float)MRT_TRANS_W, v = ((float)y + 0.5f) / (float)MRT_TRANS_H;
    // skyAccess.glsl:19-34
    float bottom = A.bottomRadius, top = A.topRadius;
    float H = sqrtf(top * top - bottom * bottom);
    float rho = H * v;
    float viewHeight = sqrtf(rho * rho + bottom * bottom);
    float d_min = top - viewHeight, d_max = rho + H;
    float d = d_min + u * (d_max - d_min);
    float cosZen = d == 0.0f ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * viewHeight * d);
    cosZen = clampf(cosZen, -1.0f, 1.0f);

    float3 worldPos = f3(0.0f, 0.0f, viewHeight);
    float3 worldDir = f3(0.0f, sqrtf(1.0f - cosZen * cosZen), cosZen);
    SkyLuts none{nullptr, nullptr, nullptr};
    Scatter r = integrate_scattered<false, false>(A, none, worldPos, worldDir, f3s(1.0f), false, 40.0f, false, false,
                                                  9000000.0f, f3s(1.0f));
    store_rgba16f(out16, outf, y * MRT_TRANS_W + x, exp3(-r.opticalDepth));
}

// genMultiScattering.comp:26-146: one block of 64 directions per texel, shared-memory tree sum
__global__ void k_gen_multiscattering(mrt_atmosphere_params A, SkyLuts luts, uint16_t* out16, float4* outf) {
    __shared__ float3 shMS[64], shL[64];
    int x = blockIdx.x, y = blockIdx.y, z = threadIdx.x;
    float u = ((float)x + 0.5f) / (float)MRT_MULTI_W, v = ((float)y + 0.5f) / (float)MRT_MULTI_H;
    u = sky_sub_uv_to_unit(u, (float)MRT_MULTI_W);
    v = sky_sub_uv_to_unit(v, (float)MRT_MULTI_H);
    float cosSunZen = u * 2.0f - 1.0f;
    float3 sunDir = f3(0.0f, sqrtf(clampf(1.0f - cosSunZen * cosSunZen, 0.0f, 1.0f)), cosSunZen);
    float viewHeight = A.bottomRadius + clampf(v + SKY_PLANET_RADIUS_OFFSET, 0.0f, 1.0f) *
                                            (A.topRadius - A.bottomRadius - SKY_PLANET_RADIUS_OFFSET);
    float3 worldPos = f3(0.0f, 0.0f, viewHeight);
    const float sphereSolidAngle = 4.0f * SKY_PI;
    const float isotropicPhase = 1.0f / sphereSolidAngle;
    const float sqrtSample = 8.0f;
    float i = 0.5f + (float)(z / 8);
    float j = 0.5f + (float)(z % 8);
    float randA = i / sqrtSample, randB = j / sqrtSample;
    float theta = 2.0f * SKY_PI * randA, phi = SKY_PI * randB;
    float cosPhi = cosf(phi), sinPhi = sinf(phi), cosTheta = cosf(theta), sinTheta = sinf(theta);
    float3 worldDir = f3(cosTheta * sinPhi, sinTheta * sinPhi, cosPhi);
    Scatter r = integrate_scattered<true, false>(A, luts, worldPos, worldDir, sunDir, true, 20.0f, false, false,
                                                 9000000.0f, f3s(1.0f));
    shMS[z] = (r.multiScatAs1 * sphereSolidAngle) / (sqrtSample * sqrtSample);
    shL[z] = (r.L * sphereSolidAngle) / (sqrtSample * sqrtSample);
    __syncthreads();
    for (int stride = 32; stride >= 1; stride >>= 1) {
        if (z < stride) {
            shMS[z] = shMS[z] + shMS[z + stride];
            shL[z] = shL[z] + shL[z + stride];
        }
        __syncthreads();
    }
    if (z > 0) return;
    float3 ms1 = shMS[0] * isotropicPhase;
    float3 inL = shL[0] * isotropicPhase;
    float3 sq = ms1 * ms1;
    float3 series = (((f3s(1.0f) + ms1) + sq) + ms1 * sq) + sq * sq;  // 5-term series (:133-136)
    store_rgba16f(out16, outf, y * MRT_MULTI_W + x, inL * series);
}

// genView.comp:33-77
__global__ void k_gen_view(mrt_atmosphere_params A, SkyLuts luts, float3 probePos, float3 sunDirection, float3 sunIll,
                           uint32_t* outp, float4* outf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= MRT_VIEW_W || y >= MRT_VIEW_H) return;
    float3 worldPos = probePos + f3(0.0f, 0.0f, A.bottomRadius);
    float u = (float)x / (float)MRT_VIEW_W, v = (float)y / (float)MRT_VIEW_H;
    float viewHeight = length3(worldPos);
    // skyAccess.glsl:54-85
    u = sky_sub_uv_to_unit(u, (float)MRT_VIEW_W);
    v = sky_sub_uv_to_unit(v, (float)MRT_VIEW_H);
    float vHorizon = sqrtf(viewHeight * viewHeight - A.bottomRadius * A.bottomRadius);
    float cosBeta = vHorizon / viewHeight;
    float beta = acosf(cosBeta);
    float zenithHorizonAngle = SKY_PI - beta;
    float cosZen;
    if (v < 0.5f) {
        float coord = 2.0f * v;
        coord = 1.0f - coord;
        coord *= coord;
        coord = 1.0f - coord;
        cosZen = cosf(zenithHorizonAngle * coord);
    } else {
        float coord = v * 2.0f - 1.0f;
        coord *= coord;
        cosZen = cosf(zenithHorizonAngle + beta * coord);
    }
    float cu = u * u;
    float lightViewCos = -(cu * 2.0f - 1.0f);

    float3 up = worldPos / viewHeight;
    float sunZen = dot3(up, sunDirection);
    float3 sunDir = normalize3(f3(sqrtf(1.0f - sunZen * sunZen), 0.0f, sunZen));
    worldPos = f3(0.0f, 0.0f, viewHeight);
    float sinZen = sqrtf(1.0f - cosZen * cosZen);
    float3 worldDir = f3(sinZen * lightViewCos, sinZen * sqrtf(1.0f - lightViewCos * lightViewCos), cosZen);

    float3 Lout = f3s(0.0f);
    // sky.glsl:79-94 moveToTopAtmosphere
    bool ok = true;
    if (viewHeight > A.topRadius) {
        float tTop = sky_ray_sphere_nearest(worldPos, worldDir, f3s(0.0f), A.topRadius);
        if (tTop >= 0.0f) {
            float3 upv = worldPos / viewHeight;
            worldPos = (worldPos + worldDir * tTop) + upv * -SKY_PLANET_RADIUS_OFFSET;
        } else {
            ok = false;
        }
    }
    if (ok) {
        Scatter ss = integrate_scattered<true, true>(A, luts, worldPos, worldDir, sunDir, false, 30.0f, true, true,
                                                     9000000.0f, sunIll);
        Lout = ss.L;
    }
    uint32_t p = pack_b10g11r11(Lout);
    outp[y * MRT_VIEW_W + x] = p;
    float3 d = unpack_b10g11r11(p);
    outf[y * MRT_VIEW_W + x] = make_float4(d.x, d.y, d.z, 1.0f);
}

// Aerial-perspective camera volume (the reference declares it -- sky.ixx:190-191, skyAccess.glsl:9,119-125 --
// and never builds it).  Contract and derivation: oracle/minote_oracle.c orc_gen_aerial_perspective.  One thread per
// froxel; RGBA16F (scattered luminance, 1 - mean transmittance) + a decoded float4 copy for the trilinear taps.
__global__ void __launch_bounds__(256) k_gen_aerial(mrt_atmosphere_params A, SkyLuts luts, RayGen gen, float3 cameraPos, float3 sunDirection,
                                                    float3 sunIll, uint16_t* out16, float4* outf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MRT_AERIAL_SIZE * MRT_AERIAL_SIZE * MRT_AERIAL_SIZE) return;
    const int x = i % MRT_AERIAL_SIZE, y = (i / MRT_AERIAL_SIZE) % MRT_AERIAL_SIZE, z = i / (MRT_AERIAL_SIZE * MRT_AERIAL_SIZE);
    float3 o, worldDir;
    ray_gen(gen, (uint32_t)x, (uint32_t)y, o, worldDir);
    const float3 camPos = cameraPos + f3(0.0f, 0.0f, A.bottomRadius);
    float slice = ((float)z + 0.5f) / (float)MRT_AERIAL_SIZE;
    slice *= slice;
    slice *= (float)MRT_AERIAL_SIZE;
    float3 worldPos = camPos;
    float tMax = slice * MRT_AERIAL_KM_PER_SLICE;
    float3 newWorldPos = worldPos + worldDir * tMax;
    float viewHeight = length3(newWorldPos);
    if (viewHeight <= A.bottomRadius + SKY_PLANET_RADIUS_OFFSET) {
        newWorldPos = normalize3(newWorldPos) * (A.bottomRadius + SKY_PLANET_RADIUS_OFFSET + 0.001f);
        worldDir = normalize3(newWorldPos - camPos);
        tMax = length3(newWorldPos - camPos);
    }
    float tMaxMax = tMax;
    float4 rgba = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    bool inside = true;
    viewHeight = length3(worldPos);
    if (viewHeight >= A.topRadius) {
        const float3 prev = worldPos;
        if (viewHeight > A.topRadius) {  // sky.glsl:79-94 moveToTopAtmosphere
            float tTop = sky_ray_sphere_nearest(worldPos, worldDir, f3s(0.0f), A.topRadius);
            if (tTop >= 0.0f) {
                float3 upv = worldPos / viewHeight;
                worldPos = (worldPos + worldDir * tTop) + upv * -SKY_PLANET_RADIUS_OFFSET;
            } else {
                inside = false;
            }
        }
        if (inside) {
            float lengthToAtmosphere = length3(prev - worldPos);
            if (tMaxMax < lengthToAtmosphere) inside = false;
            tMaxMax = fmaxf(0.0f, tMaxMax - lengthToAtmosphere);
        }
    }
    if (inside) {
        const float sampleCountIni = fmaxf(1.0f, ((float)z + 1.0f) * 2.0f);
        Scatter ss = integrate_scattered<true, true>(A, luts, worldPos, worldDir, sunDirection, false, sampleCountIni, false, true,
                                                     tMaxMax, sunIll);
        const float T = (ss.transmittance.x + ss.transmittance.y + ss.transmittance.z) * (1.0f / 3.0f);
        rgba = make_float4(ss.L.x, ss.L.y, ss.L.z, 1.0f - T);
    }
    const uint16_t h[4] = {f32_to_f16_bits(rgba.x), f32_to_f16_bits(rgba.y), f32_to_f16_bits(rgba.z), f32_to_f16_bits(rgba.w)};
    for (int c = 0; c < 4; c++) out16[4 * (size_t)i + c] = h[c];
    outf[i] = make_float4(f16_bits_to_f32(h[0]), f16_bits_to_f32(h[1]), f16_bits_to_f32(h[2]), f16_bits_to_f32(h[3]));
}

}  // namespace

int sky_gen_aerial(mrt_context* ctx, const mrt_primary_constants* c, const float cameraPos[3], const float sunDir[3], const float sunIll[3]) {
    const size_t n = (size_t)MRT_AERIAL_SIZE * MRT_AERIAL_SIZE * MRT_AERIAL_SIZE;
    MRT_TRY(dev_reserve(ctx, ctx->aerial16, 4 * n));
    MRT_TRY(dev_reserve(ctx, ctx->aerial_f, n));
    RayGen gen;
    memcpy(&gen.invView, &c->invView, sizeof(Mat4));
    memcpy(&gen.invProj, &c->invProjection, sizeof(Mat4));
    gen.W = gen.H = MRT_AERIAL_SIZE;
    SkyLuts luts{ctx->trans_f.p, ctx->multi_f.p, nullptr};
    k_gen_aerial<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->atmo, luts, gen, f3(cameraPos[0], cameraPos[1], cameraPos[2]),
                                                          f3(sunDir[0], sunDir[1], sunDir[2]), f3(sunIll[0], sunIll[1], sunIll[2]),
                                                          ctx->aerial16.p, ctx->aerial_f.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "sky_gen_aerial");
}

int sky_gen_atmosphere(mrt_context* ctx) {
    MRT_TRY(dev_reserve(ctx, ctx->trans16, (size_t)MRT_TRANS_W * MRT_TRANS_H * 4));
    MRT_TRY(dev_reserve(ctx, ctx->trans_f, (size_t)MRT_TRANS_W * MRT_TRANS_H));
    MRT_TRY(dev_reserve(ctx, ctx->multi16, (size_t)MRT_MULTI_W * MRT_MULTI_H * 4));
    MRT_TRY(dev_reserve(ctx, ctx->multi_f, (size_t)MRT_MULTI_W * MRT_MULTI_H));
    dim3 b(8, 8), g(MRT_TRANS_W / 8, MRT_TRANS_H / 8);
    k_gen_transmittance<<<g, b, 0, ctx->stream>>>(ctx->atmo, ctx->trans16.p, ctx->trans_f.p);
    MRT_LAUNCHED(ctx);
    SkyLuts luts{ctx->trans_f.p, nullptr, nullptr};
    k_gen_multiscattering<<<dim3(MRT_MULTI_W, MRT_MULTI_H), 64, 0, ctx->stream>>>(ctx->atmo, luts, ctx->multi16.p,
                                                                                 ctx->multi_f.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "sky_gen_atmosphere");
}

int sky_gen_view(mrt_context* ctx, const float probe[3], const float sunDir[3], const float sunIll[3]) {
    MRT_TRY(dev_reserve(ctx, ctx->view_packed, (size_t)MRT_VIEW_W * MRT_VIEW_H));
    MRT_TRY(dev_reserve(ctx, ctx->view_f, (size_t)MRT_VIEW_W * MRT_VIEW_H));
    SkyLuts luts{ctx->trans_f.p, ctx->multi_f.p, nullptr};
    dim3 b(8, 8), g(div_up(MRT_VIEW_W, 8), div_up(MRT_VIEW_H, 8));
    k_gen_view<<<g, b, 0, ctx->aux_stream>>>(ctx->atmo, luts, f3(probe[0], probe[1], probe[2]),
                                         f3(sunDir[0], sunDir[1], sunDir[2]), f3(sunIll[0], sunIll[1], sunIll[2]),
                                         ctx->view_packed.p, ctx->view_f.p);
    MRT_LAUNCHED(ctx);
    return mrt_check_cuda(ctx, cudaGetLastError(), "sky_gen_view");
}
