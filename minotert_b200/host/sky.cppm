// minote.modules.sky -- Atmosphere + Sky (src/gfx/modules/sky.ixx), same names and arguments;
// the vuk passes become mrt_atmosphere / mrt_sky_view.
module;
#include <algorithm>
#include <cstring>

#include "../../include/minotert.h"
export module minote.modules.sky;
import minote.math;
import minote.cuda;

export class Atmosphere : Cuda {
public:
    // std140 mirror of the shader struct, 144 bytes (sky.ixx:28-56)
    struct Params {
        float bottomRadius;
        float topRadius;
        float rayleighDensityExpScale;
        float _pad0;
        vec3 rayleighScattering;
        float mieDensityExpScale;
        vec3 mieScattering;
        float _pad1;
        vec3 mieExtinction;
        float _pad2;
        vec3 mieAbsorption;
        float miePhaseG;
        float absorptionDensity0LayerWidth;
        float absorptionDensity0ConstantTerm;
        float absorptionDensity0LinearTerm;
        float absorptionDensity1ConstantTerm;
        float absorptionDensity1LinearTerm;
        float _pad3, _pad4, _pad5;
        vec3 absorptionExtinction;
        float _pad6;
        vec3 groundAlbedo;
        float _pad7;

        // Earth (sky.ixx:59-83)
        static auto earth() -> Params {
            Params p{};
            constexpr float rayleighScaleHeight = 8.0f, mieScaleHeight = 1.2f;
            p.bottomRadius = 6360.0f;
            p.topRadius = 6460.0f;
            p.rayleighDensityExpScale = -1.0f / rayleighScaleHeight;
            p.rayleighScattering = {0.005802f, 0.013558f, 0.033100f};
            p.mieDensityExpScale = -1.0f / mieScaleHeight;
            p.mieScattering = {0.003996f, 0.003996f, 0.003996f};
            p.mieExtinction = {0.004440f, 0.004440f, 0.004440f};
            for (int i = 0; i < 3; i++) p.mieAbsorption[i] = std::max(p.mieExtinction[i] - p.mieScattering[i], 0.0f);
            p.miePhaseG = 0.8f;
            p.absorptionDensity0LayerWidth = 25.0f;
            p.absorptionDensity0ConstantTerm = -2.0f / 3.0f;
            p.absorptionDensity0LinearTerm = 1.0f / 15.0f;
            p.absorptionDensity1ConstantTerm = 8.0f / 3.0f;
            p.absorptionDensity1LinearTerm = -1.0f / 15.0f;
            p.absorptionExtinction = {0.000650f, 0.001881f, 0.000085f};
            p.groundAlbedo = {0.0f, 0.0f, 0.0f};
            return p;
        }
    };
    static_assert(sizeof(Params) == sizeof(mrt_atmosphere_params));

    DeviceImage transmittance;
    DeviceImage multiScattering;
    Params params;

    // Create and precalculate the atmosphere data (sky.ixx:91-176)
    explicit Atmosphere(Params const& p) : params(p) {
        mrt_atmosphere_params raw;
        std::memcpy(&raw, &p, sizeof raw);
        Cuda::serv->check(mrt_atmosphere(Cuda::serv->ctx, &raw));
        transmittance.id = MRT_BUF_TRANSMITTANCE;
        multiScattering.id = MRT_BUF_MULTISCATTERING;
    }
};

export class Sky : Cuda {
public:
    vec3 sunDirection = {-0.435286462f, 0.818654716f, 0.374606609f};  // sky.ixx:193
    vec3 sunIlluminance = {8.0f, 8.0f, 8.0f};                         // sky.ixx:194

    // 360-degree sky view LUT at probePos (sky.ixx:199-262)
    auto createView(Atmosphere const&, vec3 probePos) -> DeviceImage {
        Cuda::serv->check(mrt_sky_view(Cuda::serv->ctx, probePos.v.data(), sunDirection.v.data(), sunIlluminance.v.data()));
        return DeviceImage{MRT_BUF_SKY_VIEW};
    }

    // The camera volume the reference declares and never builds (AerialPerspectiveFormat / AerialPerspectiveSize,
    // sky.ixx:190-191): luminance scattered towards the camera + 1 - transmittance in 32^3 froxels of the view whose
    // inverse matrices are given (the Pathtracer's primary constants); SURVEY 8f-4
    static constexpr unsigned AerialPerspectiveSize = 32;
    auto createAerialPerspective(Atmosphere const&, mrt_primary_constants const& constants, vec3 cameraPos) -> DeviceImage {
        Cuda::serv->check(mrt_sky_aerial_perspective(Cuda::serv->ctx, &constants, cameraPos.v.data(), sunDirection.v.data(),
                                                     sunIlluminance.v.data()));
        return DeviceImage{MRT_BUF_AERIAL};
    }
};
