// minote.modules.pathtracer -- Pathtracer::primaryRays / ::secondaryRays
// (src/gfx/modules/pathtracer.ixx:29-116, 118-195), same names, argument order and constants structs.
// The UBO structs are filled exactly as pathtracer.ixx:94-104 / :178-188 do and handed to the C ABI by
// pointer; matrices are never recomputed on the device.
module;
#include <cstring>

#include "../../include/minotert.h"
export module minote.modules.pathtracer;
import minote.math;
import minote.camera;
import minote.cuda;
import minote.modules.sky;

export class Pathtracer : Cuda {
public:
    struct GBuffer {
        DeviceImage visibility;  // R32_UINT
        DeviceImage depth;       // R16F
        DeviceImage normal;      // RGBA16F
        DeviceImage motion;      // RG16F
    };

    // Samples / Bounces are compile-time 8 / 8 in the reference (secondaryRays.comp:128-129);
    // here they are renderer settings with the same defaults.
    u32 samples = 8;
    u32 bounces = 8;
    bool accumulate = false;  // progressive accumulation across draw() calls (row n7)
    // Sky extensions the reference stubs out (sky.ixx:190-191, secondaryRays.comp:37; SURVEY 8f-4), triangle scenes:
    bool sunSampling = false; // the sun as a sampled light: a shadow ray into its disc at every hit that bounces
    bool skyAtHit = false;    // sky evaluated at the shaded point instead of at the camera
    bool aerialPerspective = false;  // camera volume (Sky::createAerialPerspective) applied between camera and primary hit

    struct PrimaryConstants {
        mat4 view, projection, invView, invProjection, prevView;
        u32 frameCounter;
    };
    struct SecondaryConstants {
        mat4 view, projection, invView, invProjection;
        vec3 cameraPos;
        u32 frameCounter;
    };
    static_assert(sizeof(PrimaryConstants) == sizeof(mrt_primary_constants));
    static_assert(sizeof(SecondaryConstants) == sizeof(mrt_secondary_constants));

    static auto primaryConstants(Camera const& camera, Camera const& prevCamera, u32 frame) -> PrimaryConstants {
        mat4 const view = camera.view();
        mat4 const projection = camera.projection();
        return PrimaryConstants{view, projection, inverse(view), inverse(projection), prevCamera.view(), frame};
    }
    static auto secondaryConstants(Camera const& camera, u32 frame) -> SecondaryConstants {
        mat4 const view = camera.view();
        mat4 const projection = camera.projection();
        return SecondaryConstants{view, projection, inverse(view), inverse(projection), camera.position, frame};
    }

    auto primaryRays(uvec2 size, Camera const& camera, Camera const& prevCamera) -> GBuffer {
        auto const constants = primaryConstants(camera, prevCamera, Cuda::serv->frameCount());
        mrt_primary_constants raw;
        std::memcpy(&raw, &constants, sizeof raw);
        Cuda::serv->check(mrt_primary_rays(Cuda::serv->ctx, size.x(), size.y(), &raw));
        return GBuffer{{MRT_BUF_VISIBILITY}, {MRT_BUF_DEPTH}, {MRT_BUF_NORMAL}, {MRT_BUF_MOTION}};
    }

    auto secondaryRays(GBuffer gbuffer, Camera const& camera, Atmosphere const& atmo, DeviceImage skyView,
                       DeviceImage blueNoise) -> DeviceImage {
        (void)gbuffer; (void)atmo; (void)skyView; (void)blueNoise;  // resident in the context; passed for interface parity
        auto const constants = secondaryConstants(camera, Cuda::serv->frameCount());
        mrt_secondary_constants raw;
        std::memcpy(&raw, &constants, sizeof raw);
        u32 const flags = (accumulate ? MRT_SECONDARY_ACCUMULATE : 0u) | (sunSampling ? MRT_SECONDARY_NEE_SUN : 0u) |
                                    (skyAtHit ? MRT_SECONDARY_SKY_AT_HIT : 0u) | (aerialPerspective ? MRT_SECONDARY_AERIAL : 0u);
        Cuda::serv->check(mrt_secondary_rays(Cuda::serv->ctx, &raw, samples, bounces, flags));
        return DeviceImage{accumulate ? MRT_BUF_ACCUM : MRT_BUF_COLOR};
    }
};
