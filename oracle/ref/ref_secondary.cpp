// oracle/_ref: src/gpu/secondaryRays.comp on host arrays (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
namespace glsl { namespace { namespace sh {
#include "secondaryRays.comp"
static_assert(sizeof(C) == 272, "secondary Constants block (pathtracer.ixx:170-177)");
static_assert(sizeof(AtmosphereParams) == 140, "AtmosphereParams must follow the std140 offsets of sky.ixx:28-56");
}}}
using namespace glsl;
static void bind_sky(const void* atmo144, const uint16_t* trans, const uint32_t* skyView) {
    std::memcpy((void*)&sh::u_atmo, atmo144, sizeof(sh::AtmosphereParams));
    // samplers of pathtracer.ixx:160-165
    sh::s_transmittance = Sampler{trans, 256, 64, RGBA16F, true, false, 0};      // LinearClamp
    sh::s_skyView = Sampler{skyView, 192, 108, B10G11R11, true, true, 0};        // LinearRepeat
}
extern "C" void ref_secondary_rays(uint32_t w, uint32_t h, const void* constants272, const void* atmo144,
                                   const uint32_t* visibility, const uint16_t* depth, const uint16_t* normal,
                                   const uint8_t* blueNoise, uint32_t bnW, uint32_t bnH, const uint16_t* trans,
                                   const uint32_t* skyView, uint16_t* color16) {
    std::memcpy((void*)&sh::C, constants272, 272);
    bind_sky(atmo144, trans, skyView);
    static_cast<Sampler&>(sh::s_visibility) = Sampler{visibility, (int)w, (int)h, R32UI, false, false, 0};
    sh::s_depth = Sampler{depth, (int)w, (int)h, R16F, false, false, 0};
    sh::s_normal = Sampler{normal, (int)w, (int)h, RGBA16F, false, false, 0};
    sh::s_blueNoise = Sampler{blueNoise, (int)bnW, (int)bnH, RGBA8, false, false, 0};
    sh::i_color = Image{color16, (int)w, (int)h, RGBA16F};
    dispatch_invocations(w, h, sh::shader_main);   // pathtracer.ixx:190
}
extern "C" void ref_sky_color(const void* atmo144, const uint16_t* trans, const uint32_t* skyView,
                              const float cameraPos[3], uint32_t n, const float* dir, float* out) {
    bind_sky(atmo144, trans, skyView);
    sh::C.cameraPos = vec3{cameraPos[0], cameraPos[1], cameraPos[2]};
    for (uint32_t i = 0; i < n; i++) {
        vec3 c = sh::skyColor(vec3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
