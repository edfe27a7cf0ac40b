// oracle/_ref: src/gpu/sky/genView.comp (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
#include "ref_sky.inc"
#define GLSL_SPEC_CONSTANT_0 192
#define GLSL_SPEC_CONSTANT_1 108
namespace glsl { namespace { namespace sh {
#include "sky/genView.comp"
}}}
using namespace glsl;
extern "C" void ref_gen_sky_view(const void* atmo144, const uint16_t* trans, const uint16_t* multi,
                                 const float probePos[3], const float sunDir[3], const float sunIlluminance[3],
                                 uint32_t* b10g11r11) {
    std::memcpy((void*)&sh::u_atmo, atmo144, sizeof(sh::AtmosphereParams));
    sh::s_transmittance = Sampler{trans, TRANS_W, TRANS_H, RGBA16F, true, false, 0};      // sky.ixx:235
    sh::s_multiscattering = Sampler{multi, MULTI_W, MULTI_H, RGBA16F, true, false, 0};    // sky.ixx:236
    sh::i_view = Image{b10g11r11, VIEW_W, VIEW_H, B10G11R11};
    sh::C.probePos = vec3{probePos[0], probePos[1], probePos[2]};                         // push constants, sky.ixx:239-250
    sh::C.sunDirection = vec3{sunDir[0], sunDir[1], sunDir[2]};
    sh::C.sunIlluminance = vec3{sunIlluminance[0], sunIlluminance[1], sunIlluminance[2]};
    dispatch_invocations(VIEW_W, VIEW_H, sh::shader_main);
}
