// oracle/_ref: src/gpu/sky/genMultiScattering.comp -- work groups of 64 invocations with shared memory and
// barrier(), run as 64 host threads (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
#include "ref_sky.inc"
#define GLSL_SPEC_CONSTANT_0 32
#define GLSL_SPEC_CONSTANT_1 32
namespace glsl { namespace { namespace sh {
#include "sky/genMultiScattering.comp"
}}}
using namespace glsl;
extern "C" void ref_gen_multiscattering(const void* atmo144, const uint16_t* trans, uint16_t* rgba16f) {
    std::memcpy((void*)&sh::u_atmo, atmo144, sizeof(sh::AtmosphereParams));
    sh::s_transmittance = Sampler{trans, TRANS_W, TRANS_H, RGBA16F, true, false, 0};   // LinearClamp, sky.ixx:161
    sh::i_multiScattering = Image{rgba16f, MULTI_W, MULTI_H, RGBA16F};
    dispatch_groups_z(MULTI_W, MULTI_H, sh::gl_WorkGroupSizeZ, sh::shader_main);        // sky.ixx dispatch of 32x32 groups
}
