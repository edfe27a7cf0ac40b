// oracle/_ref: src/gpu/sky/genTransmittance.comp (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
#include "ref_sky.inc"
#define GLSL_SPEC_CONSTANT_0 256   /* cmd.specialize_constants(0, TransmittanceSize.x()), sky.ixx */
#define GLSL_SPEC_CONSTANT_1 64
namespace glsl { namespace { namespace sh {
#include "sky/genTransmittance.comp"
}}}
using namespace glsl;
extern "C" void ref_gen_transmittance(const void* atmo144, uint16_t* rgba16f) {
    std::memcpy((void*)&sh::u_atmo, atmo144, sizeof(sh::AtmosphereParams));
    sh::i_transmittance = Image{rgba16f, TRANS_W, TRANS_H, RGBA16F};
    dispatch_invocations(TRANS_W, TRANS_H, sh::shader_main);
}
