// oracle/_ref: tonemap dispatcher + src/gpu/denoise/bilateral.comp (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
namespace glsl { namespace { namespace sh {
#include "denoise/bilateral.comp"
static_assert(sizeof(C) == 20, "bilateral push constants (denoiser.ixx:78-91)");
}}}
using namespace glsl;
extern "C" {
#define TM(name) void ref_tonemap_##name(uint32_t, uint32_t, const void*, int, float, const float*, uint8_t*);
TM(linear) TM(reinhard) TM(hable) TM(aces) TM(uchimura) TM(amd)
#undef TM
void ref_tonemap(int mode, uint32_t w, uint32_t h, const void* src, int src_fmt, float exposure, const float* params,
                 uint8_t* rgba8) {
    switch (mode) {
    case 0: ref_tonemap_linear(w, h, src, src_fmt, exposure, params, rgba8); break;
    case 1: ref_tonemap_reinhard(w, h, src, src_fmt, exposure, params, rgba8); break;
    case 2: ref_tonemap_hable(w, h, src, src_fmt, exposure, params, rgba8); break;
    case 3: ref_tonemap_aces(w, h, src, src_fmt, exposure, params, rgba8); break;
    case 4: ref_tonemap_uchimura(w, h, src, src_fmt, exposure, params, rgba8); break;
    default: ref_tonemap_amd(w, h, src, src_fmt, exposure, params, rgba8); break;
    }
}
void ref_denoise_bilateral(uint32_t w, uint32_t h, const uint16_t* color16, const uint16_t* depth16,
                           const uint16_t* normal16, float sigma, float kSigma, float threshold, float nearPlane,
                           uint32_t frameCounter, uint8_t* rgba8) {
    // LinearClamp on all three (denoiser.ixx:71-73).  Sub-texel precision 8 bits, zero-weight texels not read: the
    // shader's taps sit on texel centres in x and the colour image holds +inf on the sun disc, see glsl_shim.hpp.
    sh::s_color = Sampler{color16, (int)w, (int)h, RGBA16F, true, false, 8};
    sh::s_depth = Sampler{depth16, (int)w, (int)h, R16F, true, false, 8};
    sh::s_normal = Sampler{normal16, (int)w, (int)h, RGBA16F, true, false, 8};
    sh::i_dst = Image{rgba8, (int)w, (int)h, RGBA8};
    sh::C.sigma = sigma; sh::C.kSigma = kSigma; sh::C.threshold = threshold; sh::C.nearPlane = nearPlane;
    sh::C.frameCounter = frameCounter;
    dispatch_invocations(w, h, sh::shader_main);
}
}
