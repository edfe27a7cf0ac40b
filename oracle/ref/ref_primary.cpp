// oracle/_ref: src/gpu/primaryRay.comp on host arrays (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
namespace glsl { namespace { namespace sh {
#include "intersect.glsl"
// primaryRay.comp:33-34 reads Spheres[primitiveId] with primitiveId == -1u on a miss: undefined in GLSL, a wild host
// read here.  Robust access: out-of-range indices read element 0.  The values derived from it (position, depth and
// motion of MISS pixels) are garbage in the reference too and are masked out of every comparison; the normal is
// overwritten with the ray direction (:69-70) and the id is -1u.
#define Spheres Spheres_storage
#include "scene.glsl"
#undef Spheres
static constexpr struct SpheresRobust {
    const Sphere& operator[](uint i) const { return Spheres_storage[i < SphereCount ? i : 0u]; }
} Spheres{};
#include "primaryRay.comp"
static_assert(sizeof(C) == 324, "primary Constants block (pathtracer.ixx:86-93)");
}}}
using namespace glsl;
extern "C" void ref_primary_rays(uint32_t w, uint32_t h, const void* constants324, uint32_t* visibility,
                                 uint16_t* depth, uint16_t* normal, uint16_t* motion) {
    std::memcpy((void*)&sh::C, constants324, 324);
    sh::i_visibility.data = visibility; sh::i_visibility.w = w; sh::i_visibility.h = h; sh::i_visibility.fmt = R32UI;
    sh::i_depth = Image{depth, (int)w, (int)h, R16F};
    sh::i_normal = Image{normal, (int)w, (int)h, RGBA16F};
    sh::i_motion = Image{motion, (int)w, (int)h, RG16F};
    dispatch_invocations(w, h, sh::shader_main);   // cmd.dispatch_invocations(size.x(), size.y()), pathtracer.ixx:106
}
