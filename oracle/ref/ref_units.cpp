// oracle/_ref: unit-level entry points into the reference's random.glsl / intersect.glsl / scene.glsl (TEST INFRASTRUCTURE)
#include "glsl_shim.hpp"
#include "minote_ref.h"
#include <omp.h>
namespace glsl { namespace { namespace sh {
#include "intersect.glsl"
#include "random.glsl"
#include "scene.glsl"
}}}
using namespace glsl;
extern "C" {
uint32_t ref_pcg(uint32_t* state) { return sh::pcg(*state); }
float ref_random_float(uint32_t* state) { return sh::randomFloat(*state); }
void ref_random_sphere_point(float rx, float ry, float out[3]) {
    vec3 p = sh::randomSpherePoint(vec2{rx, ry});
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
float ref_ray_sphere(const float o[3], const float d[3], const float s[7]) {
    sh::Ray ray; ray.origin = vec3{o[0], o[1], o[2]}; ray.direction = vec3{d[0], d[1], d[2]};
    sh::Sphere sp; sp.center = vec3{s[0], s[1], s[2]}; sp.radius = s[3]; sp.albedo = vec3{s[4], s[5], s[6]};
    return sh::raySphereIntersect(ray, sp);
}
uint32_t ref_scene_spheres(float* out7, uint32_t max_spheres) {
    for (uint32_t i = 0; i < sh::SphereCount && i < max_spheres; i++) {
        const sh::Sphere& s = sh::Spheres[i];
        float v[7] = {s.center.x, s.center.y, s.center.z, s.radius, s.albedo.x, s.albedo.y, s.albedo.z};
        for (int k = 0; k < 7; k++) out7[7 * i + k] = v[k];
    }
    return sh::SphereCount;
}
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_num_threads(void) { return omp_get_max_threads(); }
}
