// oracle/_ref: the reference's HOST math and camera (src/stx/math.ixx, src/gfx/camera.ixx, Atmosphere::Params of
// src/gfx/modules/sky.ixx), module syntax stripped by ixx2hpp.py, everything else their own text.  TEST INFRASTRUCTURE.
// Pins SURVEY row a1: the matrices and constant blocks the ray generation depends on bit for bit.
#include "minote.camera.hpp"
#include "sky_params.hpp"
#include "minote_ref.h"
#include <cstring>

static_assert(sizeof(mat4) == 64 && sizeof(vec3) == 12, "column-major packed floats (math.ixx:531-539)");
static_assert(sizeof(Params) == 140, "Atmosphere::Params: the 144-byte std140 block minus the tail pad");

static Camera camera_from(const ref_camera* c) {
    Camera cam{};
    cam.viewport = uvec2{c->viewport[0], c->viewport[1]};
    cam.verticalFov = c->verticalFov; cam.nearPlane = c->nearPlane;
    cam.position = vec3{c->position[0], c->position[1], c->position[2]};
    cam.yaw = c->yaw; cam.pitch = c->pitch; cam.lookSpeed = c->lookSpeed; cam.moveSpeed = c->moveSpeed;
    return cam;
}
static void camera_to(const Camera& cam, ref_camera* c) {
    c->position[0] = cam.position.x(); c->position[1] = cam.position.y(); c->position[2] = cam.position.z();
    c->yaw = cam.yaw; c->pitch = cam.pitch;
}

extern "C" {
// Pathtracer::primaryRays, src/gfx/modules/pathtracer.ixx:86-104 (the block layout and the fill, restated: the lines
// sit inside a vuk render-pass lambda)
void ref_primary_constants_fill(const ref_camera* cam, const ref_camera* prev, uint32_t frame, void* out324) {
    struct Constants { mat4 view, projection, invView, invProjection, prevView; uint frameCounter; };
    static_assert(sizeof(Constants) == 324);
    Camera camera = camera_from(cam), prevCamera = camera_from(prev);
    mat4 view = camera.view();
    mat4 projection = camera.projection();
    Constants c{view, projection, inverse(view), inverse(projection), prevCamera.view(), frame};
    std::memcpy(out324, &c, sizeof c);
}
// Pathtracer::secondaryRays, src/gfx/modules/pathtracer.ixx:170-188
void ref_secondary_constants_fill(const ref_camera* cam, uint32_t frame, void* out272) {
    struct Constants { mat4 view, projection, invView, invProjection; vec3 cameraPos; uint frameCounter; };
    static_assert(sizeof(Constants) == 272);
    Camera camera = camera_from(cam);
    mat4 view = camera.view();
    mat4 projection = camera.projection();
    Constants c{view, projection, inverse(view), inverse(projection), camera.position, frame};
    std::memcpy(out272, &c, sizeof c);
}
void ref_camera_direction(const ref_camera* cam, float out[3]) {
    vec3 d = camera_from(cam).direction();
    out[0] = d.x(); out[1] = d.y(); out[2] = d.z();
}
void ref_camera_rotate(ref_camera* cam, float horz, float vert) { Camera c = camera_from(cam); c.rotate(horz, vert); camera_to(c, cam); }
void ref_camera_shift(ref_camera* cam, const float d[3]) { Camera c = camera_from(cam); c.shift(vec3{d[0], d[1], d[2]}); camera_to(c, cam); }
void ref_camera_roam(ref_camera* cam, const float d[3]) { Camera c = camera_from(cam); c.roam(vec3{d[0], d[1], d[2]}); camera_to(c, cam); }
void ref_perspective(float vFov, float aspect, float zNear, float out16[16]) { mat4 m = perspective(vFov, aspect, zNear); std::memcpy(out16, &m, 64); }
void ref_look(const float pos[3], const float dir[3], const float up[3], float out16[16]) {
    mat4 m = look(vec3{pos[0], pos[1], pos[2]}, vec3{dir[0], dir[1], dir[2]}, vec3{up[0], up[1], up[2]});
    std::memcpy(out16, &m, 64);
}
void ref_inverse(const float in16[16], float out16[16]) { mat4 m; std::memcpy(&m, in16, 64); mat4 r = inverse(m); std::memcpy(out16, &r, 64); }
void ref_mat_mul(const float a16[16], const float b16[16], float out16[16]) {
    mat4 a, b; std::memcpy(&a, a16, 64); std::memcpy(&b, b16, 64); mat4 r = a * b; std::memcpy(out16, &r, 64);
}
float ref_deg(float degrees) { return radians(degrees); }   // the _deg literal, math.ixx:872-886
// Atmosphere::Params::earth(), src/gfx/modules/sky.ixx:59-83, into the 144-byte block (tail pad zero)
void ref_atmosphere_earth(void* out144) {
    Params p = Params::earth();
    std::memset(out144, 0, 144);
    std::memcpy(out144, &p, sizeof p);
}
}
