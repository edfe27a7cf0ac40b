/*
 * minote_ref.h -- C entry points of oracle/_ref/libminote_ref.so: the REFERENCE'S OWN GLSL compute shaders
 * (Tearnote/MinoteRT src/gpu, read from /root/reference at build time, never committed) compiled as C++ through
 * oracle/ref/glsl_shim.hpp and run on host arrays.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  It pins oracle/minote_oracle.c (the restatement):
 * tests/test_ref_pins_oracle.py runs both on the same inputs.  Image layouts = the reference's formats
 * (src/gfx/modules/pathtracer.ixx:42-69,139-145; sky.ixx:22-26,187-188; tonemapper.ixx:73; denoiser.ixx:54).
 * Not reentrant: shader uniforms are globals, one call at a time.
 */
#ifndef MINOTE_REF_H
#define MINOTE_REF_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* src/gpu/random.glsl:10-31, src/gpu/intersect.glsl:26-37 */
uint32_t ref_pcg(uint32_t* state);
float ref_random_float(uint32_t* state);
void ref_random_sphere_point(float rx, float ry, float out[3]);
/* sphere = {center[3], radius, albedo[3]} */
float ref_ray_sphere(const float o[3], const float d[3], const float sphere[7]);
/* the scene compiled into the shaders (src/gpu/scene.glsl:4-11): n spheres x 7 floats; returns the count */
uint32_t ref_scene_spheres(float* out7, uint32_t max_spheres);

/* src/gpu/primaryRay.comp (constants = the 324-byte block of pathtracer.ixx:86-104) */
void ref_primary_rays(uint32_t w, uint32_t h, const void* constants324, uint32_t* visibility, uint16_t* depth,
                      uint16_t* normal /*4/px*/, uint16_t* motion /*2/px*/);
/* src/gpu/secondaryRays.comp: 8 samples x 8 bounces as compiled in (constants = the 272-byte block of
 * pathtracer.ixx:170-188; atmo = the 144-byte sky.ixx:28-56 block) */
void ref_secondary_rays(uint32_t w, uint32_t h, const void* constants272, const void* atmo144,
                        const uint32_t* visibility, const uint16_t* depth, const uint16_t* normal,
                        const uint8_t* blueNoise, uint32_t bnW, uint32_t bnH, const uint16_t* trans,
                        const uint32_t* skyView, uint16_t* color16);
/* skyColor() of src/gpu/secondaryRays.comp:36-58 for n directions (dir: n x 3, out: n x 3) */
void ref_sky_color(const void* atmo144, const uint16_t* trans, const uint32_t* skyView, const float cameraPos[3],
                   uint32_t n, const float* dir, float* out);

/* src/gpu/sky/gen{Transmittance,MultiScattering,View}.comp at the sizes of sky.ixx:22-26,187-188 */
void ref_gen_transmittance(const void* atmo144, uint16_t* rgba16f /*256*64*4*/);
void ref_gen_multiscattering(const void* atmo144, const uint16_t* trans, uint16_t* rgba16f /*32*32*4*/);
void ref_gen_sky_view(const void* atmo144, const uint16_t* trans, const uint16_t* multi, const float probePos[3],
                      const float sunDir[3], const float sunIlluminance[3], uint32_t* b10g11r11 /*192*108*/);

/* src/gpu/tonemap/{linear,reinhard,hable,aces,uchimura,amd}.comp; mode 0..5 in that order.
 * src_fmt: 0 RGBA32F, 1 RGBA16F (what the reference binds), 2 RGBA8 unorm (the denoiser's output).
 * params = the push constants after exposure (reinhard 1, uchimura 6, amd 5 floats). */
void ref_tonemap(int mode, uint32_t w, uint32_t h, const void* src, int src_fmt, float exposure, const float* params,
                 uint8_t* rgba8);
/* src/gpu/denoise/bilateral.comp */
void ref_denoise_bilateral(uint32_t w, uint32_t h, const uint16_t* color16, const uint16_t* depth16,
                           const uint16_t* normal16, float sigma, float kSigma, float threshold, float nearPlane,
                           uint32_t frameCounter, uint8_t* rgba8);

/* ---- host side: src/stx/math.ixx + src/gfx/camera.ixx + Atmosphere::Params (sky.ixx:28-84), compiled from the
 * reference's text with the module syntax stripped (oracle/ref/ixx2hpp.py) ---- */
typedef struct {
    uint32_t viewport[2];
    float verticalFov, nearPlane;
    float position[3];
    float yaw, pitch, lookSpeed, moveSpeed;
} ref_camera; /* src/gfx/camera.ixx:8-22 */
void ref_primary_constants_fill(const ref_camera* cam, const ref_camera* prev, uint32_t frame, void* out324);
void ref_secondary_constants_fill(const ref_camera* cam, uint32_t frame, void* out272);
void ref_camera_direction(const ref_camera* cam, float out[3]);
void ref_camera_rotate(ref_camera* cam, float horz, float vert);
void ref_camera_shift(ref_camera* cam, const float d[3]);
void ref_camera_roam(ref_camera* cam, const float d[3]);
void ref_perspective(float vFov, float aspect, float zNear, float out16[16]);
void ref_look(const float pos[3], const float dir[3], const float up[3], float out16[16]);
void ref_inverse(const float in16[16], float out16[16]);
void ref_mat_mul(const float a16[16], const float b16[16], float out16[16]);
float ref_deg(float degrees);
void ref_atmosphere_earth(void* out144);

void ref_set_num_threads(int n);
int ref_num_threads(void);
#ifdef __cplusplus
}
#endif
#endif
