/*
 * minote_oracle.h -- CPU ORACLE for the MinoteRT ray-tracing hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (minotert_b200/csrc, libminotert.so) never links or calls it.
 *
 * PARITY PINNED (sphere path, sky, tonemap, denoiser, host matrices) against the reference's own code:
 * the reference ships no tests or golden vectors and its renderer cannot be built or run here (Win32 +
 * MSVC + Vulkan), but its GLSL compute shaders and its host math/camera modules DO compile as C++ --
 * oracle/ref/ (glsl2cpp.py + glsl_shim.hpp, ixx2hpp.py) builds them from /root/reference into
 * oracle/_ref/libminote_ref.so, and tests/test_ref_pins_oracle.py asserts that every function below that
 * restates reference code reproduces that library BIT FOR BIT (full 960x540 8x8 frame, the three sky LUTs,
 * six tonemappers, bilateral denoiser, 64 k-direction skyColor sweep, RNG stream, 2000 random cameras), and
 * reproduces the golden vectors generated from it (tests/golden/ref_v1.npz).  What stays pinned only by its
 * own contract: the triangle/BVH extensions (no reference counterpart, see the end of this comment).
 * This file is a plain-C restatement of the reference's GLSL + host matrix code, each function citing
 * the reference file:line it follows.  Every implementation-defined Vulkan behaviour is fixed explicitly
 * (identically in oracle/ref/glsl_shim.hpp):
 *   fp32 -> fp16          : IEEE round-to-nearest-even, overflow -> inf
 *   fp32 -> B10G11R11     : RNE to 6/6/5-bit mantissa, negatives/NaN -> 0, saturate to max finite
 *   fp32 -> unorm8        : rint(clamp(x,0,1)*255), NaN -> 0
 *   bilinear filtering    : fp32 lerp on texel-centre coordinates, clamp/repeat per sampler
 *                           (denoiser only: weights held to 8 fractional bits, zero-weight texels
 *                           not read -- see texn_bilinear in minote_oracle.c for why)
 *   mat4*vec4             : sum of columns scaled by components, left to right, no FMA
 *   normalize(v)          : v / sqrt(dot(v,v)) per component
 *   transcendentals       : glibc libm (sinf, cosf, acosf, powf, expf, sqrtf)
 * All arithmetic is fp32 without contraction (build with -ffp-contract=off).
 *
 * Extensions beyond the reference (north_star rows n1-n7; semantics inherited from the
 * sphere path): brute-force and BVH-accelerated watertight ray/triangle closest hit with
 * the lexicographic (t, primitive id) tie rule, parametric spp/bounces, fp32 progressive
 * accumulation.
 */
#ifndef MINOTE_ORACLE_H
#define MINOTE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PODs shared with the product's C ABI (layouts restated, not included) ---- */

/* src/gfx/camera.ixx:8-22 */
typedef struct {
    uint32_t viewport[2];
    float verticalFov;
    float nearPlane;
    float position[3];
    float yaw;
    float pitch;
    float lookSpeed;
    float moveSpeed;
} orc_camera;

/* column-major 4x4: m[col][row]  (src/stx/math.ixx:531-539) */
typedef struct { float m[4][4]; } orc_mat4;

/* src/gfx/modules/pathtracer.ixx:86-93  <-> src/gpu/primaryRay.comp:14-21 */
typedef struct {
    orc_mat4 view, projection, invView, invProjection, prevView;
    uint32_t frameCounter;
} orc_primary_constants;

/* src/gfx/modules/pathtracer.ixx:170-177 <-> src/gpu/secondaryRays.comp:25-32 */
typedef struct {
    orc_mat4 view, projection, invView, invProjection;
    float cameraPos[3];
    uint32_t frameCounter;
} orc_secondary_constants;

/* src/gpu/intersect.glsl:9-13 */
typedef struct { float center[3]; float radius; float albedo[3]; } orc_sphere;

/* src/gfx/modules/sky.ixx:28-56 (std140 mirror, 144 bytes) */
typedef struct {
    float bottomRadius, topRadius, rayleighDensityExpScale, _pad0;
    float rayleighScattering[3]; float mieDensityExpScale;
    float mieScattering[3]; float _pad1;
    float mieExtinction[3]; float _pad2;
    float mieAbsorption[3]; float miePhaseG;
    float absorptionDensity0LayerWidth, absorptionDensity0ConstantTerm,
          absorptionDensity0LinearTerm, absorptionDensity1ConstantTerm;
    float absorptionDensity1LinearTerm, _pad3, _pad4, _pad5;
    float absorptionExtinction[3]; float _pad6;
    float groundAlbedo[3]; float _pad7;
} orc_atmosphere_params;

enum { ORC_TRANS_W = 256, ORC_TRANS_H = 64, ORC_MULTI_W = 32, ORC_MULTI_H = 32,
       ORC_VIEW_W = 192, ORC_VIEW_H = 108 };

enum { ORC_TONEMAP_LINEAR = 0, ORC_TONEMAP_REINHARD = 1, ORC_TONEMAP_HABLE = 2,
       ORC_TONEMAP_ACES = 3, ORC_TONEMAP_UCHIMURA = 4, ORC_TONEMAP_AMD = 5 };

/* ---- storage formats ---- */
uint16_t orc_f32_to_f16(float f);
float    orc_f16_to_f32(uint16_t h);
uint32_t orc_pack_b10g11r11(const float rgb[3]);
void     orc_unpack_b10g11r11(uint32_t p, float rgb[3]);
uint8_t  orc_unorm8(float f);

/* ---- host matrices (a1) ---- */
void orc_camera_direction(const orc_camera* c, float out[3]);
void orc_camera_view(const orc_camera* c, orc_mat4* out);
void orc_camera_projection(const orc_camera* c, orc_mat4* out);
void orc_look(const float pos[3], const float dir[3], const float up[3], orc_mat4* out);
void orc_perspective(float vFov, float aspect, float zNear, orc_mat4* out);
void orc_inverse(const orc_mat4* m, orc_mat4* out);
void orc_mat_mul(const orc_mat4* a, const orc_mat4* b, orc_mat4* out);
void orc_primary_constants_fill(const orc_camera* cam, const orc_camera* prev, uint32_t frame,
                                orc_primary_constants* out);
void orc_secondary_constants_fill(const orc_camera* cam, uint32_t frame,
                                  orc_secondary_constants* out);
void orc_atmosphere_earth(orc_atmosphere_params* out);
/* Camera::rotate / shift / roam (src/gfx/camera.ixx:47-63) */
void orc_camera_rotate(orc_camera* c, float horz, float vert);
void orc_camera_shift(orc_camera* c, const float d[3]);
void orc_camera_roam(orc_camera* c, const float d[3]);

/* ---- RNG / sampling (a7-a9) ---- */
uint32_t orc_pcg(uint32_t* state);
float    orc_random_float(uint32_t* state);
void     orc_random_sphere_point(float rx, float ry, float out[3]);

/* ---- primitives (a3, n4) ---- */
float orc_ray_sphere(const float o[3], const float d[3], const orc_sphere* s);
/* returns 1 on hit; *t,*u,*v set (u,v barycentrics of v1,v2) */
int orc_ray_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
                     const float v2[3], float* t, float* u, float* v);
/* the same test for n rays against one triangle (o, d: n x 3 floats); hit[i] = 0/1, t[i] valid where hit */
void orc_ray_triangle_batch(uint32_t n, const float* o, const float* d, const float v0[3], const float v1[3],
                            const float v2[3], uint8_t* hit, float* t);
void orc_ray_gen(const orc_mat4* invView, const orc_mat4* invProjection, uint32_t x, uint32_t y,
                 uint32_t w, uint32_t h, float origin[3], float dir[3]);

/* ---- sky (a11, a13) ---- */
void orc_gen_transmittance(const orc_atmosphere_params* p, uint16_t* rgba16f /*256*64*4*/);
void orc_gen_multiscattering(const orc_atmosphere_params* p, const uint16_t* trans,
                             uint16_t* rgba16f /*32*32*4*/);
void orc_gen_sky_view(const orc_atmosphere_params* p, const uint16_t* trans, const uint16_t* multi,
                      const float probePos[3], const float sunDir[3], const float sunIlluminance[3],
                      uint32_t* b10g11r11 /*192*108*/);
void orc_sky_color(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView,
                   const float cameraPos[3], const float dir[3], float out[3]);

/* the same for n directions (dirs, out: n x 3 floats) */
void orc_sky_color_batch(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView,
                         const float cameraPos[3], uint32_t n, const float* dirs, float* out);

/* ---- reference sphere path, faithful mode (a2-a12) ---- */
void orc_primary_rays_spheres(uint32_t w, uint32_t h, const orc_primary_constants* c,
                              const orc_sphere* spheres, uint32_t nspheres,
                              uint32_t* visibility, uint16_t* depth, uint16_t* normal /*4/px*/,
                              uint16_t* motion /*2/px*/);
/* color16: RGBA16F (8 B/px) as the reference stores it; color32 (optional, may be NULL):
 * the same average before fp16 rounding.  rays_out (optional): secondary rays traced. */
void orc_secondary_rays_spheres(uint32_t w, uint32_t h, const orc_secondary_constants* c,
                                const orc_sphere* spheres, uint32_t nspheres,
                                const uint32_t* visibility, const uint16_t* depth,
                                const uint16_t* normal, const uint8_t* blueNoise, uint32_t bnW,
                                uint32_t bnH, const orc_atmosphere_params* atmo,
                                const uint16_t* trans, const uint32_t* skyView, uint32_t spp,
                                uint32_t bounces, uint16_t* color16, float* color32,
                                uint64_t* rays_out);

/* ---- bilateral denoiser (src/gpu/denoise/bilateral.comp, denoiser.ixx:36-97) ---- */
void orc_denoise_bilateral(uint32_t w, uint32_t h, const uint16_t* color16 /*RGBA16F*/,
                           const uint16_t* depth16 /*R16F*/, const uint16_t* normal16 /*RGBA16F*/,
                           float sigma, float kSigma, float threshold, float nearPlane,
                           uint32_t frameCounter, uint8_t* rgba8);

/* ---- compressed 8-wide BVH node: quantiser + slab test (north_star row n3; no reference counterpart) ----
 * CPU restatement of the product's node arithmetic, operation for operation, so that its central claim can be
 * checked WITHOUT a GPU: the fp32 slab test on the quantised planes never culls a child box the exact ray touches.
 *   orc_wide_node_quantize : minotert_b200/csrc/bvh_build.cu grid_exponent + k_emit_nodes (grid origin two steps
 *                            below the node box, power-of-two step with 250 steps across the box, planes rounded
 *                            outward with >= 1/64 step of slack, empty slot = (255, 0))
 *   orc_wide_node_test     : minotert_b200/csrc/trace.cuh lane_begin + lane_node_step (1/d scaled by 1 -/+ 2^-21,
 *                            far planes decoded as the float 2^15 + q, near planes as 2^15 + q/2, one FMA per plane,
 *                            hit <=> no sign bit in (tmax - tmin) | (tlimit - tmin) | tmax)
 * Node words: w[0..2] grid origin (fp32 bits), w[3] = ex | ey << 8 | ez << 16 (biased exponents of the step),
 * w[4..9] = qlo x[0..3], x[4..7], y.., z..; w[10..15] = qhi likewise.  present: bit s = slot s holds a child. */
typedef struct { uint32_t w[16]; } orc_wide_node;
void orc_wide_node_quantize(const float lo[8][3], const float hi[8][3], uint32_t present, orc_wide_node* out);
/* bit s of the result: child s passes the slab test for the ray (o, d) with the current closest hit at t_best
 * (pass 3.0e38f for "none").  Rays are tested in batches: o, d: n x 3 floats, t_best: n floats, hits: n words. */
void orc_wide_node_test(const orc_wide_node* node, uint32_t n, const float* o, const float* d, const float* t_best,
                        uint32_t* hits);

/* ---- temporal reprojection ----
 * No reference counterpart: the reference WRITES the motion of every primary hit between prevView and view
 * (src/gpu/primaryRay.comp:73-75, RG16F, pathtracer.ixx:63-69) and nothing reads it (renderer.ixx:61).  This is the
 * consumer: an exponential moving average of the frame's radiance with the history fetched at the reprojected pixel.
 *   cur      = accum.rgb / accum.w                              (the frame's average, row n7)
 *   miss     : visibility == 0xFFFFFFFF -> out = cur, count 1   (the sky is noise-free and carries motion 0)
 *   prev px  : (x + 0.5 - m.x / 2, y + 0.5 + m.y / 2)           (m = motion texel; ndc.y is flipped, primaryRay.comp:46)
 *   history  : bilinear over the 4 nearest history texels, taps outside the image or whose primitive id differs
 *              from the pixel's are dropped and the weights renormalised; total weight <= 1/256 -> out = cur, count 1
 *   blend    : n = min(history count, maxHistory); out = hist + (cur - hist) / (n + 1); count = n + 1
 * accum: RGBA32F, vis: R32_UINT, motion16: RG16F; hist_rgba/hist_vis: previous outputs (have_history = 0: ignored);
 * out_rgba: RGBA32F (rgb, 1); out_count: R32F. */
void orc_temporal_accumulate(uint32_t w, uint32_t h, const float* accum, const uint32_t* vis, const uint16_t* motion16,
                             int have_history, const float* hist_rgba, const float* hist_count, const uint32_t* hist_vis,
                             float maxHistory, float* out_rgba, float* out_count);

/* ---- tonemap (a14) ---- */
/* src: RGBA32F if src_is_f16 == 0, RGBA16F if 1, RGBA8 unorm (denoiser output) if 2.  params: mode-specific push constants after exposure
 * (reinhard: hdrMax; uchimura: 6 floats; amd: 5 floats). */
void orc_tonemap(int mode, uint32_t w, uint32_t h, const void* src, int src_is_f16, float exposure,
                 const float* params, uint8_t* rgba8);
void orc_tonemap_pixel(int mode, const float rgb_in[3], float exposure, const float* params,
                       float rgb_out[3]);

/* ---- triangle scenes (n1-n7) ---- */
typedef struct orc_scene orc_scene;
orc_scene* orc_scene_create(const float* positions, uint32_t nverts, const uint32_t* indices,
                            uint32_t ntris, const float* albedo /*3/tri*/);
void orc_scene_destroy(orc_scene* s);
/* closest hit for one ray. use_bvh=0 -> brute force. returns prim id or 0xFFFFFFFF */
uint32_t orc_scene_closest_hit(const orc_scene* s, const float o[3], const float d[3], int use_bvh,
                               float* t, float* u, float* v);
/* G-buffer + fp32 hit (t; 0 on miss) for every pixel in rows [y0,y1) */
void orc_primary_rays_tris(const orc_scene* s, uint32_t w, uint32_t h,
                           const orc_primary_constants* c, int use_bvh, uint32_t y0, uint32_t y1,
                           uint32_t* visibility, uint16_t* depth, uint16_t* normal,
                           uint16_t* motion, float* hit_t);
/* native path trace: accum (RGBA32F, += sample sums, .w += spp) for rows [y0,y1).
 * rays_out: [0] primary rays, [1] secondary rays traced. */
void orc_render_tris(const orc_scene* s, uint32_t w, uint32_t h, const orc_primary_constants* pc,
                     const orc_secondary_constants* sc, const uint8_t* blueNoise, uint32_t bnW,
                     uint32_t bnH, const orc_atmosphere_params* atmo, const uint16_t* trans,
                     const uint32_t* skyView, uint32_t spp, uint32_t bounces, int use_bvh,
                     uint32_t y0, uint32_t y1, float* accum, uint32_t* visibility,
                     uint64_t* rays_out);
/* SURVEY 8f-4 extensions of the native path tracer (contract in minote_oracle.c at orc_render_tris_ext) */
#define ORC_EXT_NEE_SUN 1u
#define ORC_EXT_SKY_AT_HIT 2u
#define ORC_EXT_AERIAL 4u
void orc_render_tris_ext(const orc_scene* s, uint32_t w, uint32_t h, const orc_primary_constants* pc,
                         const orc_secondary_constants* sc, const uint8_t* blueNoise, uint32_t bnW,
                         uint32_t bnH, const orc_atmosphere_params* atmo, const uint16_t* trans,
                         const uint32_t* skyView, uint32_t spp, uint32_t bounces, int use_bvh,
                         uint32_t y0, uint32_t y1, float* accum, uint32_t* visibility,
                         uint64_t* rays_out, uint32_t ext, const uint16_t* aerial /* 32^3 RGBA16F or NULL */);
/* aerial-perspective volume (32 x 32 x 32 RGBA16F, x fastest) for a camera, and its lookup for a surface at distance t
 * (km) seen through image position (u, v) in [0,1]^2 (v down) */
void orc_gen_aerial_perspective(const orc_atmosphere_params* p, const uint16_t* trans, const uint16_t* multi,
                                const orc_mat4* invView, const orc_mat4* invProjection, const float cameraPos[3],
                                const float sunDirection[3], const float sunIlluminance[3], uint16_t* out);
void orc_aerial_perspective_lookup(const uint16_t* vol, float u, float v, float t, float out[4]);
/* one sun sample: direction inside the disc and the weight limb * Omega / pi; the sun centre's radiance seen from pos */
void orc_nee_sun_sample(float u0, float u1, float l[3], float* weight);
void orc_sun_centre_radiance(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView,
                             const float pos[3], float out[3]);
/* accum (RGBA32F sums, w = spp) -> RGBA32F average */
void orc_resolve(uint32_t npixels, const float* accum, float* color32);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
