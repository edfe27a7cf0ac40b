/*
 * minote_oracle.c -- CPU ORACLE (test infrastructure only; see minote_oracle.h).
 * PARITY UNPINNED by the reference (no reference tests / goldens / runnable build exist).
 *
 * Plain C11 restatement of the MinoteRT hot path.  Citations are relative to the reference
 * repository root.  Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include "minote_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NONE_ID 0xFFFFFFFFu

/* =============================== small vector helpers =============================== */

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float vlen(v3 a) { return sqrtf(vdot(a, a)); }
static inline v3 vnorm(v3 a) { return vdivs(a, vlen(a)); }
/* GLSL cross(); same component formula as src/stx/math.ixx:312-318 */
static inline v3 vcross(v3 a, v3 b) {
    return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 vexp(v3 a) { return V(expf(a.x), expf(a.y), expf(a.z)); }
static inline v3 vmaxs(float s, v3 a) { return V(fmaxf(s, a.x), fmaxf(s, a.y), fmaxf(s, a.z)); }
static inline v3 vsplat(float s) { return V(s, s, s); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline v3 v3p(const float* p) { return V(p[0], p[1], p[2]); }

typedef struct { float x, y, z, w; } v4;
/* GLSL mat4 * vec4 : columns scaled by components, summed left to right */
static inline v4 mat_vec(const orc_mat4* M, float x, float y, float z, float w) {
    v4 r;
    r.x = M->m[0][0] * x + M->m[1][0] * y + M->m[2][0] * z + M->m[3][0] * w;
    r.y = M->m[0][1] * x + M->m[1][1] * y + M->m[2][1] * z + M->m[3][1] * w;
    r.z = M->m[0][2] * x + M->m[1][2] * y + M->m[2][2] * z + M->m[3][2] * w;
    r.w = M->m[0][3] * x + M->m[1][3] * y + M->m[2][3] * z + M->m[3][3] * w;
    return r;
}

/* =============================== storage formats =============================== */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

uint16_t orc_f32_to_f16(float f) {
    uint32_t x = f2u(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7C00u | ((ax > 0x7F800000u) ? 0x200u : 0u));
    }
    if (ax >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* rounds to >= 65520 -> inf */
    if (ax < 0x38800000u) {                                    /* subnormal half or zero */
        if (ax < 0x33000000u) return (uint16_t)sign;           /* < 2^-25 -> 0 */
        uint32_t exp = ax >> 23;
        uint32_t mant = (ax & 0x7FFFFFu) | 0x800000u;
        uint32_t shift = 126u - exp; /* 14..24 */
        uint32_t half = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1u);
        if (rem > halfway || (rem == halfway && (half & 1u))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t r = ax - 0x38000000u; /* rebias 127 -> 15 */
    uint32_t half = r >> 13;
    uint32_t rem = r & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
    return (uint16_t)(sign | half);
}

float orc_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t mant = h & 0x3FFu;
    if (exp == 0) {
        float v = (float)mant * 5.9604644775390625e-8f; /* 2^-24 */
        return sign ? -v : v;
    }
    if (exp == 31) return u2f(sign | 0x7F800000u | (mant << 13));
    return u2f(sign | ((exp + 112u) << 23) | (mant << 13));
}

/* unsigned small float with 5-bit exponent (bias 15) and `mb` mantissa bits */
static uint32_t ufloat_pack(float f, int mb) {
    uint32_t maxv = (30u << mb) | ((1u << mb) - 1u);
    if (!(f > 0.0f)) return 0u; /* negatives, -0, NaN -> 0 */
    uint32_t x = f2u(f);
    if (x >= 0x7F800000u) return maxv;
    if (x < 0x38800000u) { /* below 2^-14: denormal */
        float q = rintf(f * (float)(1u << (14 + mb)));
        return (uint32_t)q; /* q == 1<<mb encodes exp=1,mant=0 correctly */
    }
    int drop = 23 - mb;
    uint32_t r = x - 0x38000000u;
    uint32_t v = r >> drop;
    uint32_t rem = r & ((1u << drop) - 1u);
    uint32_t halfway = 1u << (drop - 1);
    if (rem > halfway || (rem == halfway && (v & 1u))) v++;
    if (v > maxv) v = maxv;
    return v;
}

static float ufloat_unpack(uint32_t v, int mb) {
    uint32_t exp = v >> mb;
    uint32_t mant = v & ((1u << mb) - 1u);
    if (exp == 0) return (float)mant * (6.103515625e-05f / (float)(1u << mb));
    if (exp == 31) return mant ? NAN : INFINITY;
    return u2f(((exp + 112u) << 23) | (mant << (23 - mb)));
}

uint32_t orc_pack_b10g11r11(const float rgb[3]) {
    return ufloat_pack(rgb[0], 6) | (ufloat_pack(rgb[1], 6) << 11) | (ufloat_pack(rgb[2], 5) << 22);
}

void orc_unpack_b10g11r11(uint32_t p, float rgb[3]) {
    rgb[0] = ufloat_unpack(p & 0x7FFu, 6);
    rgb[1] = ufloat_unpack((p >> 11) & 0x7FFu, 6);
    rgb[2] = ufloat_unpack((p >> 22) & 0x3FFu, 5);
}

uint8_t orc_unorm8(float f) {
    if (!(f == f)) return 0;
    return (uint8_t)rintf(clampf(f, 0.0f, 1.0f) * 255.0f);
}

/* =============================== host matrices (a1) =============================== */

/* src/stx/math.ixx:27 : radians<T, Prec = float>(deg) = Prec(deg) * Tau_v<Prec> / Prec(360) -- the _deg literals
 * (:884-886) pass a double but leave Prec at float, so the arithmetic is fp32 (pinned by oracle/_ref: 22_deg and
 * 45.5_deg differ in the last bit from the double-then-narrow value; 60, 89, 90, 360 do not) */
static float deg_lit(double deg) { return (float)deg * (3.14159265358979323846f * 2.0f) / 360.0f; }

/* src/gfx/camera.ixx:26-32 */
void orc_camera_direction(const orc_camera* c, float out[3]) {
    out[0] = cosf(c->pitch) * cosf(c->yaw);
    out[1] = cosf(c->pitch) * sinf(c->yaw);
    out[2] = sinf(c->pitch);
}

/* src/stx/math.ixx:823-844 */
void orc_look(const float pos[3], const float dir[3], const float up[3], orc_mat4* out) {
    v3 p = v3p(pos), d = v3p(dir), u0 = v3p(up);
    v3 s = vnorm(vcross(u0, d));
    v3 u = vcross(d, s);
    memset(out, 0, sizeof *out);
    out->m[0][0] = -s.x; out->m[1][0] = -s.y; out->m[2][0] = -s.z;
    out->m[0][1] = u.x;  out->m[1][1] = u.y;  out->m[2][1] = u.z;
    out->m[0][2] = d.x;  out->m[1][2] = d.y;  out->m[2][2] = d.z;
    /* dot(): accumulation starts from 0 (src/stx/math.ixx:304-309) */
    out->m[3][0] = (0.0f + s.x * p.x + s.y * p.y + s.z * p.z);
    out->m[3][1] = -(0.0f + u.x * p.x + u.y * p.y + u.z * p.z);
    out->m[3][2] = -(0.0f + d.x * p.x + d.y * p.y + d.z * p.z);
    out->m[3][3] = 1.0f;
}

/* src/stx/math.ixx:849-860 : inverted infinite-Z */
void orc_perspective(float vFov, float aspect, float zNear, orc_mat4* out) {
    float h = 1.0f / tanf(0.5f * vFov);
    float w = h * aspect;
    memset(out, 0, sizeof *out);
    out->m[0][0] = w;
    out->m[1][1] = h;
    out->m[2][3] = 1.0f;
    out->m[3][2] = zNear;
}

/* src/gfx/camera.ixx:36-44 */
void orc_camera_view(const orc_camera* c, orc_mat4* out) {
    float d[3], up[3] = {0.0f, 0.0f, 1.0f};
    orc_camera_direction(c, d);
    orc_look(c->position, d, up, out);
}
void orc_camera_projection(const orc_camera* c, orc_mat4* out) {
    orc_perspective(c->verticalFov, (float)c->viewport[1] / (float)c->viewport[0], c->nearPlane, out);
}

/* src/stx/math.ixx:761-815 (4x4 cofactor inverse; operation order kept) */
void orc_inverse(const orc_mat4* M, orc_mat4* out) {
    const float (*m)[4] = M->m;
    float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    float a0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]};
    float a1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    float a2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]};
    float a3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    static const float sa[4] = {1, -1, 1, -1}, sb[4] = {-1, 1, -1, 1};
    float inv[4][4];
    for (int i = 0; i < 4; i++) {
        inv[0][i] = (a1[i] * f0[i] - a2[i] * f1[i] + a3[i] * f2[i]) * sa[i];
        inv[1][i] = (a0[i] * f0[i] - a2[i] * f3[i] + a3[i] * f4[i]) * sb[i];
        inv[2][i] = (a0[i] * f1[i] - a1[i] * f3[i] + a3[i] * f5[i]) * sa[i];
        inv[3][i] = (a0[i] * f2[i] - a1[i] * f4[i] + a2[i] * f5[i]) * sb[i];
    }
    float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0],
          d3 = m[0][3] * inv[3][0];
    float det = (d0 + d1) + (d2 + d3);
    float ood = 1.0f / det;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out->m[c][r] = inv[c][r] * ood;
}

/* src/stx/math.ixx:680-696 == GLSL mat4*mat4 */
void orc_mat_mul(const orc_mat4* a, const orc_mat4* b, orc_mat4* out) {
    orc_mat4 r;
    for (int c = 0; c < 4; c++) {
        v4 col = mat_vec(a, b->m[c][0], b->m[c][1], b->m[c][2], b->m[c][3]);
        r.m[c][0] = col.x; r.m[c][1] = col.y; r.m[c][2] = col.z; r.m[c][3] = col.w;
    }
    *out = r;
}

/* src/gfx/modules/pathtracer.ixx:94-104 */
void orc_primary_constants_fill(const orc_camera* cam, const orc_camera* prev, uint32_t frame,
                                orc_primary_constants* out) {
    orc_camera_view(cam, &out->view);
    orc_camera_projection(cam, &out->projection);
    orc_inverse(&out->view, &out->invView);
    orc_inverse(&out->projection, &out->invProjection);
    orc_camera_view(prev, &out->prevView);
    out->frameCounter = frame;
}

/* src/gfx/modules/pathtracer.ixx:178-188 */
void orc_secondary_constants_fill(const orc_camera* cam, uint32_t frame, orc_secondary_constants* out) {
    orc_camera_view(cam, &out->view);
    orc_camera_projection(cam, &out->projection);
    orc_inverse(&out->view, &out->invView);
    orc_inverse(&out->projection, &out->invProjection);
    out->cameraPos[0] = cam->position[0];
    out->cameraPos[1] = cam->position[1];
    out->cameraPos[2] = cam->position[2];
    out->frameCounter = frame;
}

/* src/gfx/camera.ixx:47-63 */
void orc_camera_rotate(orc_camera* c, float horz, float vert) {
    c->yaw -= horz * c->lookSpeed;
    if (c->yaw < deg_lit(0)) c->yaw += deg_lit(360);
    if (c->yaw >= deg_lit(360)) c->yaw -= deg_lit(360);
    c->pitch += vert * c->lookSpeed;
    c->pitch = fmaxf(-deg_lit(89), fminf(c->pitch, deg_lit(89)));
}
void orc_camera_shift(orc_camera* c, const float d[3]) {
    for (int i = 0; i < 3; i++) c->position[i] += d[i] * c->moveSpeed;
}
void orc_camera_roam(orc_camera* c, const float d[3]) {
    orc_mat4 view, inv;
    orc_camera_view(c, &view);
    orc_inverse(&view, &inv);
    /* mat*vec on the host = dot(row_i, v) accumulated from 0 (src/stx/math.ixx:699-706) */
    float r[3];
    for (int i = 0; i < 3; i++)
        r[i] = 0.0f + inv.m[0][i] * d[0] + inv.m[1][i] * d[1] + inv.m[2][i] * d[2] + inv.m[3][i] * 0.0f;
    orc_camera_shift(c, r);
}

/* src/gfx/modules/sky.ixx:59-83 */
void orc_atmosphere_earth(orc_atmosphere_params* p) {
    memset(p, 0, sizeof *p);
    p->bottomRadius = 6360.0f;
    p->topRadius = 6460.0f;
    p->rayleighDensityExpScale = -1.0f / 8.0f;
    p->rayleighScattering[0] = 0.005802f; p->rayleighScattering[1] = 0.013558f; p->rayleighScattering[2] = 0.033100f;
    p->mieDensityExpScale = -1.0f / 1.2f;
    for (int i = 0; i < 3; i++) {
        p->mieScattering[i] = 0.003996f;
        p->mieExtinction[i] = 0.004440f;
        p->mieAbsorption[i] = fmaxf(0.004440f - 0.003996f, 0.0f);
    }
    p->miePhaseG = 0.8f;
    p->absorptionDensity0LayerWidth = 25.0f;
    p->absorptionDensity0ConstantTerm = -2.0f / 3.0f;
    p->absorptionDensity0LinearTerm = 1.0f / 15.0f;
    p->absorptionDensity1ConstantTerm = 8.0f / 3.0f;
    p->absorptionDensity1LinearTerm = -1.0f / 15.0f;
    p->absorptionExtinction[0] = 0.000650f; p->absorptionExtinction[1] = 0.001881f; p->absorptionExtinction[2] = 0.000085f;
}

/* =============================== RNG / sampling (a7-a9) =============================== */

/* src/gpu/random.glsl:22-27 */
uint32_t orc_pcg(uint32_t* v) {
    uint32_t state = *v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    *v = (word >> 22u) ^ word;
    return *v;
}

/* src/gpu/random.glsl:29-31 */
float orc_random_float(uint32_t* state) { return (float)(orc_pcg(state) & 0xFFFFFFu) / 16777216.0f; }

static const float kPi = 3.14159265359f; /* src/gpu/util.glsl:4 */

/* src/gpu/random.glsl:10-19 */
static v3 random_sphere_point(float rx, float ry) {
    float ang1 = (rx + 1.0f) * kPi;
    float u = ry;
    float u2 = u * u;
    float s = sqrtf(1.0f - u2);
    return V(s * cosf(ang1), s * sinf(ang1), u);
}
void orc_random_sphere_point(float rx, float ry, float out[3]) {
    v3 p = random_sphere_point(rx, ry);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

/* src/gpu/secondaryRays.comp:60-62 */
static inline float rotated_random(uint32_t* state, float rotation) {
    float x = orc_random_float(state) + rotation;
    return x - floorf(x);
}

/* =============================== primitives (a3, n4) =============================== */

/* src/gpu/intersect.glsl:26-37 */
static float ray_sphere(v3 o, v3 d, const orc_sphere* s) {
    v3 oc = vsub(o, v3p(s->center));
    float a = vdot(d, d);
    float half_b = vdot(oc, d);
    float c = vdot(oc, oc) - s->radius * s->radius;
    float disc = half_b * half_b - a * c;
    if (disc < 0.0f) return -1.0f;
    return (-half_b - sqrtf(disc)) / a;
}
float orc_ray_sphere(const float o[3], const float d[3], const orc_sphere* s) {
    return ray_sphere(v3p(o), v3p(d), s);
}

/* Watertight ray/triangle test (Woop, Benthin, Wald, JCGT 2013), fp32, no culling, with the
 * paper's double-precision fallback when an edge function is exactly zero.  North_star row n4;
 * no reference counterpart.  Accepts t >= 0 as src/gpu/intersect.glsl / primaryRay.comp:28 do. */
static inline float comp(v3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

typedef struct { int kx, ky, kz; float Sx, Sy, Sz; } ray_shear;

static ray_shear make_shear(v3 d) {
    ray_shear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = (r.kz + 1) % 3;
    r.ky = (r.kx + 1) % 3;
    if (comp(d, r.kz) < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    float dz = comp(d, r.kz);
    r.Sx = comp(d, r.kx) / dz;
    r.Sy = comp(d, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

static int ray_triangle(v3 o, const ray_shear* rs, v3 p0, v3 p1, v3 p2, float* t, float* u, float* v) {
    v3 A = vsub(p0, o), B = vsub(p1, o), C = vsub(p2, o);
    float Akz = comp(A, rs->kz), Bkz = comp(B, rs->kz), Ckz = comp(C, rs->kz);
    float Ax = comp(A, rs->kx) - rs->Sx * Akz, Ay = comp(A, rs->ky) - rs->Sy * Akz;
    float Bx = comp(B, rs->kx) - rs->Sx * Bkz, By = comp(B, rs->ky) - rs->Sy * Bkz;
    float Cx = comp(C, rs->kx) - rs->Sx * Ckz, Cy = comp(C, rs->ky) - rs->Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float Vv = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || Vv == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        Vv = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || Vv < 0.0f || W < 0.0f) && (U > 0.0f || Vv > 0.0f || W > 0.0f)) return 0;
    float det = U + Vv + W;
    if (det == 0.0f) return 0;
    float Az = rs->Sz * Akz, Bz = rs->Sz * Bkz, Cz = rs->Sz * Ckz;
    float T = U * Az + Vv * Bz + W * Cz;
    float tt = T / det;
    if (!(tt >= 0.0f)) return 0;
    *t = tt;
    *u = Vv / det;
    *v = W / det;
    return 1;
}

int orc_ray_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
                     const float v2[3], float* t, float* u, float* v) {
    ray_shear rs = make_shear(v3p(d));
    return ray_triangle(v3p(o), &rs, v3p(v0), v3p(v1), v3p(v2), t, u, v);
}

/* src/gpu/primaryRay.comp:40-56 */
static void ray_gen(const orc_mat4* invView, const orc_mat4* invProj, uint32_t x, uint32_t y, uint32_t w,
                    uint32_t h, v3* origin, v3* dir) {
    float pitchx = 1.0f / (float)w, pitchy = 1.0f / (float)h;
    float u = ((float)x + 0.5f) * pitchx;
    float v = ((float)y + 0.5f) * pitchy;
    v = 1.0f - v;
    v4 o = mat_vec(invView, 0.0f, 0.0f, 0.0f, 1.0f);
    v4 vd = mat_vec(invProj, u * 2.0f - 1.0f, v * 2.0f - 1.0f, 1.0f, 1.0f);
    v4 wd = mat_vec(invView, vd.x, vd.y, vd.z, 0.0f);
    *origin = V(o.x, o.y, o.z);
    *dir = vnorm(V(wd.x, wd.y, wd.z));
}
void orc_ray_gen(const orc_mat4* invView, const orc_mat4* invProjection, uint32_t x, uint32_t y,
                 uint32_t w, uint32_t h, float origin[3], float dir[3]) {
    v3 o, d;
    ray_gen(invView, invProjection, x, y, w, h, &o, &d);
    origin[0] = o.x; origin[1] = o.y; origin[2] = o.z;
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

/* =============================== textures =============================== */

typedef struct { int w, h; float* px; /* rgb triples */ } tex3;

static tex3 tex_from_rgba16f(const uint16_t* src, int w, int h) {
    tex3 t = {w, h, (float*)malloc(sizeof(float) * 3 * (size_t)w * h)};
    for (int i = 0; i < w * h; i++)
        for (int c = 0; c < 3; c++) t.px[3 * i + c] = orc_f16_to_f32(src[4 * i + c]);
    return t;
}
static tex3 tex_from_b10g11r11(const uint32_t* src, int w, int h) {
    tex3 t = {w, h, (float*)malloc(sizeof(float) * 3 * (size_t)w * h)};
    for (int i = 0; i < w * h; i++) orc_unpack_b10g11r11(src[i], &t.px[3 * i]);
    return t;
}
static void tex_free(tex3* t) { free(t->px); t->px = NULL; }

/* bilinear, Vulkan unnormalized-coordinate rule; repeat=0 -> clamp to edge
 * (samplers: src/gfx/samplers.ixx:14-26) */
static v3 tex_bilinear(const tex3* t, float u, float v, int repeat) {
    float x = u * (float)t->w - 0.5f, y = v * (float)t->h - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = x - fx0, fy = y - fy0;
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (repeat) {
        x0 %= t->w; if (x0 < 0) x0 += t->w;
        x1 %= t->w; if (x1 < 0) x1 += t->w;
        y0 %= t->h; if (y0 < 0) y0 += t->h;
        y1 %= t->h; if (y1 < 0) y1 += t->h;
    } else {
        x0 = x0 < 0 ? 0 : (x0 > t->w - 1 ? t->w - 1 : x0);
        x1 = x1 < 0 ? 0 : (x1 > t->w - 1 ? t->w - 1 : x1);
        y0 = y0 < 0 ? 0 : (y0 > t->h - 1 ? t->h - 1 : y0);
        y1 = y1 < 0 ? 0 : (y1 > t->h - 1 ? t->h - 1 : y1);
    }
    const float* p00 = &t->px[3 * (y0 * t->w + x0)];
    const float* p10 = &t->px[3 * (y0 * t->w + x1)];
    const float* p01 = &t->px[3 * (y1 * t->w + x0)];
    const float* p11 = &t->px[3 * (y1 * t->w + x1)];
    float gx = 1.0f - fx, gy = 1.0f - fy;
    v3 top = V(p00[0] * gx + p10[0] * fx, p00[1] * gx + p10[1] * fx, p00[2] * gx + p10[2] * fx);
    v3 bot = V(p01[0] * gx + p11[0] * fx, p01[1] * gx + p11[1] * fx, p01[2] * gx + p11[2] * fx);
    return V(top.x * gy + bot.x * fy, top.y * gy + bot.y * fy, top.z * gy + bot.z * fy);
}

/* =============================== sky (a11, a13) =============================== */

static const float kPI = 3.1415926535897932384626433832795f; /* src/gpu/constants.glsl:4 */
#define PLANET_RADIUS_OFFSET 0.01f /* src/gpu/sky/sky.glsl:13 */
#define RAYMARCH_MIN_SPP 4.0f
#define RAYMARCH_MAX_SPP 14.0f

/* src/gpu/sky/skyAccess.glsl:11-17 */
static float fromUnitToSubUvs(float u, float res) { return (u + 0.5f / res) * (res / (res + 1.0f)); }
static float fromSubUvsToUnit(float u, float res) { return (u - 0.5f / res) * (res / (res - 1.0f)); }

/* src/gpu/sky/skyAccess.glsl:19-34 */
static void uvToLutTransmittanceParams(float* viewHeight, float* viewZenithCosAngle, float u, float v,
                                       float bottom, float top) {
    float x_mu = u, x_r = v;
    float H = sqrtf(top * top - bottom * bottom);
    float rho = H * x_r;
    *viewHeight = sqrtf(rho * rho + bottom * bottom);
    float d_min = top - *viewHeight;
    float d_max = rho + H;
    float d = d_min + x_mu * (d_max - d_min);
    float c = d == 0.0f ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * *viewHeight * d);
    *viewZenithCosAngle = clampf(c, -1.0f, 1.0f);
}

/* src/gpu/sky/skyAccess.glsl:36-52 */
static void lutTransmittanceParamsToUv(float viewHeight, float viewZenithCosAngle, float* u, float* v,
                                       float bottom, float top) {
    float H = sqrtf(fmaxf(0.0f, top * top - bottom * bottom));
    float rho = sqrtf(fmaxf(0.0f, viewHeight * viewHeight - bottom * bottom));
    float disc = viewHeight * viewHeight * (viewZenithCosAngle * viewZenithCosAngle - 1.0f) + top * top;
    float d = fmaxf(0.0f, (-viewHeight * viewZenithCosAngle + sqrtf(disc)));
    float d_min = top - viewHeight;
    float d_max = rho + H;
    *u = (d - d_min) / (d_max - d_min);
    *v = rho / H;
}

/* src/gpu/sky/skyAccess.glsl:54-85 */
static void uvToSkyViewLutParams(float* viewZenithCosAngle, float* lightViewCosAngle, float sizeW,
                                 float sizeH, float viewHeight, float u, float v, float bottom) {
    u = fromSubUvsToUnit(u, sizeW);
    v = fromSubUvsToUnit(v, sizeH);
    float vHorizon = sqrtf(viewHeight * viewHeight - bottom * bottom);
    float cosBeta = vHorizon / viewHeight;
    float beta = acosf(cosBeta);
    float zenithHorizonAngle = kPI - beta;
    if (v < 0.5f) {
        float coord = 2.0f * v;
        coord = 1.0f - coord;
        coord *= coord;
        coord = 1.0f - coord;
        *viewZenithCosAngle = cosf(zenithHorizonAngle * coord);
    } else {
        float coord = v * 2.0f - 1.0f;
        coord *= coord;
        *viewZenithCosAngle = cosf(zenithHorizonAngle + beta * coord);
    }
    float coord = u;
    coord *= coord;
    *lightViewCosAngle = -(coord * 2.0f - 1.0f);
}

/* src/gpu/sky/skyAccess.glsl:87-117 */
static void skyViewLutParamsToUv(int intersectGround, float viewZenithCosAngle, float lightViewCosAngle,
                                 float sizeW, float sizeH, float viewHeight, float* u, float* v,
                                 float bottom) {
    float vHorizon = sqrtf(viewHeight * viewHeight - bottom * bottom);
    float cosBeta = vHorizon / viewHeight;
    float beta = acosf(cosBeta);
    float zenithHorizonAngle = kPI - beta;
    if (!intersectGround) {
        float coord = acosf(viewZenithCosAngle) / zenithHorizonAngle;
        coord = 1.0f - coord;
        coord = sqrtf(coord);
        coord = 1.0f - coord;
        *v = coord * 0.5f;
    } else {
        float coord = (acosf(viewZenithCosAngle) - zenithHorizonAngle) / beta;
        coord = sqrtf(coord);
        *v = coord * 0.5f + 0.5f;
    }
    float coord = -lightViewCosAngle * 0.5f + 0.5f;
    coord = sqrtf(coord);
    *u = coord;
    *u = fromUnitToSubUvs(*u, sizeW);
    *v = fromUnitToSubUvs(*v, sizeH);
}

/* src/gpu/sky/sky.glsl:58-77 */
static float raySphereIntersectNearest(v3 r0, v3 rd, v3 s0, float sR) {
    float a = vdot(rd, rd);
    v3 s0_r0 = vsub(r0, s0);
    float b = 2.0f * vdot(rd, s0_r0);
    float c = vdot(s0_r0, s0_r0) - (sR * sR);
    float delta = b * b - 4.0f * a * c;
    if (delta < 0.0f || a == 0.0f) return -1.0f;
    float sol0 = (-b - sqrtf(delta)) / (2.0f * a);
    float sol1 = (-b + sqrtf(delta)) / (2.0f * a);
    if (sol0 < 0.0f && sol1 < 0.0f) return -1.0f;
    if (sol0 < 0.0f) return fmaxf(0.0f, sol1);
    else if (sol1 < 0.0f) return fmaxf(0.0f, sol0);
    return fmaxf(0.0f, fminf(sol0, sol1));
}

/* src/gpu/sky/sky.glsl:79-94 */
static int moveToTopAtmosphere(v3* worldPos, v3 worldDir, float top) {
    float viewHeight = vlen(*worldPos);
    if (viewHeight > top) {
        float tTop = raySphereIntersectNearest(*worldPos, worldDir, V(0, 0, 0), top);
        if (tTop >= 0.0f) {
            v3 upVector = vdivs(*worldPos, viewHeight);
            v3 upOffset = vscale(upVector, -PLANET_RADIUS_OFFSET);
            *worldPos = vadd(vadd(*worldPos, vscale(worldDir, tTop)), upOffset);
        } else {
            return 0;
        }
    }
    return 1;
}

typedef struct { v3 scattering, extinction, scatteringMie, scatteringRay; } medium_t;

/* src/gpu/sky/sky.glsl:96-125 */
static medium_t sampleMediumRGB(const orc_atmosphere_params* A, v3 worldPos) {
    float viewHeight = vlen(worldPos) - A->bottomRadius;
    float densityMie = expf(A->mieDensityExpScale * viewHeight);
    float densityRay = expf(A->rayleighDensityExpScale * viewHeight);
    float densityOzo = clampf(viewHeight < A->absorptionDensity0LayerWidth
                                  ? A->absorptionDensity0LinearTerm * viewHeight + A->absorptionDensity0ConstantTerm
                                  : A->absorptionDensity1LinearTerm * viewHeight + A->absorptionDensity1ConstantTerm,
                              0.0f, 1.0f);
    medium_t s;
    v3 scatteringMie = vscale(v3p(A->mieScattering), densityMie);
    v3 absorptionMie = vscale(v3p(A->mieAbsorption), densityMie);
    v3 extinctionMie = vscale(v3p(A->mieExtinction), densityMie);
    v3 scatteringRay = vscale(v3p(A->rayleighScattering), densityRay);
    v3 absorptionRay = V(0, 0, 0);
    v3 extinctionRay = vadd(scatteringRay, absorptionRay);
    v3 scatteringOzo = V(0, 0, 0);
    v3 absorptionOzo = vscale(v3p(A->absorptionExtinction), densityOzo);
    v3 extinctionOzo = vadd(scatteringOzo, absorptionOzo);
    (void)absorptionMie;
    s.scatteringMie = scatteringMie;
    s.scatteringRay = scatteringRay;
    s.scattering = vadd(vadd(scatteringMie, scatteringRay), scatteringOzo);
    s.extinction = vadd(vadd(extinctionMie, extinctionRay), extinctionOzo);
    return s;
}

/* src/gpu/sky/sky.glsl:42-50 */
static float cornetteShanksMiePhaseFunction(float g, float cosTheta) {
    float k = 3.0f / (8.0f * kPI) * (1.0f - g * g) / (2.0f + g * g);
    return k * (1.0f + cosTheta * cosTheta) / powf(1.0f + g * g - 2.0f * g * -cosTheta, 1.5f);
}
static float rayleighPhase(float cosTheta) {
    float factor = 3.0f / (16.0f * kPI);
    return factor * (1.0f + cosTheta * cosTheta);
}

/* src/gpu/sky/sky.glsl:161-172 */
static v3 getMultipleScattering(const orc_atmosphere_params* A, const tex3* multi, v3 worldPos,
                                float viewZenithCosAngle) {
    float u = clampf(viewZenithCosAngle * 0.5f + 0.5f, 0.0f, 1.0f);
    float v = clampf((vlen(worldPos) - A->bottomRadius) / (A->topRadius - A->bottomRadius), 0.0f, 1.0f);
    u = fromUnitToSubUvs(u, (float)multi->w);
    v = fromUnitToSubUvs(v, (float)multi->h);
    return tex_bilinear(multi, u, v, 0);
}

typedef struct { v3 L, opticalDepth, transmittance, multiScatAs1; } scatter_result;

/* src/gpu/sky/sky.glsl:176-343.  trans / multi may be NULL (the S_TRANSMITTANCE /
 * S_MULTISCATTERING macros being undefined in the including shader). */
static scatter_result integrateScatteredLuminance(const orc_atmosphere_params* A, const tex3* trans,
                                                  const tex3* multi, v3 worldPos, v3 worldDir, v3 sunDir,
                                                  int ground, float sampleCountIni, int variableSampleCount,
                                                  int mieRayPhase, float tMaxMax, v3 sunIlluminance) {
    scatter_result result;
    memset(&result, 0, sizeof result);
    v3 earthO = V(0, 0, 0);
    float tBottom = raySphereIntersectNearest(worldPos, worldDir, earthO, A->bottomRadius);
    float tTop = raySphereIntersectNearest(worldPos, worldDir, earthO, A->topRadius);
    float tMax = 0.0f;
    if (tBottom < 0.0f) {
        if (tTop < 0.0f) {
            return result;
        } else {
            tMax = tTop;
        }
    } else if (tTop > 0.0f) {
        tMax = fminf(tTop, tBottom);
    }
    tMax = fminf(tMax, tMaxMax);

    float sampleCount = sampleCountIni;
    float sampleCountFloor = sampleCountIni;
    float tMaxFloor = tMax;
    if (variableSampleCount) {
        float a = clampf(tMax * 0.01f, 0.0f, 1.0f);
        sampleCount = RAYMARCH_MIN_SPP * (1.0f - a) + RAYMARCH_MAX_SPP * a; /* mix() */
        sampleCountFloor = floorf(sampleCount);
        tMaxFloor = tMax * sampleCountFloor / sampleCount;
    }
    float dt = tMax / sampleCount;

    float uniformPhase = 1.0f / (4.0f * kPI);
    float cosTheta = vdot(sunDir, worldDir);
    float miePhaseValue = cornetteShanksMiePhaseFunction(A->miePhaseG, -cosTheta);
    float rayleighPhaseValue = rayleighPhase(cosTheta);

    v3 globalL = sunIlluminance;
    v3 L = V(0, 0, 0), throughput = V(1, 1, 1), opticalDepth = V(0, 0, 0);
    float t = 0.0f;
    float sampleSegmentT = 0.3f;
    for (float s = 0.0f; s < sampleCount; s += 1.0f) {
        if (variableSampleCount) {
            float t0 = s / sampleCountFloor;
            float t1 = (s + 1.0f) / sampleCountFloor;
            t0 = t0 * t0;
            t1 = t1 * t1;
            t0 = tMaxFloor * t0;
            if (t1 > 1.0f) t1 = tMax;
            else t1 = tMaxFloor * t1;
            t = t0 + (t1 - t0) * sampleSegmentT;
            dt = t1 - t0;
        } else {
            float newT = tMax * (s + sampleSegmentT) / sampleCount;
            dt = newT - t;
            t = newT;
        }
        v3 P = vadd(worldPos, vscale(worldDir, t));

        medium_t medium = sampleMediumRGB(A, P);
        v3 sampleOpticalDepth = vscale(medium.extinction, dt);
        v3 sampleTransmittance = vexp(vneg(sampleOpticalDepth));
        opticalDepth = vadd(opticalDepth, sampleOpticalDepth);

        float pHeight = vlen(P);
        v3 upVector = vdivs(P, pHeight);
        float sunZenithCosAngle = vdot(sunDir, upVector);
        float u, v;
        lutTransmittanceParamsToUv(pHeight, sunZenithCosAngle, &u, &v, A->bottomRadius, A->topRadius);
        v3 transmittanceToSun = trans ? tex_bilinear(trans, u, v, 0) : V(0, 0, 0);

        v3 phaseTimesScattering;
        if (mieRayPhase)
            phaseTimesScattering = vadd(vscale(medium.scatteringMie, miePhaseValue),
                                        vscale(medium.scatteringRay, rayleighPhaseValue));
        else
            phaseTimesScattering = vscale(medium.scattering, uniformPhase);

        float tEarth = raySphereIntersectNearest(P, sunDir, vadd(earthO, vscale(upVector, PLANET_RADIUS_OFFSET)),
                                                 A->bottomRadius);
        float earthShadow = tEarth >= 0.0f ? 0.0f : 1.0f;

        v3 multiScatteredLuminance = multi ? getMultipleScattering(A, multi, P, sunZenithCosAngle) : V(0, 0, 0);

        /* S = globalL * (earthShadow * transmittanceToSun * phaseTimesScattering + msL * scattering) */
        v3 S = vmul(globalL, vadd(vmul(vscale(transmittanceToSun, earthShadow), phaseTimesScattering),
                                  vmul(multiScatteredLuminance, medium.scattering)));

        v3 MS = vscale(medium.scattering, 1.0f);
        v3 MSint = vdiv(vsub(MS, vmul(MS, sampleTransmittance)), medium.extinction);
        result.multiScatAs1 = vadd(result.multiScatAs1, vmul(throughput, MSint));

        v3 Sint = vdiv(vsub(S, vmul(S, sampleTransmittance)), medium.extinction);
        L = vadd(L, vmul(throughput, Sint));
        throughput = vmul(throughput, sampleTransmittance);
    }

    if (ground && tMax == tBottom && tBottom > 0.0f) {
        v3 P = vadd(worldPos, vscale(worldDir, tBottom));
        float pHeight = vlen(P);
        v3 upVector = vdivs(P, pHeight);
        float sunZenithCosAngle = vdot(sunDir, upVector);
        float u, v;
        lutTransmittanceParamsToUv(pHeight, sunZenithCosAngle, &u, &v, A->bottomRadius, A->topRadius);
        v3 transmittanceToSun = trans ? tex_bilinear(trans, u, v, 0) : V(0, 0, 0);
        float NdotL = clampf(vdot(vnorm(upVector), vnorm(sunDir)), 0.0f, 1.0f);
        /* L += globalL * transmittanceToSun * throughput * NdotL * groundAlbedo / PI */
        v3 term = vmul(vmul(globalL, transmittanceToSun), throughput);
        term = vscale(term, NdotL);
        term = vmul(term, v3p(A->groundAlbedo));
        term = vdivs(term, kPI);
        L = vadd(L, term);
    }

    result.L = L;
    result.opticalDepth = opticalDepth;
    result.transmittance = throughput;
    return result;
}

/* src/gpu/sky/genTransmittance.comp:21-44 ; size/format src/gfx/modules/sky.ixx:22-23 */
void orc_gen_transmittance(const orc_atmosphere_params* A, uint16_t* out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < ORC_TRANS_H; y++)
        for (int x = 0; x < ORC_TRANS_W; x++) {
            float u = ((float)x + 0.5f) / (float)ORC_TRANS_W;
            float v = ((float)y + 0.5f) / (float)ORC_TRANS_H;
            float viewHeight, viewZenithCosAngle;
            uvToLutTransmittanceParams(&viewHeight, &viewZenithCosAngle, u, v, A->bottomRadius, A->topRadius);
            v3 worldPos = V(0.0f, 0.0f, viewHeight);
            v3 worldDir = V(0.0f, sqrtf(1.0f - viewZenithCosAngle * viewZenithCosAngle), viewZenithCosAngle);
            scatter_result r = integrateScatteredLuminance(A, NULL, NULL, worldPos, worldDir, V(1, 1, 1), 0,
                                                           40.0f, 0, 0, 9000000.0f, V(1, 1, 1));
            v3 res = vexp(vneg(r.opticalDepth));
            uint16_t* o = &out[4 * (y * ORC_TRANS_W + x)];
            o[0] = orc_f32_to_f16(res.x); o[1] = orc_f32_to_f16(res.y); o[2] = orc_f32_to_f16(res.z);
            o[3] = orc_f32_to_f16(1.0f);
        }
}

/* src/gpu/sky/genMultiScattering.comp:26-146 */
void orc_gen_multiscattering(const orc_atmosphere_params* A, const uint16_t* trans16, uint16_t* out) {
    tex3 trans = tex_from_rgba16f(trans16, ORC_TRANS_W, ORC_TRANS_H);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < ORC_MULTI_H; y++)
        for (int x = 0; x < ORC_MULTI_W; x++) {
            float u = ((float)x + 0.5f) / (float)ORC_MULTI_W;
            float v = ((float)y + 0.5f) / (float)ORC_MULTI_H;
            u = fromSubUvsToUnit(u, (float)ORC_MULTI_W);
            v = fromSubUvsToUnit(v, (float)ORC_MULTI_H);
            float cosSunZenithAngle = u * 2.0f - 1.0f;
            v3 sunDir = V(0.0f, sqrtf(clampf(1.0f - cosSunZenithAngle * cosSunZenithAngle, 0.0f, 1.0f)),
                          cosSunZenithAngle);
            float viewHeight = A->bottomRadius + clampf(v + PLANET_RADIUS_OFFSET, 0.0f, 1.0f) *
                                                     (A->topRadius - A->bottomRadius - PLANET_RADIUS_OFFSET);
            v3 worldPos = V(0.0f, 0.0f, viewHeight);
            float sphereSolidAngle = 4.0f * kPI;
            float isotropicPhase = 1.0f / sphereSolidAngle;
            float sqrtSample = 8.0f;
            v3 shMS[64], shL[64];
            for (int z = 0; z < 64; z++) {
                float i = 0.5f + (float)(z / 8);
                float j = 0.5f + (float)(z % 8);
                float randA = i / sqrtSample, randB = j / sqrtSample;
                float theta = 2.0f * kPI * randA;
                float phi = kPI * randB;
                float cosPhi = cosf(phi), sinPhi = sinf(phi), cosTheta = cosf(theta), sinTheta = sinf(theta);
                v3 worldDir = V(cosTheta * sinPhi, sinTheta * sinPhi, cosPhi);
                scatter_result r = integrateScatteredLuminance(A, &trans, NULL, worldPos, worldDir, sunDir, 1,
                                                               20.0f, 0, 0, 9000000.0f, V(1, 1, 1));
                shMS[z] = vdivs(vscale(r.multiScatAs1, sphereSolidAngle), sqrtSample * sqrtSample);
                shL[z] = vdivs(vscale(r.L, sphereSolidAngle), sqrtSample * sqrtSample);
            }
            /* shared-memory tree reduction, same pairing (:79-116) */
            for (int stride = 32; stride >= 1; stride >>= 1)
                for (int z = 0; z < stride; z++) {
                    shMS[z] = vadd(shMS[z], shMS[z + stride]);
                    shL[z] = vadd(shL[z], shL[z + stride]);
                }
            v3 multiScatAs1 = vscale(shMS[0], isotropicPhase);
            v3 inScatteredLuminance = vscale(shL[0], isotropicPhase);
            /* MULTI_SCATTERING_POWER_SERIE is undefined => evaluates as 0 => 5-term series (:133-136) */
            v3 sq = vmul(multiScatAs1, multiScatAs1);
            v3 series = vadd(vadd(vadd(vadd(vsplat(1.0f), multiScatAs1), sq), vmul(multiScatAs1, sq)), vmul(sq, sq));
            v3 Lr = vmul(inScatteredLuminance, series);
            uint16_t* o = &out[4 * (y * ORC_MULTI_W + x)];
            o[0] = orc_f32_to_f16(Lr.x); o[1] = orc_f32_to_f16(Lr.y); o[2] = orc_f32_to_f16(Lr.z);
            o[3] = orc_f32_to_f16(1.0f);
        }
    tex_free(&trans);
}

/* src/gpu/sky/genView.comp:33-77 ; size/format src/gfx/modules/sky.ixx:187-188 */
void orc_gen_sky_view(const orc_atmosphere_params* A, const uint16_t* trans16, const uint16_t* multi16,
                      const float probePos[3], const float sunDirection[3], const float sunIlluminance[3],
                      uint32_t* out) {
    tex3 trans = tex_from_rgba16f(trans16, ORC_TRANS_W, ORC_TRANS_H);
    tex3 multi = tex_from_rgba16f(multi16, ORC_MULTI_W, ORC_MULTI_H);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < ORC_VIEW_H; y++)
        for (int x = 0; x < ORC_VIEW_W; x++) {
            v3 worldPos = vadd(v3p(probePos), V(0, 0, A->bottomRadius));
            float u = (float)x / (float)ORC_VIEW_W;
            float v = (float)y / (float)ORC_VIEW_H;
            float viewHeight = vlen(worldPos);
            float viewZenithCosAngle, lightViewCosAngle;
            uvToSkyViewLutParams(&viewZenithCosAngle, &lightViewCosAngle, (float)ORC_VIEW_W, (float)ORC_VIEW_H,
                                 viewHeight, u, v, A->bottomRadius);
            v3 upVector = vdivs(worldPos, viewHeight);
            float sunZenithCosAngle = vdot(upVector, v3p(sunDirection));
            v3 sunDir = vnorm(V(sqrtf(1.0f - sunZenithCosAngle * sunZenithCosAngle), 0.0f, sunZenithCosAngle));
            worldPos = V(0.0f, 0.0f, viewHeight);
            float viewZenithSinAngle = sqrtf(1.0f - viewZenithCosAngle * viewZenithCosAngle);
            v3 worldDir = V(viewZenithSinAngle * lightViewCosAngle,
                            viewZenithSinAngle * sqrtf(1.0f - lightViewCosAngle * lightViewCosAngle),
                            viewZenithCosAngle);
            float rgb[3] = {0, 0, 0};
            if (moveToTopAtmosphere(&worldPos, worldDir, A->topRadius)) {
                scatter_result ss = integrateScatteredLuminance(A, &trans, &multi, worldPos, worldDir, sunDir, 0,
                                                                30.0f, 1, 1, 9000000.0f, v3p(sunIlluminance));
                rgb[0] = ss.L.x; rgb[1] = ss.L.y; rgb[2] = ss.L.z;
            }
            out[y * ORC_VIEW_W + x] = orc_pack_b10g11r11(rgb);
        }
    tex_free(&trans);
    tex_free(&multi);
}

typedef struct {
    const orc_atmosphere_params* A;
    tex3 trans, view;
    v3 cameraPos;
} sky_ctx;

/* src/gpu/sky/sky.glsl:129-155 */
static v3 getSunLuminance(const sky_ctx* S, v3 worldPos, v3 worldDir, v3 sunDirection, v3 sunIlluminance) {
    const orc_atmosphere_params* A = S->A;
    float SunRadius = 0.5f * 0.505f * 3.14159f / 180.0f;
    if (vdot(worldDir, sunDirection) > cosf(SunRadius)) {
        float t = raySphereIntersectNearest(worldPos, worldDir, V(0, 0, 0), A->bottomRadius);
        if (t < 0.0f) {
            float uUp, vUp;
            lutTransmittanceParamsToUv(A->bottomRadius, 1.0f, &uUp, &vUp, A->bottomRadius, A->topRadius);
            float pHeight = vlen(worldPos);
            v3 upVector = vdivs(worldPos, pHeight);
            float sunZenithCosAngle = vdot(sunDirection, upVector);
            float uSun, vSun;
            lutTransmittanceParamsToUv(pHeight, sunZenithCosAngle, &uSun, &vSun, A->bottomRadius, A->topRadius);
            float cosAngle = vdot(worldDir, sunDirection);
            float angle = acosf(clampf(cosAngle, -1.0f, 1.0f));
            float radiusRatio = angle / SunRadius;
            float limbDarkening = sqrtf(clampf(1.0f - radiusRatio * radiusRatio, 0.0001f, 1.0f));
            v3 sunLuminanceInSpace = vdiv(sunIlluminance, tex_bilinear(&S->trans, uUp, vUp, 0));
            return vscale(vmul(sunLuminanceInSpace, tex_bilinear(&S->trans, uSun, vSun, 0)), limbDarkening);
        }
    }
    return V(0, 0, 0);
}

/* src/gpu/secondaryRays.comp:36-58; `pos` is C.cameraPos there (the SKY_AT_HIT extension passes the ray origin),
 * parts: bit 0 the sky-view LUT term, bit 1 the sun disc */
static v3 sky_color_at(const sky_ctx* S, v3 pos, v3 dir, int parts) {
    const orc_atmosphere_params* A = S->A;
    v3 worldPos = vadd(pos, V(0.0f, 0.0f, A->bottomRadius));
    v3 upVector = vnorm(worldPos);
    float viewZenithCosAngle = vdot(dir, upVector);
    float viewHeight = vlen(worldPos);
    const v3 sunDirection = V(-0.435286462f, 0.818654716f, 0.374606609f);
    const v3 sunIlluminance = V(8.0f, 8.0f, 8.0f);
    v3 sideVector = vnorm(vcross(upVector, dir));
    v3 forwardVector = vnorm(vcross(sideVector, upVector));
    float lx = vdot(sunDirection, forwardVector), ly = vdot(sunDirection, sideVector);
    float ll = sqrtf(lx * lx + ly * ly);
    lx = lx / ll;
    float lightViewCosAngle = lx;
    int intersectGround = raySphereIntersectNearest(worldPos, dir, V(0, 0, 0), A->bottomRadius) >= 0.0f;
    float u, v;
    skyViewLutParamsToUv(intersectGround, viewZenithCosAngle, lightViewCosAngle, (float)S->view.w,
                         (float)S->view.h, viewHeight, &u, &v, A->bottomRadius);
    v3 skyView = tex_bilinear(&S->view, u, v, 1);
    if (parts == 1) return skyView;
    v3 sun = vmul(getSunLuminance(S, worldPos, dir, sunDirection, sunIlluminance),
                  vdiv(vsplat(120000.0f), sunIlluminance));
    return vadd(skyView, sun);
}
static v3 sky_color(const sky_ctx* S, v3 dir) { return sky_color_at(S, S->cameraPos, dir, 3); }

/* ---- SURVEY 8f-4 extension (no reference counterpart; contract written here): the sun as a sampled light ----
 * Radiance of the sun's centre seen from `pos` through the atmosphere: getSunLuminance's expression without the disc
 * test and the limb factor, times the 120000 / illuminance scale of skyColor (secondaryRays.comp:54-55).  Zero when the
 * sun's centre direction meets the ground sphere. */
static v3 sun_centre_radiance(const sky_ctx* S, v3 pos) {
    const orc_atmosphere_params* A = S->A;
    const v3 sunDirection = V(-0.435286462f, 0.818654716f, 0.374606609f);
    const v3 sunIlluminance = V(8.0f, 8.0f, 8.0f);
    v3 worldPos = vadd(pos, V(0.0f, 0.0f, A->bottomRadius));
    if (raySphereIntersectNearest(worldPos, sunDirection, V(0, 0, 0), A->bottomRadius) >= 0.0f) return V(0, 0, 0);
    float uUp, vUp, uSun, vSun;
    lutTransmittanceParamsToUv(A->bottomRadius, 1.0f, &uUp, &vUp, A->bottomRadius, A->topRadius);
    float pHeight = vlen(worldPos);
    v3 upVector = vdivs(worldPos, pHeight);
    float sunZenithCosAngle = vdot(sunDirection, upVector);
    lutTransmittanceParamsToUv(pHeight, sunZenithCosAngle, &uSun, &vSun, A->bottomRadius, A->topRadius);
    v3 inSpace = vdiv(sunIlluminance, tex_bilinear(&S->trans, uUp, vUp, 0));
    return vmul(vmul(inSpace, tex_bilinear(&S->trans, uSun, vSun, 0)), vdiv(vsplat(120000.0f), sunIlluminance));
}

/* One sun sample from two uniform numbers: direction l uniform in solid angle inside the sun's cone (half-angle
 * 0.5 * 0.505 deg, the disc of getSunLuminance), and the scalar weight  limb(u0) * Omega / pi  that multiplies
 * throughput * sun_centre_radiance * max(0, n.l):  limb darkening sqrt(clamp(1 - r^2, 1e-4, 1)) with r^2 = u0 (the
 * squared radius ratio of a cone sample with cos(theta) = 1 - u0 (1 - cos R), to O(R^2)), Omega = 2 pi (1 - cos R). */
static void nee_sun_sample(float u0, float u1, v3* l, float* weight) {
    const v3 sun = V(-0.435286462f, 0.818654716f, 0.374606609f);
    const float SunRadius = 0.5f * 0.505f * 3.14159f / 180.0f;
    const float oneMinusCos = 1.0f - cosf(SunRadius);
    float cosT = 1.0f - u0 * oneMinusCos;
    float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
    float phi = (u1 * 2.0f) * 3.14159274101257324f;
    v3 t = vnorm(vcross(V(0.0f, 0.0f, 1.0f), sun)); /* the sun is never vertical: |sun.z| = 0.37 */
    v3 b = vcross(sun, t);
    v3 dir = vadd(vadd(vscale(t, cosf(phi) * sinT), vscale(b, sinf(phi) * sinT)), vscale(sun, cosT));
    *l = vnorm(dir);
    float limb = sqrtf(clampf(1.0f - u0, 0.0001f, 1.0f));
    *weight = limb * ((2.0f * oneMinusCos)); /* Omega / pi = 2 (1 - cos R) */
}


static sky_ctx sky_ctx_make(const orc_atmosphere_params* A, const uint16_t* trans, const uint32_t* skyView,
                            const float cameraPos[3]) {
    sky_ctx S;
    S.A = A;
    S.trans = tex_from_rgba16f(trans, ORC_TRANS_W, ORC_TRANS_H);
    S.view = tex_from_b10g11r11(skyView, ORC_VIEW_W, ORC_VIEW_H);
    S.cameraPos = v3p(cameraPos);
    return S;
}
static void sky_ctx_free(sky_ctx* S) { tex_free(&S->trans); tex_free(&S->view); }

void orc_nee_sun_sample(float u0, float u1, float l[3], float* weight) {
    v3 d;
    nee_sun_sample(u0, u1, &d, weight);
    l[0] = d.x; l[1] = d.y; l[2] = d.z;
}
void orc_sun_centre_radiance(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView, const float pos[3],
                             float out[3]) {
    sky_ctx S = sky_ctx_make(p, trans, skyView, pos);
    v3 e = sun_centre_radiance(&S, v3p(pos));
    out[0] = e.x; out[1] = e.y; out[2] = e.z;
    sky_ctx_free(&S);
}

void orc_sky_color(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView,
                   const float cameraPos[3], const float dir[3], float out[3]) {
    sky_ctx S = sky_ctx_make(p, trans, skyView, cameraPos);
    v3 c = sky_color(&S, v3p(dir));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    sky_ctx_free(&S);
}

void orc_sky_color_batch(const orc_atmosphere_params* p, const uint16_t* trans, const uint32_t* skyView,
                         const float cameraPos[3], uint32_t n, const float* dirs, float* out) {
    sky_ctx S = sky_ctx_make(p, trans, skyView, cameraPos);
    for (uint32_t i = 0; i < n; i++) {
        v3 c = sky_color(&S, v3p(dirs + 3 * (size_t)i));
        out[3 * (size_t)i] = c.x; out[3 * (size_t)i + 1] = c.y; out[3 * (size_t)i + 2] = c.z;
    }
    sky_ctx_free(&S);
}

/* =============================== sphere path, faithful (a2-a12) =============================== */

static void store_f16x4(uint16_t* dst, float a, float b, float c, float d) {
    dst[0] = orc_f32_to_f16(a); dst[1] = orc_f32_to_f16(b); dst[2] = orc_f32_to_f16(c); dst[3] = orc_f32_to_f16(d);
}

/* G-buffer depth / motion of a hit position (src/gpu/primaryRay.comp:62-64,73-75) */
static void project_hit(const orc_primary_constants* c, const orc_mat4* PV, const orc_mat4* PVprev, v3 pos,
                        uint32_t w, uint32_t h, float* depth, float motion[2]) {
    (void)c;
    v4 ph = mat_vec(PV, pos.x, pos.y, pos.z, 1.0f);
    ph.x /= ph.w; ph.y /= ph.w; ph.z /= ph.w;
    *depth = ph.z;
    v4 pp = mat_vec(PVprev, pos.x, pos.y, pos.z, 1.0f);
    pp.x /= pp.w; pp.y /= pp.w;
    float pitchx = 1.0f / (float)w, pitchy = 1.0f / (float)h;
    motion[0] = (ph.x - pp.x) / pitchx;
    motion[1] = (ph.y - pp.y) / pitchy;
}

/* src/gpu/primaryRay.comp:23-76.  On a miss the reference leaves t/position uninitialised
 * (UB, :33-34); the oracle defines depth = 0 (infinitely far in inverted-Z) and motion = 0. */
void orc_primary_rays_spheres(uint32_t w, uint32_t h, const orc_primary_constants* c, const orc_sphere* spheres,
                              uint32_t nspheres, uint32_t* visibility, uint16_t* depth, uint16_t* normal,
                              uint16_t* motion) {
    orc_mat4 PV, PVprev;
    orc_mat_mul(&c->projection, &c->view, &PV);
    orc_mat_mul(&c->projection, &c->prevView, &PVprev);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < (int)h; y++)
        for (uint32_t x = 0; x < w; x++) {
            v3 o, d;
            ray_gen(&c->invView, &c->invProjection, x, (uint32_t)y, w, h, &o, &d);
            uint32_t id = NONE_ID;
            float best = 0.0f;
            for (uint32_t i = 0; i < nspheres; i++) {
                float t = ray_sphere(o, d, &spheres[i]);
                if (t >= 0.0f && (id == NONE_ID || t < best)) { best = t; id = i; }
            }
            size_t px = (size_t)y * w + x;
            visibility[px] = id;
            float dep = 0.0f, mo[2] = {0.0f, 0.0f};
            v3 n = d;
            if (id != NONE_ID) {
                v3 pos = vadd(o, vscale(d, best));
                n = vnorm(vsub(pos, v3p(spheres[id].center)));
                project_hit(c, &PV, &PVprev, pos, w, h, &dep, mo);
            }
            depth[px] = orc_f32_to_f16(dep);
            store_f16x4(&normal[4 * px], n.x, n.y, n.z, 0.0f);
            motion[2 * px] = orc_f32_to_f16(mo[0]);
            motion[2 * px + 1] = orc_f32_to_f16(mo[1]);
        }
}

/* src/gpu/secondaryRays.comp:64-135 with Samples / Bounces (:128-129) as parameters */
void orc_secondary_rays_spheres(uint32_t w, uint32_t h, const orc_secondary_constants* c,
                                const orc_sphere* spheres, uint32_t nspheres, const uint32_t* visibility,
                                const uint16_t* depth, const uint16_t* normal, const uint8_t* blueNoise,
                                uint32_t bnW, uint32_t bnH, const orc_atmosphere_params* atmo,
                                const uint16_t* trans, const uint32_t* skyView, uint32_t spp, uint32_t bounces,
                                uint16_t* color16, float* color32, uint64_t* rays_out) {
    sky_ctx S = sky_ctx_make(atmo, trans, skyView, c->cameraPos);
    uint64_t rays = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays)
    for (int y = 0; y < (int)h; y++)
        for (uint32_t x = 0; x < w; x++) {
            size_t px = (size_t)y * w + x;
            float pitchx = 1.0f / (float)w, pitchy = 1.0f / (float)h;
            float u = ((float)x + 0.5f) * pitchx;
            float v = ((float)y + 0.5f) * pitchy;
            v = 1.0f - v;
            v4 co = mat_vec(&c->invView, 0.0f, 0.0f, 0.0f, 1.0f);
            /* reconstruct primary hit (:114-123) */
            uint32_t pid = visibility[px];
            v3 pn = V(orc_f16_to_f32(normal[4 * px]), orc_f16_to_f32(normal[4 * px + 1]),
                      orc_f16_to_f32(normal[4 * px + 2]));
            float dep = orc_f16_to_f32(depth[px]);
            v4 vp = mat_vec(&c->invProjection, u * 2.0f - 1.0f, v * 2.0f - 1.0f, dep, 1.0f);
            vp.x /= vp.w; vp.y /= vp.w; vp.z /= vp.w;
            v4 wp = mat_vec(&c->invView, vp.x, vp.y, vp.z, 1.0f);
            v3 ppos = V(wp.x, wp.y, wp.z);
            (void)co;
            uint32_t seed = (c->frameCounter << 1u) | 1u;
            const uint8_t* bn = &blueNoise[4 * ((size_t)((uint32_t)y % bnH) * bnW + (x % bnW))];
            float rotx = (float)bn[0] / 255.0f, roty = (float)bn[1] / 255.0f;

            v3 color = V(0, 0, 0);
            for (uint32_t s = 0; s < spp; s++) {
                v3 thr = V(1, 1, 1);
                uint32_t hid = pid;
                v3 hpos = ppos, hn = pn;
                v3 contrib = V(0, 0, 0);
                for (uint32_t i = 0; i < bounces + 1u; i++) {
                    if (i > 0) {
                        v3 ro = vadd(hpos, vscale(hn, 0.000001f));
                        float r0 = rotated_random(&seed, rotx);
                        float r1 = rotated_random(&seed, roty);
                        v3 rd = vnorm(vadd(hn, random_sphere_point(r0 * 2.0f - 1.0f, r1 * 2.0f - 1.0f)));
                        rays++;
                        hid = NONE_ID;
                        float ht = -1.0f;
                        for (uint32_t k = 0; k < nspheres; k++) {
                            float t = ray_sphere(ro, rd, &spheres[k]);
                            if (t >= 0.0f && (t < ht || ht < 0.0f)) { hid = k; ht = t; }
                        }
                        if (hid != NONE_ID) {
                            hpos = vadd(ro, vscale(rd, ht));
                            hn = vnorm(vsub(hpos, v3p(spheres[hid].center)));
                        } else {
                            hn = rd;
                        }
                    }
                    if (hid != NONE_ID) {
                        thr = vmul(thr, v3p(spheres[hid].albedo));
                    } else {
                        contrib = vmul(thr, sky_color(&S, hn));
                        break;
                    }
                }
                color = vadd(color, contrib);
            }
            color = vdivs(color, (float)spp);
            if (color16) store_f16x4(&color16[4 * px], color.x, color.y, color.z, 1.0f);
            if (color32) {
                color32[4 * px] = color.x; color32[4 * px + 1] = color.y; color32[4 * px + 2] = color.z;
                color32[4 * px + 3] = 1.0f;
            }
        }
    if (rays_out) *rays_out = rays;
    sky_ctx_free(&S);
}

/* =============================== bilateral denoiser =============================== */

typedef struct { int w, h, ch; const float* px; } texn;

/* texture(tex, uv) with the LinearClamp sampler (src/gfx/modules/denoiser.ixx:72-74,
 * src/gfx/samplers.ixx:14-19), ch channels.  Filtering rule fixed for THIS stage: the fractional
 * texel position is held with 8 fractional bits (VkPhysicalDeviceLimits::subTexelPrecisionBits,
 * 8 on every desktop implementation), weights are k/256, and a texel whose weight is zero is not
 * read.  The shader's taps are built to land on texel centres in x (d.x is integral) and the
 * colour image holds +inf on the sun disc (fp16 overflow of 1.2e5 nits): with unquantised fp32
 * weights the 1e-4-texel rounding residue of uv + d/size would decide, tap by tap, between inf,
 * NaN (0 * inf) and a finite value, which no GPU does.  (The sky LUT lookups keep the plain fp32
 * lerp of tex_bilinear: they are smooth and finite, so weight precision is immaterial there.) */
static void texn_bilinear(const texn* t, float u, float v, float* out) {
    float x = u * (float)t->w - 0.5f, y = v * (float)t->h - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = rintf((x - fx0) * 256.0f) * (1.0f / 256.0f), fy = rintf((y - fy0) * 256.0f) * (1.0f / 256.0f);
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 > t->w - 1 ? t->w - 1 : x0);
    x1 = x1 < 0 ? 0 : (x1 > t->w - 1 ? t->w - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 > t->h - 1 ? t->h - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > t->h - 1 ? t->h - 1 : y1);
    const float* p00 = &t->px[(size_t)t->ch * ((size_t)y0 * t->w + x0)];
    const float* p10 = &t->px[(size_t)t->ch * ((size_t)y0 * t->w + x1)];
    const float* p01 = &t->px[(size_t)t->ch * ((size_t)y1 * t->w + x0)];
    const float* p11 = &t->px[(size_t)t->ch * ((size_t)y1 * t->w + x1)];
    float gx = 1.0f - fx, gy = 1.0f - fy;
    for (int c = 0; c < t->ch; c++) {
        float top = fx == 0.0f ? p00[c] : (gx == 0.0f ? p10[c] : p00[c] * gx + p10[c] * fx);
        float bot = fx == 0.0f ? p01[c] : (gx == 0.0f ? p11[c] : p01[c] * gx + p11[c] * fx);
        out[c] = fy == 0.0f ? top : (gy == 0.0f ? bot : top * gy + bot * fy);
    }
}

/* smartDeNoise + main of src/gpu/denoise/bilateral.comp:23-76, push constants of
 * src/gfx/modules/denoiser.ixx:78-91 (defaults sigma 5, kSigma 2, threshold 0.12, :27-33).
 * Inputs are the reference's images: colour RGBA16F, depth R16F, normal RGBA16F; output RGBA8 unorm
 * (denoiser.ixx:56), i.e. the HDR colour is clamped to [0,1] BEFORE the tonemapper sees it.
 * Fixed here where GLSL leaves it open: round() = roundf (half away from zero); clamp(NaN,0,1) = 0
 * (inf - inf arises where both taps are sky: depth 0 in inverted Z). */
void orc_denoise_bilateral(uint32_t w, uint32_t h, const uint16_t* color16, const uint16_t* depth16,
                           const uint16_t* normal16, float sigma, float kSigma, float threshold,
                           float nearPlane, uint32_t frameCounter, uint8_t* rgba8) {
    size_t n = (size_t)w * h;
    float* col = (float*)malloc(sizeof(float) * 4 * n);
    float* dep = (float*)malloc(sizeof(float) * n);
    float* nor = (float*)malloc(sizeof(float) * 3 * n);
    for (size_t i = 0; i < n; i++) {
        for (int c = 0; c < 4; c++) col[4 * i + c] = orc_f16_to_f32(color16[4 * i + c]);
        dep[i] = orc_f16_to_f32(depth16[i]);
        for (int c = 0; c < 3; c++) nor[3 * i + c] = orc_f16_to_f32(normal16[4 * i + c]);
    }
    const texn tc = {(int)w, (int)h, 4, col}, tz = {(int)w, (int)h, 1, dep}, tn = {(int)w, (int)h, 3, nor};
    const float INV_SQRT_OF_2PI = 0.39894228040143267793994605993439f;
    const float INV_PI = 0.31830988618379067153776752674503f;
    const float radius = roundf(kSigma * sigma);
    const float radQ = radius * radius;
    const float invSigmaQx2 = .5f / (sigma * sigma);
    const float invSigmaQx2PI = INV_PI * invSigmaQx2;
    const float invThresholdSqx2 = .5f / (threshold * threshold);
    const float invThresholdSqrt2PI = INV_SQRT_OF_2PI / threshold;
    const float sizeX = (float)w, sizeY = (float)h;
#pragma omp parallel for schedule(dynamic, 4)
    for (long long gy = 0; gy < (long long)h; gy++)
        for (uint32_t gx = 0; gx < w; gx++) {
            float uvx = ((float)gx + 0.5f) / sizeX, uvy = ((float)gy + 0.5f) / sizeY;
            float filtered[4], centrPx[4], centrZ, centrN[3];
            texn_bilinear(&tc, uvx, uvy, centrPx);
            texn_bilinear(&tz, uvx, uvy, &centrZ);
            if (centrZ < 0.0f) {
                memcpy(filtered, centrPx, sizeof filtered);
            } else {
                texn_bilinear(&tn, uvx, uvy, centrN);
                float zBuff = 0.0f, aBuff[4] = {0, 0, 0, 0};
                for (float dx = -radius; dx <= radius; dx++) {
                    float pt = sqrtf(radQ - dx * dx);
                    for (float dy = -pt; dy <= pt; dy++) {
                        float blurFactor = expf(-(dx * dx + dy * dy) * invSigmaQx2) * invSigmaQx2PI;
                        float u = uvx + dx / sizeX, v = uvy + dy / sizeY;
                        float walkPx[4], walkZ, walkN[3];
                        texn_bilinear(&tc, u, v, walkPx);
                        texn_bilinear(&tz, u, v, &walkZ);
                        texn_bilinear(&tn, u, v, walkN);
                        float dZ = nearPlane / walkZ - nearPlane / centrZ;
                        dZ *= 100.0f;
                        float dN = walkN[0] * centrN[0] + walkN[1] * centrN[1] + walkN[2] * centrN[2];
                        float deltaFactor = expf(clampf(dN - dZ * dZ, 0.0f, 1.0f) * invThresholdSqx2) *
                                            invThresholdSqrt2PI * blurFactor;
                        zBuff += deltaFactor;
                        for (int c = 0; c < 4; c++) aBuff[c] += deltaFactor * walkPx[c];
                    }
                }
                for (int c = 0; c < 4; c++) filtered[c] = aBuff[c] / zBuff;
            }
            uint32_t seed = gx * 709u + (uint32_t)gy * 1153u + frameCounter * 1361u;
            float noise = orc_random_float(&seed) * 0.005f;
            size_t i = (size_t)gy * w + gx;
            rgba8[4 * i] = orc_unorm8(filtered[0] + noise);
            rgba8[4 * i + 1] = orc_unorm8(filtered[1] + noise);
            rgba8[4 * i + 2] = orc_unorm8(filtered[2] + noise);
            rgba8[4 * i + 3] = orc_unorm8(filtered[3]);
        }
    free(col); free(dep); free(nor);
}

/* =============================== tonemap (a14) =============================== */

/* src/gpu/util.glsl:13-18 */
static float srgb1(float c) { return c < 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }

/* src/gpu/tonemap/amd.comp:22-40 */
static float ColToneB(float hdrMax, float contrast, float shoulder, float midIn, float midOut) {
    return -((-powf(midIn, contrast) +
              (midOut * (powf(hdrMax, contrast * shoulder) * powf(midIn, contrast) -
                         powf(hdrMax, contrast) * powf(midIn, contrast * shoulder) * midOut)) /
                  (powf(hdrMax, contrast * shoulder) * midOut - powf(midIn, contrast * shoulder) * midOut)) /
             (powf(midIn, contrast * shoulder) * midOut));
}
static float ColToneC(float hdrMax, float contrast, float shoulder, float midIn, float midOut) {
    return (powf(hdrMax, contrast * shoulder) * powf(midIn, contrast) -
            powf(hdrMax, contrast) * powf(midIn, contrast * shoulder) * midOut) /
           (powf(hdrMax, contrast * shoulder) * midOut - powf(midIn, contrast * shoulder) * midOut);
}
static float ColTone(float x, float p0, float p1, float p2, float p3) {
    float z = powf(x, p0);
    return z / (powf(z, p1) * p2 + p3);
}
static float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }

/* src/gpu/tonemap/amd.comp:42-69 ; params {hdrMax, contrast, shoulder, midIn, midOut} */
static v3 tm_amd(v3 color, const float* p) {
    float hdrMax = p[0], contrast = p[1], shoulder = p[2], midIn = p[3], midOut = p[4];
    float b = ColToneB(hdrMax, contrast, shoulder, midIn, midOut);
    float c = ColToneC(hdrMax, contrast, shoulder, midIn, midOut);
    float peak = fmaxf(color.x, fmaxf(color.y, color.z));
    peak = fmaxf(1e-6f, peak);
    v3 ratio = vdivs(color, peak);
    peak = ColTone(peak, contrast, shoulder, b, c);
    float crosstalk = 4.0f;
    float saturation = contrast;
    float crossSaturation = contrast * 16.0f;
    float white = 1.0f;
    float e0 = saturation / crossSaturation;
    ratio = V(powf(fabsf(ratio.x), e0), powf(fabsf(ratio.y), e0), powf(fabsf(ratio.z), e0));
    float a = powf(peak, crosstalk);
    ratio = V(mixf(ratio.x, white, a), mixf(ratio.y, white, a), mixf(ratio.z, white, a));
    ratio = V(powf(fabsf(ratio.x), crossSaturation), powf(fabsf(ratio.y), crossSaturation),
              powf(fabsf(ratio.z), crossSaturation));
    return vscale(ratio, peak);
}

/* src/gpu/tonemap/reinhard.comp:16-19 */
static v3 tm_reinhard(v3 v, float max_white) {
    v3 mw = vsplat(max_white * max_white);
    v3 numerator = vmul(v, vadd(vsplat(1.0f), vdiv(v, mw)));
    return vdiv(numerator, vadd(vsplat(1.0f), v));
}

/* src/gpu/tonemap/hable.comp:14-28 */
static float hable1(float x) {
    float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
static v3 tm_hable(v3 c) {
    float W = 11.2f;
    float d = hable1(W);
    return V(hable1(2.0f * c.x) / d, hable1(2.0f * c.y) / d, hable1(2.0f * c.z) / d);
}

/* src/gpu/tonemap/aces.comp:15-43.  GLSL `v * M` treats v as a row vector: result[i] = dot(v, M[i]),
 * M[i] being the i-th brace group (a column in GLSL's constructor order). */
static v3 row_mul(v3 v, const float M[3][3]) {
    return V(v.x * M[0][0] + v.y * M[0][1] + v.z * M[0][2], v.x * M[1][0] + v.y * M[1][1] + v.z * M[1][2],
             v.x * M[2][0] + v.y * M[2][1] + v.z * M[2][2]);
}
static v3 tm_aces(v3 color) {
    static const float In[3][3] = {{0.59719f, 0.35458f, 0.04823f}, {0.07600f, 0.90834f, 0.01566f},
                                   {0.02840f, 0.13383f, 0.83777f}};
    static const float Out[3][3] = {{1.60475f, -0.53108f, -0.07367f}, {-0.10208f, 1.10813f, -0.00605f},
                                    {-0.00327f, -0.07276f, 1.07602f}};
    color = row_mul(color, In);
    v3 a = vsub(vmul(color, vadd(color, vsplat(0.0245786f))), vsplat(0.000090537f));
    v3 b = vadd(vmul(color, vadd(vscale(color, 0.983729f), vsplat(0.4329510f))), vsplat(0.238081f));
    color = vdiv(a, b);
    color = row_mul(color, Out);
    return V(clampf(color.x, 0.0f, 1.0f), clampf(color.y, 0.0f, 1.0f), clampf(color.z, 0.0f, 1.0f));
}

/* src/gpu/tonemap/uchimura.comp:20-38 ; params {P, a, m, l, c, b} */
static float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
static float uchimura1(float x, float P, float a, float m, float l, float c, float b) {
    float l0 = ((P - m) * l) / a;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = (a * P) / (P - S1);
    float CP = -C2 / P;
    float w0 = 1.0f - smoothstepf(0.0f, m, x);
    float w2 = (x < m + l0) ? 0.0f : 1.0f; /* step(edge, x) */
    float w1 = 1.0f - w0 - w2;
    float T = m * powf(x / m, c) + b;
    float Sc = P - (P - S1) * expf(CP * (x - S0));
    float Lc = m + a * (x - m);
    return T * w0 + Lc * w1 + Sc * w2;
}

void orc_ray_triangle_batch(uint32_t n, const float* o, const float* d, const float v0[3], const float v1[3],
                            const float v2[3], uint8_t* hit, float* t) {
    for (uint32_t i = 0; i < n; i++) {
        float u, v;
        t[i] = 0.0f;
        hit[i] = (uint8_t)orc_ray_triangle(o + 3 * i, d + 3 * i, v0, v1, v2, &t[i], &u, &v);
    }
}

/* ---- compressed 8-wide BVH node (contract in minote_oracle.h) ---- */
static inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* bvh_build.cu grid_exponent: biased exponent e with 2^(e-127) * 250 >= ext, and never below 2 ulp of the largest
 * coordinate magnitude m on the axis (nor below 2^-126): a finer grid could not be resolved by fp32 coordinates, and
 * the grid origin nlo - 2 step would round by more than a fraction of a step */
static uint32_t wide_grid_exponent(float ext, float m) {
    float s = ext / 250.0f;
    uint32_t b = f_bits(s);
    uint32_t e = (b >> 23) & 0xFFu;
    if (b & 0x7FFFFFu) e += 1;
    while (e < 254u && bits_f(e << 23) * 250.0f < ext) e++;
    const uint32_t em = (f_bits(m) >> 23) & 0xFFu;
    const uint32_t emin = em > 23u ? em - 22u : 1u;
    if (e < emin) e = emin;
    return e > 254u ? 254u : e;
}

/* bvh_build.cu k_emit_nodes, the quantisation part.  Plane positions origin + q * step are evaluated in double
 * (exact for fp32 operands), because that is the plane the traversal's (origin - o) / d + q * (step / d) sees. */
void orc_wide_node_quantize(const float lo[8][3], const float hi[8][3], uint32_t present, orc_wide_node* out) {
    float nlo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, nhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int s = 0; s < 8; s++) {
        if (!(present & (1u << s))) continue;
        for (int a = 0; a < 3; a++) {
            nlo[a] = fminf(nlo[a], lo[s][a]);
            nhi[a] = fmaxf(nhi[a], hi[s][a]);
        }
    }
    uint32_t e[3];
    float sc[3], org[3];
    for (int a = 0; a < 3; a++) {
        e[a] = wide_grid_exponent(nhi[a] - nlo[a], fmaxf(fabsf(nlo[a]), fabsf(nhi[a])));
        sc[a] = bits_f(e[a] << 23);
        org[a] = nlo[a] - 2.0f * sc[a];
    }
    memset(out, 0, sizeof *out);
    for (int a = 0; a < 3; a++) out->w[a] = f_bits(org[a]);
    out->w[3] = e[0] | (e[1] << 8) | (e[2] << 16);
    for (int s = 0; s < 8; s++) {
        for (int a = 0; a < 3; a++) {
            float ql = 255.0f, qh = 0.0f; /* empty slot: inverted box */
            if (present & (1u << s)) {
                const double o = (double)org[a], st = (double)sc[a], slack = st * 0.015625;
                ql = fminf(fmaxf(floorf((lo[s][a] - org[a]) / sc[a] - 0.02f), 0.0f), 255.0f);
                qh = fminf(fmaxf(ceilf((hi[s][a] - org[a]) / sc[a] + 0.02f), 0.0f), 255.0f);
                while (ql > 0.0f && o + (double)ql * st > (double)lo[s][a] - slack) ql -= 1.0f;
                while (qh < 255.0f && o + (double)qh * st < (double)hi[s][a] + slack) qh += 1.0f;
            }
            out->w[4 + 2 * a + (s >> 2)] |= (uint32_t)ql << (8 * (s & 3));
            out->w[10 + 2 * a + (s >> 2)] |= (uint32_t)qh << (8 * (s & 3));
        }
    }
}

/* trace.cuh lane_begin + lane_node_step (TRACE_DP4A_NEAR = 1) */
void orc_wide_node_test(const orc_wide_node* node, uint32_t n, const float* o, const float* d, const float* t_best,
                        uint32_t* hits) {
    const float KNEAR = 0.99999952f, KFAR = 1.00000048f, tiny = 1e-20f;
    const uint32_t K = 0x47000000u; /* 2^15 */
    for (uint32_t r = 0; r < n; r++) {
        float idn[3], idf[3];
        int pos[3];
        for (int a = 0; a < 3; a++) {
            float da = d[3 * r + a];
            float dd = fabsf(da) > tiny ? da : copysignf(tiny, da);
            float idir = 1.0f / dd;
            idn[a] = idir * KNEAR;
            idf[a] = idir * KFAR;
            pos[a] = !(idir < 0.0f);
        }
        const float tlimit = t_best[r] >= 3.0e38f ? 3.0e38f : t_best[r] * KFAR;
        float sn2[3], bn[3], sf[3], bf[3];
        for (int a = 0; a < 3; a++) {
            const float st = bits_f(((node->w[3] >> (8 * a)) & 0xFFu) << 23);
            const float dx = bits_f(node->w[a]) - o[3 * r + a];
            const float sn = st * idn[a];
            sf[a] = st * idf[a];
            bn[a] = fmaf(-65536.0f, sn, dx * idn[a]);
            sn2[a] = sn * 2.0f;
            bf[a] = fmaf(-32768.0f, sf[a], dx * idf[a]);
        }
        uint32_t mask = 0;
        for (int s = 0; s < 8; s++) {
            float t0[3], t1[3];
            for (int a = 0; a < 3; a++) {
                const uint32_t ql = (node->w[4 + 2 * a + (s >> 2)] >> (8 * (s & 3))) & 0xFFu;
                const uint32_t qh = (node->w[10 + 2 * a + (s >> 2)] >> (8 * (s & 3))) & 0xFFu;
                const uint32_t qn = pos[a] ? ql : qh, qf = pos[a] ? qh : ql;
                t0[a] = fmaf(bits_f(K + 128u * qn), sn2[a], bn[a]); /* IDP.4A decode: 2^15 + q/2 */
                t1[a] = fmaf(bits_f(K | (qf << 8)), sf[a], bf[a]);  /* PRMT decode: 2^15 + q */
            }
            const float tmin = fmaxf(fmaxf(t0[0], t0[1]), t0[2]);
            const float tmax = fminf(fminf(t1[0], t1[1]), t1[2]);
            const uint32_t neg = f_bits(tmax - tmin) | f_bits(tlimit - tmin) | f_bits(tmax);
            if (!(neg >> 31)) mask |= 1u << s;
        }
        hits[r] = mask;
    }
}

/* ---- temporal reprojection (contract in minote_oracle.h) ----
 * Consumes the motion buffer of src/gpu/primaryRay.comp:73-75.  fp32, no contraction; the same expression order as
 * k_temporal (minotert_b200/csrc/temporal.cu), so the two agree bit for bit. */
void orc_temporal_accumulate(uint32_t w, uint32_t h, const float* accum, const uint32_t* vis, const uint16_t* motion16,
                             int have_history, const float* hist_rgba, const float* hist_count, const uint32_t* hist_vis,
                             float maxHistory, float* out_rgba, float* out_count) {
#pragma omp parallel for schedule(static)
    for (long long y = 0; y < (long long)h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const size_t p = (size_t)y * w + x;
            const float aw = accum[4 * p + 3];
            float cur[3];
            for (int c = 0; c < 3; c++) cur[c] = aw > 0.0f ? accum[4 * p + c] / aw : 0.0f;
            const uint32_t id = vis[p];
            float out[3] = {cur[0], cur[1], cur[2]}, count = 1.0f;
            if (have_history && id != 0xFFFFFFFFu) {
                const float mx = orc_f16_to_f32(motion16[2 * p]), my = orc_f16_to_f32(motion16[2 * p + 1]);
                /* texel-centre coordinates of the previous position, minus the half texel of the bilinear footprint */
                const float gx = ((float)x + 0.5f - mx * 0.5f) - 0.5f, gy = ((float)y + 0.5f + my * 0.5f) - 0.5f;
                /* NaN / far outside (also what an inf motion gives): no history */
                if (gx > -2.0f && gy > -2.0f && gx < (float)w + 1.0f && gy < (float)h + 1.0f) {
                    const float fx0 = floorf(gx), fy0 = floorf(gy);
                    const float wx = gx - fx0, wy = gy - fy0;
                    const int x0 = (int)fx0, y0 = (int)fy0;
                    float sum[3] = {0.0f, 0.0f, 0.0f}, nsum = 0.0f, wsum = 0.0f;
                    for (int j = 0; j < 2; j++)
                        for (int i = 0; i < 2; i++) {
                            const int tx = x0 + i, ty = y0 + j;
                            if (tx < 0 || ty < 0 || tx >= (int)w || ty >= (int)h) continue;
                            const size_t q = (size_t)ty * w + (size_t)tx;
                            if (hist_vis[q] != id) continue;
                            const float wt = (i ? wx : 1.0f - wx) * (j ? wy : 1.0f - wy);
                            for (int c = 0; c < 3; c++) sum[c] = sum[c] + wt * hist_rgba[4 * q + c];
                            nsum = nsum + wt * hist_count[q];
                            wsum = wsum + wt;
                        }
                    if (wsum > 0.00390625f) {
                        float n = nsum / wsum;
                        n = n < maxHistory ? n : maxHistory;
                        const float a = 1.0f / (n + 1.0f);
                        for (int c = 0; c < 3; c++) {
                            const float hc = sum[c] / wsum;
                            out[c] = hc + (cur[c] - hc) * a;
                        }
                        count = n + 1.0f;
                    }
                }
            }
            out_rgba[4 * p] = out[0]; out_rgba[4 * p + 1] = out[1]; out_rgba[4 * p + 2] = out[2]; out_rgba[4 * p + 3] = 1.0f;
            out_count[p] = count;
        }
}


void orc_tonemap_pixel(int mode, const float in[3], float exposure, const float* p, float out[3]) {
    v3 src = vscale(v3p(in), exposure);
    v3 mapped;
    switch (mode) {
    case ORC_TONEMAP_LINEAR: mapped = src; break;
    case ORC_TONEMAP_REINHARD: mapped = tm_reinhard(src, p[0]); break;
    case ORC_TONEMAP_HABLE: mapped = tm_hable(src); break;
    case ORC_TONEMAP_ACES: mapped = tm_aces(src); break;
    case ORC_TONEMAP_UCHIMURA:
        mapped = V(uchimura1(src.x, p[0], p[1], p[2], p[3], p[4], p[5]),
                   uchimura1(src.y, p[0], p[1], p[2], p[3], p[4], p[5]),
                   uchimura1(src.z, p[0], p[1], p[2], p[3], p[4], p[5]));
        break;
    default: mapped = tm_amd(src, p); break;
    }
    out[0] = srgb1(mapped.x); out[1] = srgb1(mapped.y); out[2] = srgb1(mapped.z);
}

/* main() of src/gpu/tonemap/<op>.comp ; output RGBA8 unorm (src/gfx/modules/tonemapper.ixx:332) */
void orc_tonemap(int mode, uint32_t w, uint32_t h, const void* src, int src_is_f16, float exposure,
                 const float* params, uint8_t* rgba8) {
    size_t n = (size_t)w * h;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        float in[3], out[3];
        if (src_is_f16 == 2) { /* the denoiser's RGBA8 unorm image (denoiser.ixx:56): texel = k/255 */
            const uint8_t* s = (const uint8_t*)src + 4 * i;
            in[0] = (float)s[0] / 255.0f; in[1] = (float)s[1] / 255.0f; in[2] = (float)s[2] / 255.0f;
        } else if (src_is_f16) {
            const uint16_t* s = (const uint16_t*)src + 4 * i;
            in[0] = orc_f16_to_f32(s[0]); in[1] = orc_f16_to_f32(s[1]); in[2] = orc_f16_to_f32(s[2]);
        } else {
            const float* s = (const float*)src + 4 * i;
            in[0] = s[0]; in[1] = s[1]; in[2] = s[2];
        }
        orc_tonemap_pixel(mode, in, exposure, params, out);
        rgba8[4 * i] = orc_unorm8(out[0]);
        rgba8[4 * i + 1] = orc_unorm8(out[1]);
        rgba8[4 * i + 2] = orc_unorm8(out[2]);
        rgba8[4 * i + 3] = 255;
    }
}

/* =============================== triangle scenes (n1-n7) =============================== */

typedef struct { float lo[3], hi[3]; uint32_t left; /* children left,left+1 */ uint32_t first, count; } bnode;

struct orc_scene {
    uint32_t ntris;
    float* tri;    /* 9 floats per triangle, upload order */
    float* albedo; /* 3 per triangle */
    /* median-split binary BVH over triangle ids */
    bnode* nodes;
    uint32_t nnodes;
    uint32_t* order;
};

static const float* g_sort_cent; /* build is single-threaded */
static int g_sort_axis;
static int cmp_cent(const void* a, const void* b) {
    uint32_t ia = *(const uint32_t*)a, ib = *(const uint32_t*)b;
    float ca = g_sort_cent[3 * ia + g_sort_axis], cb = g_sort_cent[3 * ib + g_sort_axis];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return (ia > ib) - (ia < ib);
}

static void tri_bounds(const float* t, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; k++) {
        lo[k] = fminf(t[k], fminf(t[3 + k], t[6 + k]));
        hi[k] = fmaxf(t[k], fmaxf(t[3 + k], t[6 + k]));
    }
}

static void build_rec(orc_scene* s, const float* cent, uint32_t node, uint32_t first, uint32_t count) {
    bnode* n = &s->nodes[node];
    for (int k = 0; k < 3; k++) { n->lo[k] = INFINITY; n->hi[k] = -INFINITY; }
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = first; i < first + count; i++) {
        float lo[3], hi[3];
        tri_bounds(&s->tri[9 * (size_t)s->order[i]], lo, hi);
        for (int k = 0; k < 3; k++) {
            n->lo[k] = fminf(n->lo[k], lo[k]); n->hi[k] = fmaxf(n->hi[k], hi[k]);
            float c = cent[3 * (size_t)s->order[i] + k];
            clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c);
        }
    }
    n->first = first; n->count = count; n->left = 0;
    if (count <= 4) return;
    int axis = 0;
    float ext = chi[0] - clo[0];
    if (chi[1] - clo[1] > ext) { axis = 1; ext = chi[1] - clo[1]; }
    if (chi[2] - clo[2] > ext) { axis = 2; }
    g_sort_cent = cent; g_sort_axis = axis;
    qsort(&s->order[first], count, sizeof(uint32_t), cmp_cent);
    uint32_t half = count / 2;
    uint32_t left = s->nnodes;
    s->nnodes += 2;
    n->left = left;
    n->count = 0;
    build_rec(s, cent, left, first, half);
    build_rec(s, cent, left + 1, first + half, count - half);
}

orc_scene* orc_scene_create(const float* positions, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                            const float* albedo) {
    (void)nverts;
    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    s->ntris = ntris;
    s->tri = (float*)malloc(sizeof(float) * 9 * (size_t)ntris);
    s->albedo = (float*)malloc(sizeof(float) * 3 * (size_t)ntris);
    memcpy(s->albedo, albedo, sizeof(float) * 3 * (size_t)ntris);
    float* cent = (float*)malloc(sizeof(float) * 3 * (size_t)ntris);
    for (uint32_t i = 0; i < ntris; i++) {
        for (int v = 0; v < 3; v++)
            for (int k = 0; k < 3; k++) s->tri[9 * (size_t)i + 3 * v + k] = positions[3 * (size_t)indices[3 * (size_t)i + v] + k];
        float lo[3], hi[3];
        tri_bounds(&s->tri[9 * (size_t)i], lo, hi);
        for (int k = 0; k < 3; k++) cent[3 * (size_t)i + k] = 0.5f * (lo[k] + hi[k]);
    }
    s->order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(ntris ? ntris : 1));
    for (uint32_t i = 0; i < ntris; i++) s->order[i] = i;
    s->nodes = (bnode*)malloc(sizeof(bnode) * (2 * (size_t)ntris + 2));
    s->nnodes = 1;
    if (ntris) build_rec(s, cent, 0, 0, ntris);
    else { memset(&s->nodes[0], 0, sizeof(bnode)); }
    free(cent);
    return s;
}

void orc_scene_destroy(orc_scene* s) {
    if (!s) return;
    free(s->tri); free(s->albedo); free(s->nodes); free(s->order); free(s);
}

typedef struct { uint32_t id; float t, u, v; } hit_t;

/* closest hit = lexicographic minimum of (t, id) over all accepted hits
 * (tie rule inherited from src/gpu/primaryRay.comp:28: ascending index, strict '<') */
static inline void consider(const orc_scene* s, v3 o, const ray_shear* rs, uint32_t id, hit_t* h) {
    const float* t9 = &s->tri[9 * (size_t)id];
    float t, u, v;
    if (ray_triangle(o, rs, v3p(t9), v3p(t9 + 3), v3p(t9 + 6), &t, &u, &v)) {
        if (h->id == NONE_ID || t < h->t || (t == h->t && id < h->id)) { h->id = id; h->t = t; h->u = u; h->v = v; }
    }
}

static hit_t closest_hit(const orc_scene* s, v3 o, v3 d, int use_bvh) {
    hit_t h = {NONE_ID, 0.0f, 0.0f, 0.0f};
    ray_shear rs = make_shear(d);
    if (!use_bvh) {
        for (uint32_t i = 0; i < s->ntris; i++) consider(s, o, &rs, i, &h);
        return h;
    }
    if (!s->ntris) return h;
    float inv[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    float oo[3] = {o.x, o.y, o.z};
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const bnode* n = &s->nodes[stack[--sp]];
        /* conservative slab test */
        float tn = 0.0f, tf = INFINITY;
        int ok = 1;
        for (int k = 0; k < 3; k++) {
            float t0 = (n->lo[k] - oo[k]) * inv[k], t1 = (n->hi[k] - oo[k]) * inv[k];
            if (t0 != t0 || t1 != t1) { /* 0 * inf: origin on a slab plane of an axis-parallel ray */
                if (oo[k] < n->lo[k] || oo[k] > n->hi[k]) { ok = 0; break; }
                continue;
            }
            float a = fminf(t0, t1), b = fmaxf(t0, t1);
            tn = fmaxf(tn, a);
            tf = fminf(tf, b);
        }
        if (!ok) continue;
        if (tn * 0.9999990f > tf * 1.0000010f) continue;
        if (h.id != NONE_ID && tn * 0.9999990f > h.t) continue;
        if (n->left == 0) {
            for (uint32_t i = n->first; i < n->first + n->count; i++) consider(s, o, &rs, s->order[i], &h);
        } else {
            if (sp + 2 > 128) abort();
            stack[sp++] = n->left;
            stack[sp++] = n->left + 1;
        }
    }
    return h;
}

uint32_t orc_scene_closest_hit(const orc_scene* s, const float o[3], const float d[3], int use_bvh, float* t,
                               float* u, float* v) {
    hit_t h = closest_hit(s, v3p(o), v3p(d), use_bvh);
    if (t) *t = h.t;
    if (u) *u = h.u;
    if (v) *v = h.v;
    return h.id;
}

/* geometric normal, two-sided: flipped to face the incoming ray (row n4) */
static v3 tri_normal(const orc_scene* s, uint32_t id, v3 d) {
    const float* t9 = &s->tri[9 * (size_t)id];
    v3 p0 = v3p(t9), e1 = vsub(v3p(t9 + 3), p0), e2 = vsub(v3p(t9 + 6), p0);
    v3 n = vnorm(vcross(e1, e2));
    if (vdot(n, d) > 0.0f) n = vneg(n);
    return n;
}

void orc_primary_rays_tris(const orc_scene* s, uint32_t w, uint32_t h, const orc_primary_constants* c, int use_bvh,
                           uint32_t y0, uint32_t y1, uint32_t* visibility, uint16_t* depth, uint16_t* normal,
                           uint16_t* motion, float* hit_t_out) {
    orc_mat4 PV, PVprev;
    orc_mat_mul(&c->projection, &c->view, &PV);
    orc_mat_mul(&c->projection, &c->prevView, &PVprev);
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = (int)y0; y < (int)y1; y++)
        for (uint32_t x = 0; x < w; x++) {
            v3 o, d;
            ray_gen(&c->invView, &c->invProjection, x, (uint32_t)y, w, h, &o, &d);
            hit_t hh = closest_hit(s, o, d, use_bvh);
            size_t px = (size_t)y * w + x;
            if (visibility) visibility[px] = hh.id;
            float dep = 0.0f, mo[2] = {0.0f, 0.0f};
            v3 n = d;
            if (hh.id != NONE_ID) {
                v3 pos = vadd(o, vscale(d, hh.t));
                n = tri_normal(s, hh.id, d);
                project_hit(c, &PV, &PVprev, pos, w, h, &dep, mo);
            }
            if (depth) depth[px] = orc_f32_to_f16(dep);
            if (normal) store_f16x4(&normal[4 * px], n.x, n.y, n.z, 0.0f);
            if (motion) { motion[2 * px] = orc_f32_to_f16(mo[0]); motion[2 * px + 1] = orc_f32_to_f16(mo[1]); }
            if (hit_t_out) hit_t_out[px] = hh.id != NONE_ID ? hh.t : 0.0f;
        }
}

/* Native (fp32-hit) path trace of a triangle scene: secondaryRays.comp:64-135 semantics with the
 * primary hit carried in fp32 instead of through the fp16 G-buffer, parametric spp/bounces and an
 * fp32 accumulator (rows n6/n7). */
void orc_render_tris(const orc_scene* s, uint32_t w, uint32_t h, const orc_primary_constants* pc,
                     const orc_secondary_constants* sc, const uint8_t* blueNoise, uint32_t bnW, uint32_t bnH,
                     const orc_atmosphere_params* atmo, const uint16_t* trans, const uint32_t* skyView,
                     uint32_t spp, uint32_t bounces, int use_bvh, uint32_t y0, uint32_t y1, float* accum,
                     uint32_t* visibility, uint64_t* rays_out) {
    orc_render_tris_ext(s, w, h, pc, sc, blueNoise, bnW, bnH, atmo, trans, skyView, spp, bounces, use_bvh, y0, y1, accum,
                        visibility, rays_out, 0u, NULL);
}

/* ---- aerial-perspective volume ----
 * The reference declares the volume (Sky::AerialPerspectiveFormat RGBA16F, AerialPerspectiveSize 32^3, sky.ixx:190-191;
 * AP_KM_PER_SLICE 4.0 and the depth<->slice maps, skyAccess.glsl:9,119-125) and never builds it.  The contract here is the
 * camera-volume pass of the technique its sky code comes from (Hillaire 2020, "A Scalable and Production Ready Sky and
 * Atmosphere Rendering Technique", sec. 5.3), written with the reference's own integrateScatteredLuminance:
 *   froxel (x, y, z): view ray of the centre of cell (x, y) of a 32 x 32 image through ray_gen (primaryRay.comp:40-56),
 *   slice s = ((z + 0.5) / 32)^2 * 32, depth tMax = s * 4 km along that ray from the camera; a froxel below the ground
 *   is pulled onto it; luminance scattered towards the camera and transmittance over [0, tMax] by the ray march with
 *   max(1, 2 (z + 1)) fixed steps, Mie + Rayleigh phase, no ground term  ->  RGBA16F (L.rgb, 1 - mean transmittance).
 * Lookup for a surface seen at distance t through pixel (px, py) of a w x h image: s = t / 4; weight = 1, and for
 * s < 0.5: weight = clamp(2 s, 0, 1), s = 0.5 (fades to nothing at the camera); trilinear, clamp to edge, at
 * ((px + 0.5) / w, (py + 0.5) / h, sqrt(s / 32)); result * weight.  ORC_EXT_AERIAL applies it to every sample of a pixel
 * whose primary ray hit: the path's throughput starts at 1 - AP.a and AP.rgb is added once per sample. */
#define AP_SIZE 32
#define AP_KM_PER_SLICE 4.0f
void orc_gen_aerial_perspective(const orc_atmosphere_params* A, const uint16_t* trans16, const uint16_t* multi16,
                                const orc_mat4* invView, const orc_mat4* invProjection, const float cameraPos[3],
                                const float sunDirection[3], const float sunIlluminance[3], uint16_t* out) {
    tex3 trans = tex_from_rgba16f(trans16, ORC_TRANS_W, ORC_TRANS_H);
    tex3 multi = tex_from_rgba16f(multi16, ORC_MULTI_W, ORC_MULTI_H);
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < AP_SIZE; z++)
        for (int y = 0; y < AP_SIZE; y++)
            for (int x = 0; x < AP_SIZE; x++) {
                v3 o, worldDir;
                ray_gen(invView, invProjection, (uint32_t)x, (uint32_t)y, AP_SIZE, AP_SIZE, &o, &worldDir);
                v3 camPos = vadd(v3p(cameraPos), V(0.0f, 0.0f, A->bottomRadius));
                float slice = ((float)z + 0.5f) / (float)AP_SIZE;
                slice *= slice;
                slice *= (float)AP_SIZE;
                v3 worldPos = camPos;
                float tMax = slice * AP_KM_PER_SLICE; /* aerialPerspectiveSliceToDepth */
                v3 newWorldPos = vadd(worldPos, vscale(worldDir, tMax));
                float viewHeight = vlen(newWorldPos);
                if (viewHeight <= A->bottomRadius + PLANET_RADIUS_OFFSET) {
                    newWorldPos = vscale(vnorm(newWorldPos), A->bottomRadius + PLANET_RADIUS_OFFSET + 0.001f);
                    worldDir = vnorm(vsub(newWorldPos, camPos));
                    tMax = vlen(vsub(newWorldPos, camPos));
                }
                float tMaxMax = tMax;
                float rgba[4] = {0.0f, 0.0f, 0.0f, 1.0f};
                int inside = 1;
                viewHeight = vlen(worldPos);
                if (viewHeight >= A->topRadius) {
                    v3 prev = worldPos;
                    if (!moveToTopAtmosphere(&worldPos, worldDir, A->topRadius)) inside = 0;
                    else {
                        float lengthToAtmosphere = vlen(vsub(prev, worldPos));
                        if (tMaxMax < lengthToAtmosphere) inside = 0;
                        tMaxMax = fmaxf(0.0f, tMaxMax - lengthToAtmosphere);
                    }
                }
                if (inside) {
                    float sampleCountIni = fmaxf(1.0f, ((float)z + 1.0f) * 2.0f);
                    scatter_result ss = integrateScatteredLuminance(A, &trans, &multi, worldPos, worldDir, v3p(sunDirection), 0,
                                                                    sampleCountIni, 0, 1, tMaxMax, v3p(sunIlluminance));
                    float T = (ss.transmittance.x + ss.transmittance.y + ss.transmittance.z) * (1.0f / 3.0f);
                    rgba[0] = ss.L.x; rgba[1] = ss.L.y; rgba[2] = ss.L.z; rgba[3] = 1.0f - T;
                }
                uint16_t* px = &out[4 * (((size_t)z * AP_SIZE + y) * AP_SIZE + x)];
                for (int c = 0; c < 4; c++) px[c] = orc_f32_to_f16(rgba[c]);
            }
    tex_free(&trans);
    tex_free(&multi);
}

static void ap_texel(const uint16_t* vol, int x, int y, int z, float out[4]) {
    x = x < 0 ? 0 : (x > AP_SIZE - 1 ? AP_SIZE - 1 : x);
    y = y < 0 ? 0 : (y > AP_SIZE - 1 ? AP_SIZE - 1 : y);
    z = z < 0 ? 0 : (z > AP_SIZE - 1 ? AP_SIZE - 1 : z);
    const uint16_t* p = &vol[4 * (((size_t)z * AP_SIZE + y) * AP_SIZE + x)];
    for (int c = 0; c < 4; c++) out[c] = orc_f16_to_f32(p[c]);
}
/* weight * trilinear(vol, u, v, sqrt(slice / 32)) for a surface at distance t (km) */
static void ap_lookup(const uint16_t* vol, float u, float v, float t, float out[4]) {
    float slice = t * (1.0f / AP_KM_PER_SLICE); /* aerialPerspectiveDepthToSlice */
    float weight = 1.0f;
    if (slice < 0.5f) {
        weight = clampf(slice * 2.0f, 0.0f, 1.0f);
        slice = 0.5f;
    }
    float w = sqrtf(slice / (float)AP_SIZE);
    float c[3] = {u * (float)AP_SIZE - 0.5f, v * (float)AP_SIZE - 0.5f, w * (float)AP_SIZE - 0.5f};
    int i0[3];
    float f[3];
    for (int a = 0; a < 3; a++) {
        float fl = floorf(c[a]);
        i0[a] = (int)fl;
        f[a] = c[a] - fl;
    }
    float acc[4] = {0, 0, 0, 0};
    for (int dz = 0; dz < 2; dz++)
        for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++) {
                float tx[4];
                ap_texel(vol, i0[0] + dx, i0[1] + dy, i0[2] + dz, tx);
                float wx = dx ? f[0] : 1.0f - f[0], wy = dy ? f[1] : 1.0f - f[1], wz = dz ? f[2] : 1.0f - f[2];
                float wt = (wx * wy) * wz;
                for (int k = 0; k < 4; k++) acc[k] += tx[k] * wt;
            }
    for (int k = 0; k < 4; k++) out[k] = acc[k] * weight;
}
void orc_aerial_perspective_lookup(const uint16_t* vol, float u, float v, float t, float out[4]) { ap_lookup(vol, u, v, t, out); }

/* The native path tracer with the SURVEY 8f-4 extensions (ext = 0: exactly secondaryRays.comp:64-135 per pixel).
 *  ORC_EXT_NEE_SUN    at every hit vertex i < bounces, BEFORE the bounce's two random numbers, two more rotated random
 *                     numbers (u0 with the pixel's x rotation, u1 with its y rotation) pick a direction l inside the sun's
 *                     disc (nee_sun_sample); if n.l > 0 a shadow ray from the bounce origin (pos + n * 1e-6) is tested for
 *                     ANY hit with t >= 0, and if there is none  throughput(after this vertex' albedo) * E * (n.l * weight)
 *                     is added to the pixel, E = sun_centre_radiance at the sky position.  Bounce rays that escape then add
 *                     the sky-view term only (the disc they would otherwise hit by chance is what the shadow rays
 *                     integrate); a PRIMARY ray that escapes still sees the disc.  Shadow rays count as secondary rays.
 *  ORC_EXT_SKY_AT_HIT the sky is evaluated at the origin of the escaping ray (and E at the shaded point) instead of at the
 *                     camera (secondaryRays.comp:37 uses C.cameraPos for every vertex).
 *  ORC_EXT_AERIAL     aerial perspective between the camera and the primary hit from the volume `aerial`
 *                     (orc_gen_aerial_perspective; contract above). */
void orc_render_tris_ext(const orc_scene* s, uint32_t w, uint32_t h, const orc_primary_constants* pc,
                         const orc_secondary_constants* sc, const uint8_t* blueNoise, uint32_t bnW, uint32_t bnH,
                         const orc_atmosphere_params* atmo, const uint16_t* trans, const uint32_t* skyView,
                         uint32_t spp, uint32_t bounces, int use_bvh, uint32_t y0, uint32_t y1, float* accum,
                         uint32_t* visibility, uint64_t* rays_out, uint32_t ext, const uint16_t* aerial) {
    const int nee = (ext & ORC_EXT_NEE_SUN) != 0, at_hit = (ext & ORC_EXT_SKY_AT_HIT) != 0;
    const int ap_on = (ext & ORC_EXT_AERIAL) != 0 && aerial != NULL;
    sky_ctx S = sky_ctx_make(atmo, trans, skyView, sc->cameraPos);
    const v3 E_cam = sun_centre_radiance(&S, S.cameraPos);
    uint64_t prim = 0, sec = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : prim, sec)
    for (int y = (int)y0; y < (int)y1; y++)
        for (uint32_t x = 0; x < w; x++) {
            size_t px = (size_t)y * w + x;
            v3 o, d;
            ray_gen(&pc->invView, &pc->invProjection, x, (uint32_t)y, w, h, &o, &d);
            hit_t h0 = closest_hit(s, o, d, use_bvh);
            prim++;
            if (visibility) visibility[px] = h0.id;
            v3 p0 = V(0, 0, 0), n0 = d;
            if (h0.id != NONE_ID) {
                p0 = vadd(o, vscale(d, h0.t));
                n0 = tri_normal(s, h0.id, d);
            }
            uint32_t seed = (sc->frameCounter << 1u) | 1u;
            const uint8_t* bn = &blueNoise[4 * ((size_t)((uint32_t)y % bnH) * bnW + (x % bnW))];
            float rotx = (float)bn[0] / 255.0f, roty = (float)bn[1] / 255.0f;
            v3 color = V(0, 0, 0);
            float ap[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (ap_on && h0.id != NONE_ID)
                ap_lookup(aerial, ((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h, h0.t, ap);
            for (uint32_t smp = 0; smp < spp; smp++) {
                v3 thr = V(1, 1, 1);
                if (ap_on && h0.id != NONE_ID) {
                    thr = vsplat(1.0f - ap[3]);
                    color = vadd(color, V(ap[0], ap[1], ap[2]));
                }
                uint32_t hid = h0.id;
                v3 hpos = p0, hn = n0;
                v3 sky_pos = S.cameraPos;  /* where the sky is evaluated for the ray that escapes next */
                for (uint32_t i = 0; i < bounces + 1u; i++) {
                    if (i > 0) {
                        v3 ro = vadd(hpos, vscale(hn, 0.000001f));
                        if (at_hit) sky_pos = ro;
                        float r0 = rotated_random(&seed, rotx);
                        float r1 = rotated_random(&seed, roty);
                        v3 rd = vnorm(vadd(hn, random_sphere_point(r0 * 2.0f - 1.0f, r1 * 2.0f - 1.0f)));
                        sec++;
                        hit_t hh = closest_hit(s, ro, rd, use_bvh);
                        hid = hh.id;
                        if (hid != NONE_ID) {
                            hpos = vadd(ro, vscale(rd, hh.t));
                            hn = tri_normal(s, hid, rd);
                        } else {
                            hn = rd;
                        }
                    }
                    if (hid != NONE_ID) {
                        thr = vmul(thr, v3p(&s->albedo[3 * (size_t)hid]));
                        if (nee && i < bounces) {
                            float u0 = rotated_random(&seed, rotx);
                            float u1 = rotated_random(&seed, roty);
                            v3 l;
                            float wgt;
                            nee_sun_sample(u0, u1, &l, &wgt);
                            float ndotl = vdot(hn, l);
                            if (ndotl > 0.0f) {
                                v3 so = vadd(hpos, vscale(hn, 0.000001f));
                                v3 E = at_hit ? sun_centre_radiance(&S, so) : E_cam;
                                if (E.x > 0.0f || E.y > 0.0f || E.z > 0.0f) {
                                    sec++;
                                    hit_t sh = closest_hit(s, so, l, use_bvh); /* any hit <=> a closest hit exists */
                                    if (sh.id == NONE_ID) color = vadd(color, vscale(vmul(thr, E), ndotl * wgt));
                                }
                            }
                        }
                    } else {
                        color = vadd(color, vmul(thr, sky_color_at(&S, sky_pos, hn, (nee && i > 0) ? 1 : 3)));
                        break;
                    }
                }
            }
            accum[4 * px] += color.x; accum[4 * px + 1] += color.y; accum[4 * px + 2] += color.z;
            accum[4 * px + 3] += (float)spp;
        }
    if (rays_out) { rays_out[0] = prim; rays_out[1] = sec; }
    sky_ctx_free(&S);
}

void orc_resolve(uint32_t npixels, const float* accum, float* color32) {
    for (uint32_t i = 0; i < npixels; i++) {
        float n = accum[4 * (size_t)i + 3];
        for (int c = 0; c < 3; c++) color32[4 * (size_t)i + c] = n > 0.0f ? accum[4 * (size_t)i + c] / n : 0.0f;
        color32[4 * (size_t)i + 3] = 1.0f;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
