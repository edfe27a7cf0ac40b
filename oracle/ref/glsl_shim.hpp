// glsl_shim.hpp -- the "Vulkan implementation" under oracle/_ref (TEST INFRASTRUCTURE, not product code).
//
// oracle/_ref compiles the reference's own GLSL compute shaders (read from /root/reference at build time, passed
// through the lexical pre-pass oracle/ref/glsl2cpp.py) as C++20.  This header supplies what a GLSL compiler + a
// Vulkan driver supply: the vector/matrix types with swizzles, the built-in functions, images and samplers over
// host arrays, and the dispatch loop (including work groups with shared memory + barrier()).
//
// Everything the GLSL/Vulkan specifications leave to the implementation is fixed here, with the same rules the
// CPU restatement (oracle/minote_oracle.c) documents, so the two can be compared bit for bit:
//   fp32 -> fp16 image store : IEEE round-to-nearest-even, overflow -> inf
//   fp32 -> B10G11R11 store  : RNE to 6/6/5-bit mantissa, negatives/NaN -> 0, saturate to max finite
//   fp32 -> unorm8 store     : rint(clamp(x,0,1)*255), NaN -> 0
//   linear filtering         : fp32 lerp on texel-centre coordinates (sampler.subtexel_bits = 0) or weights held to
//                              k/2^bits with zero-weight texels skipped (subtexel_bits = 8, the denoiser's sampler)
//   mat*vec                  : columns scaled by the vector's components, summed left to right
//   dot                      : products summed left to right; length = sqrt(dot); normalize = v / length
//   mix(x,y,a)               : x*(1-a) + y*a;  clamp = min(max(x,lo),hi);  fract = x - floor(x)
//   min/max with a NaN       : the other operand (IEEE minNum/maxNum, what GPU min/max instructions do)
//   transcendentals          : glibc libm (sinf, cosf, acosf, powf, expf, sqrtf); round = roundf
//   undefined values         : default-constructed vectors/scalars in structs are zero; out-of-range constant-array
//                              reads are the binder's business (primaryRay.comp:33-34 indexes Spheres[-1u] on a miss)
// Build with -ffp-contract=off: GLSL on the reference's path has no fused multiply-add the source does not spell.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

namespace glsl {

typedef unsigned int uint;

template <class T, int N> struct tvec;
template <class V, int N, int... I> struct swz;

// ---- component counting / access for the GLSL constructor rule ("flatten the arguments, left to right") ----
template <class A> struct ncomp { static constexpr int value = std::is_arithmetic_v<A> ? 1 : -1000; };
template <class T, int N> struct ncomp<tvec<T, N>> { static constexpr int value = N; };
template <class V, int N, int... I> struct ncomp<swz<V, N, I...>> { static constexpr int value = sizeof...(I); };

template <class T, class A> inline void append(T* dst, int& n, const A& a) {
    if constexpr (std::is_arithmetic_v<A>) dst[n++] = static_cast<T>(a);
    else for (int k = 0; k < ncomp<A>::value; k++) dst[n++] = static_cast<T>(a[k]);
}

// ---- swizzle proxy: lives in a union with the vector's components ----
template <class T, int M, int N, int... I>
struct swz<tvec<T, M>, N, I...> {
    T v[N];
    static constexpr int idx(int k) { constexpr int t[] = {I...}; return t[k]; }
    T operator[](int k) const { return v[idx(k)]; }
    operator tvec<T, M>() const { tvec<T, M> r; for (int k = 0; k < M; k++) r[k] = v[idx(k)]; return r; }
    swz& operator=(const tvec<T, M>& o) { for (int k = 0; k < M; k++) v[idx(k)] = o[k]; return *this; }
    swz& operator=(const swz& o) { tvec<T, M> t = o; return *this = t; }
    swz& operator+=(const tvec<T, M>& o) { for (int k = 0; k < M; k++) v[idx(k)] += o[k]; return *this; }
    swz& operator-=(const tvec<T, M>& o) { for (int k = 0; k < M; k++) v[idx(k)] -= o[k]; return *this; }
    swz& operator*=(const tvec<T, M>& o) { for (int k = 0; k < M; k++) v[idx(k)] *= o[k]; return *this; }
    swz& operator/=(const tvec<T, M>& o) { for (int k = 0; k < M; k++) v[idx(k)] /= o[k]; return *this; }
    swz& operator*=(T s) { for (int k = 0; k < M; k++) v[idx(k)] *= s; return *this; }
    swz& operator/=(T s) { for (int k = 0; k < M; k++) v[idx(k)] /= s; return *this; }
};

#define GLSL_SWZ2(V, N, a, b, ia, ib) swz<tvec<T, 2>, N, ia, ib> a##b;
#define GLSL_SWZ3(V, N, a, b, c, ia, ib, ic) swz<tvec<T, 3>, N, ia, ib, ic> a##b##c;

// all 2-component swizzles of the first two names
#define GLSL_SWZ_OF2(N, x, y) \
    GLSL_SWZ2(T, N, x, x, 0, 0) GLSL_SWZ2(T, N, x, y, 0, 1) GLSL_SWZ2(T, N, y, x, 1, 0) GLSL_SWZ2(T, N, y, y, 1, 1)
// the 2- and 3-component swizzles that involve a third name
#define GLSL_SWZ_OF3(N, x, y, z) \
    GLSL_SWZ_OF2(N, x, y) \
    GLSL_SWZ2(T, N, x, z, 0, 2) GLSL_SWZ2(T, N, y, z, 1, 2) GLSL_SWZ2(T, N, z, x, 2, 0) GLSL_SWZ2(T, N, z, y, 2, 1) \
    GLSL_SWZ2(T, N, z, z, 2, 2) \
    GLSL_SWZ3(T, N, x, y, z, 0, 1, 2) GLSL_SWZ3(T, N, x, z, y, 0, 2, 1) GLSL_SWZ3(T, N, y, x, z, 1, 0, 2) \
    GLSL_SWZ3(T, N, y, z, x, 1, 2, 0) GLSL_SWZ3(T, N, z, x, y, 2, 0, 1) GLSL_SWZ3(T, N, z, y, x, 2, 1, 0) \
    GLSL_SWZ3(T, N, x, x, x, 0, 0, 0) GLSL_SWZ3(T, N, y, y, y, 1, 1, 1) GLSL_SWZ3(T, N, z, z, z, 2, 2, 2)
#define GLSL_SWZ_OF4(N, x, y, z, w) \
    GLSL_SWZ_OF3(N, x, y, z) \
    GLSL_SWZ2(T, N, x, w, 0, 3) GLSL_SWZ2(T, N, y, w, 1, 3) GLSL_SWZ2(T, N, z, w, 2, 3) GLSL_SWZ2(T, N, w, w, 3, 3) \
    GLSL_SWZ2(T, N, w, x, 3, 0) GLSL_SWZ2(T, N, w, y, 3, 1) GLSL_SWZ2(T, N, w, z, 3, 2) \
    GLSL_SWZ3(T, N, x, y, w, 0, 1, 3) GLSL_SWZ3(T, N, y, z, w, 1, 2, 3) GLSL_SWZ3(T, N, x, z, w, 0, 2, 3) \
    GLSL_SWZ3(T, N, w, w, w, 3, 3, 3)

// ---- operators shared by the three vector sizes (non-template friends: swizzle proxies convert implicitly) ----
#define GLSL_VEC_COMMON(N) \
    T& operator[](int k) { return c_[k]; } \
    const T& operator[](int k) const { return c_[k]; } \
    tvec() { for (int k = 0; k < N; k++) c_[k] = T(); } \
    tvec(const tvec& o) { for (int k = 0; k < N; k++) c_[k] = o.c_[k]; } \
    tvec& operator=(const tvec& o) { for (int k = 0; k < N; k++) c_[k] = o.c_[k]; return *this; } \
    /* GLSL constructor: one scalar splats, otherwise the arguments' components in order (a longer single vector is cut) */ \
    template <class... A, std::enable_if_t<(sizeof...(A) >= 1) && ((ncomp<A>::value + ...) >= 1), int> = 0> \
    explicit(sizeof...(A) == 1) tvec(const A&... a) { \
        constexpr int total = (ncomp<A>::value + ...); \
        static_assert(total == 1 || total >= N, "not enough components"); \
        static_assert(total == 1 || total == N || sizeof...(A) == 1, "too many components"); \
        T tmp[total > N ? total : N]; int n = 0; \
        (append(tmp, n, a), ...); \
        for (int k = 0; k < N; k++) c_[k] = total == 1 ? tmp[0] : tmp[k]; \
    } \
    /* GLSL's implicit conversion int -> uint */ \
    tvec(const tvec<int, N>& o) requires std::is_same_v<T, uint> { for (int k = 0; k < N; k++) c_[k] = T(o[k]); } \
    template <int NN, int... I> tvec(const swz<tvec<int, N>, NN, I...>& o) requires std::is_same_v<T, uint> \
    { for (int k = 0; k < N; k++) c_[k] = T(o[k]); } \
    friend tvec operator+(const tvec& a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] + b[k]; return r; } \
    friend tvec operator-(const tvec& a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] - b[k]; return r; } \
    friend tvec operator*(const tvec& a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] * b[k]; return r; } \
    friend tvec operator/(const tvec& a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] / b[k]; return r; } \
    friend tvec operator%(const tvec& a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] % b[k]; return r; } \
    friend tvec operator+(const tvec& a, T b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] + b; return r; } \
    friend tvec operator-(const tvec& a, T b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] - b; return r; } \
    friend tvec operator*(const tvec& a, T b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] * b; return r; } \
    friend tvec operator/(const tvec& a, T b) { tvec r; for (int k = 0; k < N; k++) r[k] = a[k] / b; return r; } \
    friend tvec operator+(T a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a + b[k]; return r; } \
    friend tvec operator-(T a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a - b[k]; return r; } \
    friend tvec operator*(T a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a * b[k]; return r; } \
    friend tvec operator/(T a, const tvec& b) { tvec r; for (int k = 0; k < N; k++) r[k] = a / b[k]; return r; } \
    friend tvec operator-(const tvec& a) { tvec r; for (int k = 0; k < N; k++) r[k] = -a[k]; return r; } \
    tvec& operator+=(const tvec& o) { for (int k = 0; k < N; k++) c_[k] += o[k]; return *this; } \
    tvec& operator-=(const tvec& o) { for (int k = 0; k < N; k++) c_[k] -= o[k]; return *this; } \
    tvec& operator*=(const tvec& o) { for (int k = 0; k < N; k++) c_[k] *= o[k]; return *this; } \
    tvec& operator/=(const tvec& o) { for (int k = 0; k < N; k++) c_[k] /= o[k]; return *this; } \
    tvec& operator+=(T s) { for (int k = 0; k < N; k++) c_[k] += s; return *this; } \
    tvec& operator-=(T s) { for (int k = 0; k < N; k++) c_[k] -= s; return *this; } \
    tvec& operator*=(T s) { for (int k = 0; k < N; k++) c_[k] *= s; return *this; } \
    tvec& operator/=(T s) { for (int k = 0; k < N; k++) c_[k] /= s; return *this; }

template <class T> struct tvec<T, 2> {
    union {
        T c_[2];
        struct { T x, y; };
        struct { T r, g; };
        GLSL_SWZ_OF2(2, x, y)
        GLSL_SWZ_OF2(2, r, g)
    };
    GLSL_VEC_COMMON(2)
};
template <class T> struct tvec<T, 3> {
    union {
        T c_[3];
        struct { T x, y, z; };
        struct { T r, g, b; };
        GLSL_SWZ_OF3(3, x, y, z)
        GLSL_SWZ_OF3(3, r, g, b)
    };
    GLSL_VEC_COMMON(3)
};
template <class T> struct tvec<T, 4> {
    union {
        T c_[4];
        struct { T x, y, z, w; };
        struct { T r, g, b, a; };
        GLSL_SWZ_OF4(4, x, y, z, w)
        GLSL_SWZ_OF4(4, r, g, b, a)
    };
    GLSL_VEC_COMMON(4)
};

typedef tvec<float, 2> vec2;  typedef tvec<float, 3> vec3;  typedef tvec<float, 4> vec4;
typedef tvec<uint, 2> uvec2;  typedef tvec<uint, 3> uvec3;  typedef tvec<uint, 4> uvec4;
typedef tvec<int, 2> ivec2;   typedef tvec<int, 3> ivec3;   typedef tvec<int, 4> ivec4;
typedef tvec<bool, 2> bvec2;  typedef tvec<bool, 3> bvec3;  typedef tvec<bool, 4> bvec4;
static_assert(sizeof(vec2) == 8 && sizeof(vec3) == 12 && sizeof(vec4) == 16, "vectors must be packed floats");

// ---- matrices: column-major, constructed from columns ----
struct mat4 {
    vec4 c[4];
    mat4() {}
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int k) { return c[k]; }
    const vec4& operator[](int k) const { return c[k]; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int k) { return c[k]; }
    const vec3& operator[](int k) const { return c[k]; }
};
static_assert(sizeof(mat4) == 64, "mat4 must be 16 packed floats");
inline vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b) { return mat4(a * b.c[0], a * b.c[1], a * b.c[2], a * b.c[3]); }

// ---- built-in functions ----
inline float sqrt(float x) { return ::sqrtf(x); }
inline float exp(float x) { return ::expf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float round(float x) { return ::roundf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline float fract(float x) { return x - ::floorf(x); }
// min/max with a NaN operand are undefined in GLSL (SPIR-V FMin/FMax: "which operand is the result is undefined");
// GPU min/max instructions return the other operand (IEEE minNum/maxNum), so clamp(NaN, 0, 1) = 0 -- the denoiser
// relies on it where both taps are sky (inf - inf).  fminf/fmaxf have exactly that rule.
inline float min(float x, float y) { return ::fminf(x, y); }
inline float max(float x, float y) { return ::fmaxf(x, y); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

#define GLSL_VEC_BUILTINS(V, N) \
    inline float dot(const V& a, const V& b) { float s = a[0] * b[0]; for (int k = 1; k < N; k++) s = s + a[k] * b[k]; return s; } \
    inline float length(const V& a) { return ::sqrtf(dot(a, a)); } \
    inline V normalize(const V& a) { return a / length(a); } \
    inline V abs(const V& a) { V r; for (int k = 0; k < N; k++) r[k] = ::fabsf(a[k]); return r; } \
    inline V exp(const V& a) { V r; for (int k = 0; k < N; k++) r[k] = ::expf(a[k]); return r; } \
    inline V sqrt(const V& a) { V r; for (int k = 0; k < N; k++) r[k] = ::sqrtf(a[k]); return r; } \
    inline V pow(const V& a, const V& b) { V r; for (int k = 0; k < N; k++) r[k] = ::powf(a[k], b[k]); return r; } \
    inline V min(const V& a, const V& b) { V r; for (int k = 0; k < N; k++) r[k] = min(a[k], b[k]); return r; } \
    inline V max(const V& a, const V& b) { V r; for (int k = 0; k < N; k++) r[k] = max(a[k], b[k]); return r; } \
    inline V clamp(const V& a, float lo, float hi) { V r; for (int k = 0; k < N; k++) r[k] = clamp(a[k], lo, hi); return r; } \
    inline V mix(const V& x, const V& y, const V& a) { V r; for (int k = 0; k < N; k++) r[k] = mix(x[k], y[k], a[k]); return r; } \
    inline V step(float e, const V& x) { V r; for (int k = 0; k < N; k++) r[k] = step(e, x[k]); return r; } \
    inline V smoothstep(float e0, float e1, const V& x) { V r; for (int k = 0; k < N; k++) r[k] = smoothstep(e0, e1, x[k]); return r; }
GLSL_VEC_BUILTINS(vec2, 2)
GLSL_VEC_BUILTINS(vec3, 3)
GLSL_VEC_BUILTINS(vec4, 4)

inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3{dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])}; }

#define GLSL_RELATIONAL(V, B, N) \
    inline B greaterThan(const V& a, const V& b) { B r; for (int k = 0; k < N; k++) r[k] = a[k] > b[k]; return r; } \
    inline B greaterThanEqual(const V& a, const V& b) { B r; for (int k = 0; k < N; k++) r[k] = a[k] >= b[k]; return r; }
GLSL_RELATIONAL(uvec2, bvec2, 2)
GLSL_RELATIONAL(uvec3, bvec3, 3)
GLSL_RELATIONAL(vec2, bvec2, 2)
GLSL_RELATIONAL(vec3, bvec3, 3)
inline bool any(const bvec2& b) { return b.x || b.y; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

// ---- storage formats ----
inline uint16_t f32_to_f16(float f) {
    uint32_t u; std::memcpy(&u, &f, 4);
    uint32_t sign = (u >> 16) & 0x8000u, mag = u & 0x7fffffffu;
    if (mag > 0x7f800000u) return uint16_t(sign | 0x7e00u);                    // NaN
    if (mag >= 0x47800000u) return uint16_t(sign | 0x7c00u);                   // >= 65536 (incl. inf) -> inf
    if (mag >= 0x38800000u) {                                                  // normal half
        uint32_t h = (mag - 0x38000000u) >> 13, rem = mag & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;                // RNE; a carry into 0x7c00 is the inf
        return uint16_t(sign | h);
    }
    if (mag < 0x33000000u) return uint16_t(sign);                              // < 2^-25 -> 0
    int e = int(mag >> 23);                                                    // subnormal half: value = m * 2^-24
    uint32_t m = (mag & 0x7fffffu) | 0x800000u;
    int shift = 126 - e;                                                       // 14 .. 24
    uint32_t h = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) h++;
    return uint16_t(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    uint32_t sign = uint32_t(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u, u;
    if (e == 31u) u = sign | 0x7f800000u | (m << 13);
    else if (e) u = sign | ((e + 112u) << 23) | (m << 13);
    else if (m) { int s = 0; while (!(m & 1024u)) { m <<= 1; s++; } u = sign | uint32_t(113 - s) << 23 | ((m & 1023u) << 13); }
    else u = sign;
    float f; std::memcpy(&f, &u, 4); return f;
}
// unsigned small float with MB mantissa bits, 5 exponent bits (bias 15), no sign
inline uint32_t f32_to_ufloat(float f, int MB) {
    if (!(f > 0.0f)) return 0;                                                 // negatives, -0, NaN -> 0
    const uint32_t maxv = (30u << MB) | ((1u << MB) - 1u);
    uint32_t u; std::memcpy(&u, &f, 4);
    int e = int(u >> 23) - 127 + 15;
    uint32_t m = u & 0x7fffffu;
    if ((u >> 23) == 255u) return maxv;                                        // inf -> max finite
    int drop = 23 - MB;
    if (e <= 0) {                                                              // subnormal target
        if (e < -MB) return 0;
        m |= 0x800000u; drop += 1 - e; e = 0;
        if (drop > 31) return 0;
    }
    uint32_t q = m >> drop, rem = m & ((1u << drop) - 1u), half = 1u << (drop - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t r = (uint32_t(e) << MB) + q;                                      // mantissa carry bumps the exponent
    return r > maxv ? maxv : r;
}
inline float ufloat_to_f32(uint32_t v, int MB) {
    uint32_t e = v >> MB, m = v & ((1u << MB) - 1u);
    if (e == 0) return std::ldexp(float(m), -14 - MB);
    if (e == 31) return m ? NAN : INFINITY;
    return std::ldexp(float(m | (1u << MB)), int(e) - 15 - MB);
}
inline uint8_t f32_to_unorm8(float f) {
    if (!(f == f)) return 0;
    float c = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return uint8_t(::rintf(c * 255.0f));
}

// ---- images and samplers over host arrays ----
enum Format { R32UI, R16F, RG16F, RGBA16F, B10G11R11, RGBA8, RGBA32F };
struct Image { void* data = nullptr; int w = 0, h = 0; Format fmt = RGBA32F; };
typedef Image image2D;
struct UImage : Image {};
typedef UImage uimage2D;
struct Sampler {
    const void* data = nullptr; int w = 0, h = 0; Format fmt = RGBA32F;
    bool linear = false, repeat = false; int subtexel_bits = 0;
};
typedef Sampler sampler2D;
struct USampler : Sampler {};
typedef USampler usampler2D;

inline ivec2 imageSize(const Image& i) { return ivec2{i.w, i.h}; }
inline ivec2 textureSize(const Sampler& s, int) { return ivec2{s.w, s.h}; }

inline void imageStore(const Image& img, const ivec2& p, const vec4& v) {
    if (p.x < 0 || p.y < 0 || p.x >= img.w || p.y >= img.h) return;
    size_t i = size_t(p.y) * img.w + p.x;
    switch (img.fmt) {
    case R16F: ((uint16_t*)img.data)[i] = f32_to_f16(v.x); break;
    case RG16F: for (int k = 0; k < 2; k++) ((uint16_t*)img.data)[2 * i + k] = f32_to_f16(v[k]); break;
    case RGBA16F: for (int k = 0; k < 4; k++) ((uint16_t*)img.data)[4 * i + k] = f32_to_f16(v[k]); break;
    case RGBA32F: for (int k = 0; k < 4; k++) ((float*)img.data)[4 * i + k] = v[k]; break;
    case RGBA8: for (int k = 0; k < 4; k++) ((uint8_t*)img.data)[4 * i + k] = f32_to_unorm8(v[k]); break;
    case B10G11R11:
        ((uint32_t*)img.data)[i] = f32_to_ufloat(v.x, 6) | (f32_to_ufloat(v.y, 6) << 11) | (f32_to_ufloat(v.z, 5) << 22);
        break;
    default: break;
    }
}
inline void imageStore(const UImage& img, const ivec2& p, const uvec4& v) {
    if (p.x < 0 || p.y < 0 || p.x >= img.w || p.y >= img.h) return;
    ((uint32_t*)img.data)[size_t(p.y) * img.w + p.x] = v.x;
}

inline vec4 fetch_texel(const Sampler& s, int x, int y) {
    size_t i = size_t(y) * s.w + x;
    switch (s.fmt) {
    case R16F: return vec4{f16_to_f32(((const uint16_t*)s.data)[i]), 0.0f, 0.0f, 1.0f};
    case RG16F: return vec4{f16_to_f32(((const uint16_t*)s.data)[2 * i]), f16_to_f32(((const uint16_t*)s.data)[2 * i + 1]), 0.0f, 1.0f};
    case RGBA16F: { const uint16_t* p = (const uint16_t*)s.data + 4 * i;
        return vec4{f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2]), f16_to_f32(p[3])}; }
    case RGBA32F: { const float* p = (const float*)s.data + 4 * i; return vec4{p[0], p[1], p[2], p[3]}; }
    case RGBA8: { const uint8_t* p = (const uint8_t*)s.data + 4 * i;
        return vec4{float(p[0]) / 255.0f, float(p[1]) / 255.0f, float(p[2]) / 255.0f, float(p[3]) / 255.0f}; }
    case B10G11R11: { uint32_t p = ((const uint32_t*)s.data)[i];
        return vec4{ufloat_to_f32(p & 0x7ffu, 6), ufloat_to_f32((p >> 11) & 0x7ffu, 6), ufloat_to_f32(p >> 22, 5), 1.0f}; }
    default: return vec4{};
    }
}
inline vec4 texelFetch(const Sampler& s, const ivec2& p, int) {
    if (p.x < 0 || p.y < 0 || p.x >= s.w || p.y >= s.h) return vec4{};
    return fetch_texel(s, p.x, p.y);
}
inline uvec4 texelFetch(const USampler& s, const ivec2& p, int) {
    if (p.x < 0 || p.y < 0 || p.x >= s.w || p.y >= s.h) return uvec4{};
    return uvec4{((const uint32_t*)s.data)[size_t(p.y) * s.w + p.x], 0u, 0u, 1u};
}
inline int wrap_coord(int i, int n, bool repeat) {
    if (repeat) { i %= n; return i < 0 ? i + n : i; }
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}
inline vec4 textureLod(const Sampler& s, const vec2& uv, float) {
    if (!s.linear) {
        int x = wrap_coord(int(::floorf(uv.x * float(s.w))), s.w, s.repeat);
        int y = wrap_coord(int(::floorf(uv.y * float(s.h))), s.h, s.repeat);
        return fetch_texel(s, x, y);
    }
    float x = uv.x * float(s.w) - 0.5f, y = uv.y * float(s.h) - 0.5f;
    float fx0 = ::floorf(x), fy0 = ::floorf(y);
    float fx = x - fx0, fy = y - fy0;
    if (s.subtexel_bits) {
        float q = float(1 << s.subtexel_bits);
        fx = ::rintf(fx * q) * (1.0f / q); fy = ::rintf(fy * q) * (1.0f / q);
    }
    int x0 = wrap_coord(int(fx0), s.w, s.repeat), x1 = wrap_coord(int(fx0) + 1, s.w, s.repeat);
    int y0 = wrap_coord(int(fy0), s.h, s.repeat), y1 = wrap_coord(int(fy0) + 1, s.h, s.repeat);
    float gx = 1.0f - fx, gy = 1.0f - fy;
    if (s.subtexel_bits) {   // zero-weight texels are not read (an inf texel must not turn into NaN through 0 * inf)
        vec4 top = fx == 0.0f ? fetch_texel(s, x0, y0) : (gx == 0.0f ? fetch_texel(s, x1, y0)
                                : fetch_texel(s, x0, y0) * gx + fetch_texel(s, x1, y0) * fx);
        if (fy == 0.0f) return top;
        vec4 bot = fx == 0.0f ? fetch_texel(s, x0, y1) : (gx == 0.0f ? fetch_texel(s, x1, y1)
                                : fetch_texel(s, x0, y1) * gx + fetch_texel(s, x1, y1) * fx);
        return gy == 0.0f ? bot : top * gy + bot * fy;
    }
    vec4 top = fetch_texel(s, x0, y0) * gx + fetch_texel(s, x1, y0) * fx;
    vec4 bot = fetch_texel(s, x0, y1) * gx + fetch_texel(s, x1, y1) * fx;
    return top * gy + bot * fy;
}
inline vec4 texture(const Sampler& s, const vec2& uv) { return textureLod(s, uv, 0.0f); }

// ---- invocation state and dispatch ----
inline thread_local uvec3 gl_GlobalInvocationID;
inline thread_local std::barrier<>* current_group_barrier = nullptr;
inline void barrier() { if (current_group_barrier) current_group_barrier->arrive_and_wait(); }
#define GLSL_LOCAL_SIZE(x, y, z) static const uint gl_WorkGroupSizeX = x, gl_WorkGroupSizeY = y, gl_WorkGroupSizeZ = z;

// w x h invocations of a shader without barriers (the work-group shape is then immaterial)
template <class F> inline void dispatch_invocations(uint w, uint h, F shader_main) {
#pragma omp parallel for schedule(dynamic, 4)
    for (long long y = 0; y < (long long)h; y++)
        for (uint x = 0; x < w; x++) {
            gl_GlobalInvocationID = uvec3{x, uint(y), 0u};
            shader_main();
        }
}
// gx x gy work groups of (1, 1, lz) invocations that share memory and meet at barrier(): lz host threads, one group
// at a time (the shader's `shared` arrays are plain statics)
template <class F> inline void dispatch_groups_z(uint gx, uint gy, uint lz, F shader_main) {
    std::barrier<> bar((std::ptrdiff_t)lz);
    std::vector<std::thread> pool;
    for (uint z = 0; z < lz; z++)
        pool.emplace_back([&, z] {
            current_group_barrier = &bar;
            for (uint y = 0; y < gy; y++)
                for (uint x = 0; x < gx; x++) {
                    gl_GlobalInvocationID = uvec3{x, y, z};
                    shader_main();
                    bar.arrive_and_wait();   // group boundary: nobody overwrites shared memory early
                }
            current_group_barrier = nullptr;
        });
    for (auto& t : pool) t.join();
}

}  // namespace glsl
