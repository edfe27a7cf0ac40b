// minote.math -- the subset of the reference's math module the hot path depends on
// (src/stx/math.ixx: vec :53-240, mat :519-675, operator* :680-706, inverse :761-815,
// look :823-844, perspective :849-860, _deg/_m/_km literals :872-886).
// Written from scratch; only the fp32 OPERATION ORDER is kept, because the matrices that reach the
// GPU must be bit-identical to what the reference would upload (SURVEY.md §8 row a1).
module;
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
export module minote.math;

export using u32 = std::uint32_t;

export template <std::size_t N, typename T>
struct vec {
    std::array<T, N> v{};
    constexpr vec() = default;
    template <typename... A>
        requires(sizeof...(A) == N)
    constexpr vec(A... a) : v{static_cast<T>(a)...} {}
    constexpr auto operator[](std::size_t i) -> T& { return v[i]; }
    constexpr auto operator[](std::size_t i) const -> T { return v[i]; }
    constexpr auto x() const -> T { return v[0]; }
    constexpr auto y() const -> T { return v[1]; }
    constexpr auto z() const -> T requires(N > 2) { return v[2]; }
    constexpr auto w() const -> T requires(N > 3) { return v[3]; }
    constexpr auto x() -> T& { return v[0]; }
    constexpr auto y() -> T& { return v[1]; }
    constexpr auto z() -> T& requires(N > 2) { return v[2]; }
    constexpr auto w() -> T& requires(N > 3) { return v[3]; }
};

export using vec2 = vec<2, float>;
export using vec3 = vec<3, float>;
export using vec4 = vec<4, float>;
export using uvec2 = vec<2, u32>;

export template <std::size_t N, typename T>
constexpr auto operator+(vec<N, T> a, vec<N, T> const& b) -> vec<N, T> {
    for (std::size_t i = 0; i < N; i++) a[i] = a[i] + b[i];
    return a;
}
export template <std::size_t N, typename T>
constexpr auto operator-(vec<N, T> a, vec<N, T> const& b) -> vec<N, T> {
    for (std::size_t i = 0; i < N; i++) a[i] = a[i] - b[i];
    return a;
}
export template <std::size_t N, typename T>
constexpr auto operator*(vec<N, T> a, vec<N, T> const& b) -> vec<N, T> {
    for (std::size_t i = 0; i < N; i++) a[i] = a[i] * b[i];
    return a;
}
export template <std::size_t N, typename T>
constexpr auto operator*(vec<N, T> a, T s) -> vec<N, T> {
    for (std::size_t i = 0; i < N; i++) a[i] = a[i] * s;
    return a;
}
export template <std::size_t N, typename T>
constexpr auto operator/(vec<N, T> a, T s) -> vec<N, T> {
    for (std::size_t i = 0; i < N; i++) a[i] = a[i] / s;
    return a;
}
export template <std::size_t N, typename T>
constexpr auto operator+=(vec<N, T>& a, vec<N, T> const& b) -> vec<N, T>& { return a = a + b; }

// accumulation starts from T(0), element by element (math.ixx:304-309)
export template <std::size_t N, typename T>
constexpr auto dot(vec<N, T> const& a, vec<N, T> const& b) -> T {
    T r = T(0);
    for (std::size_t i = 0; i < N; i++) r += a[i] * b[i];
    return r;
}
export constexpr auto cross(vec3 const& a, vec3 const& b) -> vec3 {
    return vec3{a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};
}
export inline auto length(vec3 const& a) -> float { return std::sqrt(dot(a, a)); }
export inline auto normalize(vec3 const& a) -> vec3 { return a / length(a); }
export constexpr auto clamp(float v, float lo, float hi) -> float { return std::max(lo, std::min(v, hi)); }

// degrees -> radians as the reference's _deg literal does it: radians<double, Prec = float>(d) = float(d) * Tau_v<float> / 360.0f
// (math.ixx:27,884-886; fp32 arithmetic -- checked against the reference's own code through oracle/_ref)
export constexpr auto deg(double d) -> float { return static_cast<float>(d) * (3.14159265358979323846f * 2.0f) / 360.0f; }
export constexpr float operator""_deg(long double d) { return deg(static_cast<double>(d)); }
export constexpr float operator""_deg(unsigned long long d) { return deg(static_cast<double>(d)); }
export constexpr float operator""_m(long double d) { return static_cast<float>(static_cast<double>(d) * 0.001); }
export constexpr float operator""_km(long double d) { return static_cast<float>(d); }

// column-major 4x4, c[col][row]; byte-compatible with mrt_mat4 and GLSL mat4
export struct mat4 {
    std::array<vec4, 4> c{};
    constexpr auto operator[](std::size_t i) -> vec4& { return c[i]; }
    constexpr auto operator[](std::size_t i) const -> vec4 const& { return c[i]; }
    static constexpr auto identity() -> mat4 {
        mat4 m;
        for (std::size_t i = 0; i < 4; i++) m[i][i] = 1.0f;
        return m;
    }
};
static_assert(sizeof(mat4) == 64);

// result column j = sum_k left[k] * right[j][k], left to right (math.ixx:680-696)
export constexpr auto operator*(mat4 const& l, mat4 const& r) -> mat4 {
    mat4 out;
    for (std::size_t j = 0; j < 4; j++) out[j] = l[0] * r[j][0] + l[1] * r[j][1] + l[2] * r[j][2] + l[3] * r[j][3];
    return out;
}
// row i dotted with v (math.ixx:699-706)
export constexpr auto operator*(mat4 const& m, vec4 const& v) -> vec4 {
    vec4 out;
    for (std::size_t i = 0; i < 4; i++) out[i] = dot(vec4{m[0][i], m[1][i], m[2][i], m[3][i]}, v);
    return out;
}

// cofactor inverse; the 18 2x2 sub-determinants, the four cofactor columns and the determinant are
// formed in the reference's order (math.ixx:761-815) so the result is bit-identical
export constexpr auto inverse(mat4 const& m) -> mat4 {
    auto sub = [&](int a, int b, int c, int d, int e, int f, int g, int h) {
        return m[a][b] * m[c][d] - m[e][f] * m[g][h];
    };
    float const s00 = sub(2, 2, 3, 3, 3, 2, 2, 3), s02 = sub(1, 2, 3, 3, 3, 2, 1, 3), s03 = sub(1, 2, 2, 3, 2, 2, 1, 3);
    float const s04 = sub(2, 1, 3, 3, 3, 1, 2, 3), s06 = sub(1, 1, 3, 3, 3, 1, 1, 3), s07 = sub(1, 1, 2, 3, 2, 1, 1, 3);
    float const s08 = sub(2, 1, 3, 2, 3, 1, 2, 2), s10 = sub(1, 1, 3, 2, 3, 1, 1, 2), s11 = sub(1, 1, 2, 2, 2, 1, 1, 2);
    float const s12 = sub(2, 0, 3, 3, 3, 0, 2, 3), s14 = sub(1, 0, 3, 3, 3, 0, 1, 3), s15 = sub(1, 0, 2, 3, 2, 0, 1, 3);
    float const s16 = sub(2, 0, 3, 2, 3, 0, 2, 2), s18 = sub(1, 0, 3, 2, 3, 0, 1, 2), s19 = sub(1, 0, 2, 2, 2, 0, 1, 2);
    float const s20 = sub(2, 0, 3, 1, 3, 0, 2, 1), s22 = sub(1, 0, 3, 1, 3, 0, 1, 1), s23 = sub(1, 0, 2, 1, 2, 0, 1, 1);
    vec4 const f0{s00, s00, s02, s03}, f1{s04, s04, s06, s07}, f2{s08, s08, s10, s11};
    vec4 const f3{s12, s12, s14, s15}, f4{s16, s16, s18, s19}, f5{s20, s20, s22, s23};
    vec4 const v0{m[1][0], m[0][0], m[0][0], m[0][0]}, v1{m[1][1], m[0][1], m[0][1], m[0][1]};
    vec4 const v2{m[1][2], m[0][2], m[0][2], m[0][2]}, v3{m[1][3], m[0][3], m[0][3], m[0][3]};
    vec4 const pos{1.0f, -1.0f, 1.0f, -1.0f}, neg{-1.0f, 1.0f, -1.0f, 1.0f};
    mat4 adj;
    adj[0] = (v1 * f0 - v2 * f1 + v3 * f2) * pos;
    adj[1] = (v0 * f0 - v2 * f3 + v3 * f4) * neg;
    adj[2] = (v0 * f1 - v1 * f3 + v3 * f5) * pos;
    adj[3] = (v0 * f2 - v1 * f4 + v2 * f5) * neg;
    vec4 const d = m[0] * vec4{adj[0][0], adj[1][0], adj[2][0], adj[3][0]};
    float const ood = 1.0f / ((d[0] + d[1]) + (d[2] + d[3]));
    for (std::size_t j = 0; j < 4; j++) adj[j] = adj[j] * ood;
    return adj;
}

// view matrix from position + unit direction + unit up (math.ixx:823-844)
export inline auto look(vec3 pos, vec3 dir, vec3 up) -> mat4 {
    vec3 const s = normalize(cross(up, dir));
    vec3 const u = cross(dir, s);
    mat4 r = mat4::identity();
    for (std::size_t i = 0; i < 3; i++) {
        r[i][0] = -s[i];
        r[i][1] = u[i];
        r[i][2] = dir[i];
    }
    r[3][0] = dot(s, pos);
    r[3][1] = -dot(u, pos);
    r[3][2] = -dot(dir, pos);
    return r;
}

// inverted infinite-Z perspective: depth 1 at zNear, 0 at infinity (math.ixx:849-860)
export inline auto perspective(float vFov, float aspectRatio, float zNear) -> mat4 {
    float const h = 1.0f / std::tan(0.5f * vFov);
    mat4 r;
    r[0][0] = h * aspectRatio;
    r[1][1] = h;
    r[2][3] = 1.0f;
    r[3][2] = zNear;
    return r;
}
