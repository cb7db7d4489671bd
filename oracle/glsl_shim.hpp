// TEST INFRASTRUCTURE — part of the parity oracle, never linked into the product.
//
// GLSL 4.60 vector types and built-ins on the CPU, so that the reference's generated
// shader text (variation_table.cpp:217-265 output, built from variations.yaml) compiles
// as C++ with its original swizzles. Semantics follow the GLSL spec and the helper
// functions of /root/reference/shaders/include/math.glsl:1-22. All arithmetic is
// binary32; build with -ffp-contract=off so that only the explicit fma() calls fuse.
#pragma once
#include <cmath>
#include <cstdint>

namespace glsl {

typedef unsigned int uint;
struct vec2;

template <int A, int B>
struct swz2 {  // two-component swizzle proxy living inside a union with the components
    float d[4];
    operator vec2() const;
    swz2& operator=(const vec2& v);
    swz2& operator+=(const vec2& v);
    swz2& operator-=(const vec2& v);
    swz2& operator*=(float s);
};

struct vec2 {
    union {
        struct { float x, y; };
        swz2<0, 1> xy;
        swz2<1, 0> yx;
    };
    vec2() : x(0), y(0) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float a) : x(a), y(a) {}
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
    vec2& operator+=(const vec2& o) { x += o.x; y += o.y; return *this; }
    vec2& operator-=(const vec2& o) { x -= o.x; y -= o.y; return *this; }
    vec2& operator*=(const vec2& o) { x *= o.x; y *= o.y; return *this; }
    vec2& operator*=(float s) { x *= s; y *= s; return *this; }
    vec2& operator/=(float s) { x /= s; y /= s; return *this; }
};

template <int A, int B> swz2<A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int A, int B> swz2<A, B>& swz2<A, B>::operator=(const vec2& v) { float a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
template <int A, int B> swz2<A, B>& swz2<A, B>::operator+=(const vec2& v) { float a = v.x, b = v.y; d[A] += a; d[B] += b; return *this; }
template <int A, int B> swz2<A, B>& swz2<A, B>::operator-=(const vec2& v) { float a = v.x, b = v.y; d[A] -= a; d[B] -= b; return *this; }
template <int A, int B> swz2<A, B>& swz2<A, B>::operator*=(float s) { d[A] *= s; d[B] *= s; return *this; }

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
};

struct vec3 {
    union {
        struct { float x, y, z; };
        swz2<0, 1> xy;
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec2& a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
    vec3 xyz() const { return vec3(x, y, z); }
};
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

#define GLSL_BINOP(op)                                                                       \
    inline vec2 operator op(const vec2& a, const vec2& b) { return vec2(a.x op b.x, a.y op b.y); } \
    inline vec2 operator op(const vec2& a, float s) { return vec2(a.x op s, a.y op s); }           \
    inline vec2 operator op(float s, const vec2& a) { return vec2(s op a.x, s op a.y); }
GLSL_BINOP(+)
GLSL_BINOP(-)
GLSL_BINOP(*)
GLSL_BINOP(/)
#undef GLSL_BINOP
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }

// math.glsl:1-4
static const float PI = 3.141592653589793f;
static const float PI_2 = PI / 2.0f;
static const float EPS = (1e-10f);

inline float sin(float v) { return ::sinf(v); }
inline float cos(float v) { return ::cosf(v); }
inline float tan(float v) { return ::tanf(v); }
inline float sinh(float v) { return ::sinhf(v); }
inline float cosh(float v) { return ::coshf(v); }
inline float exp(float v) { return ::expf(v); }
inline float log(float v) { return ::logf(v); }
inline float sqrt(float v) { return ::sqrtf(v); }
inline float pow(float a, float b) { return ::powf(a, b); }
inline float atan(float a, float b) { return ::atan2f(a, b); }
inline float atan(float a) { return ::atanf(a); }
inline float acos(float v) { return ::acosf(v); }
inline float asin(float v) { return ::asinf(v); }
inline float floor(float v) { return ::floorf(v); }
inline float ceil(float v) { return ::ceilf(v); }
inline float trunc(float v) { return ::truncf(v); }
// round(): "the fraction 0.5 will round in a direction chosen by the implementation" — nearest-even here
inline float round(float v) { return ::rintf(v); }
inline float abs(float v) { return ::fabsf(v); }
inline int abs(int v) { return v < 0 ? -v : v; }
inline float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float clamp(float v, float lo, float hi) { return ::fminf(::fmaxf(v, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // GLSL 4.60 §8.3
inline float sign(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
inline float fract(float v) { return v - ::floorf(v); }
inline float mod(float a, float b) { return a - b * ::floorf(a / b); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float length(const vec2& v) { return ::sqrtf(v.x * v.x + v.y * v.y); }
inline vec2 sin(const vec2& v) { return vec2(::sinf(v.x), ::sinf(v.y)); }
inline vec2 cos(const vec2& v) { return vec2(::cosf(v.x), ::cosf(v.y)); }
inline vec2 abs(const vec2& v) { return vec2(::fabsf(v.x), ::fabsf(v.y)); }
inline vec2 mix(const vec2& a, const vec2& b, float t) { return vec2(mix(a.x, b.x, t), mix(a.y, b.y, t)); }

// math.glsl:6-22
inline vec2 sincos(float v) { return vec2(sin(v), cos(v)); }
inline vec2 sinhcosh(float v) { return vec2(sinh(v), cosh(v)); }
inline float mod2(float x, float y) { return x - y * trunc(x / y); }
inline float log10(float x) { return log(x) * 0.434294481903251827651128918916f; }
inline bool badval(float x) { return (x != x) || (x > 1e10f) || (x < -1e10f); }

}  // namespace glsl
