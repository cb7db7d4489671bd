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

struct uvec2 {
    uint x, y;
    uvec2() : x(0), y(0) {}
    uvec2(uint a, uint b) : x(a), y(b) {}
};

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const vec2& v) : x(int(v.x)), y(int(v.y)) {}
    explicit ivec2(const uvec2& v) : x(int(v.x)), y(int(v.y)) {}
};

struct vec3;
template <int A, int B, int C>
struct swz3 {  // three-component swizzle proxy (.xyz / .rgb of a vec4)
    float d[4];
    operator vec3() const;
    swz3& operator=(const vec3& v);
};

struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<0, 1> xy;
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
template <int A, int B, int C> swz3<A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int A, int B, int C> swz3<A, B, C>& swz3<A, B, C>::operator=(const vec3& v) { float a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<0, 1> xy;
        swz3<0, 1, 2> xyz;
        swz3<0, 1, 2> rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec2& a_, float c_, float d_) : x(a_.x), y(a_.y), z(c_), w(d_) {}
    vec4(const vec3& a_, float d_) : x(a_.x), y(a_.y), z(a_.z), w(d_) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    vec4& operator+=(const vec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    vec4& operator*=(float s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};
static_assert(sizeof(vec4) == 16, "vec4 must match the std430 layout of the reference's buffers");
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }

struct uvec3 {
    struct xy_t { uint d[4]; operator uvec2() const { return uvec2(d[0], d[1]); } };
    union {
        struct { uint x, y, z; };
        xy_t xy;
    };
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
struct uvec4 { uint x, y, z, w; };
struct mat4 { float m[16]; };  // column-major, as glUniformMatrix4fv(transpose = false) stores it
inline vec4 operator*(const mat4& M, const vec4& v) {
    return vec4(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * v.w, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * v.w,
                M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z + M.m[14] * v.w, M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * v.w);
}

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

#ifndef GLSL_SHIM_NO_MATH_GLSL  // the shader-text build (oracle/softgl) compiles math.glsl itself
// math.glsl:1-4
static const float PI = 3.141592653589793f;
static const float PI_2 = PI / 2.0f;
static const float EPS = (1e-10f);
#endif

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

inline vec2 floor(const vec2& v) { return vec2(::floorf(v.x), ::floorf(v.y)); }
inline vec3 pow(const vec3& a, const vec3& b) { return vec3(::powf(a.x, b.x), ::powf(a.y, b.y), ::powf(a.z, b.z)); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline uint min(int a, uint b) { return uint(a) < b ? uint(a) : b; }  // GLSL converts the int operand to uint
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline float uintBitsToFloat(uint u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
inline uint atomicAdd(uint& mem, int v) { return __atomic_fetch_add(&mem, uint(v), __ATOMIC_RELAXED); }

#ifndef GLSL_SHIM_NO_MATH_GLSL
// math.glsl:6-22
inline vec2 sincos(float v) { return vec2(sin(v), cos(v)); }
inline vec2 sinhcosh(float v) { return vec2(sinh(v), cosh(v)); }
inline float mod2(float x, float y) { return x - y * trunc(x / y); }
inline float log10(float x) { return log(x) * 0.434294481903251827651128918916f; }
inline bool badval(float x) { return (x != x) || (x > 1e10f) || (x < -1e10f); }
#endif

}  // namespace glsl
