// Device prelude of every generated chaos-game module (compiled by NVRTC for sm_100a).
// It gives the variation snippets of variations.yaml — which are GLSL — the types and
// built-ins they use (vec2/vec3/vec4, sincos, 2-argument atan, mix, mod2, ...), with
// the semantics of shaders/include/math.glsl:1-22, and the JSF32 generator of
// shaders/include/random.glsl:29-41.
namespace rfk_glsl {

typedef unsigned int uint;
typedef uint4 rfk_rng;  // (a, b, c, d) = local_random_state.xyzw (random.glsl:19-23)

// flame.glsl:27 `shared float fp[1024]`: this CTA's temporal sample of the parameter buffer.
// Namespace scope, so that every fp[k] of the generated code is one LDS with an immediate offset.
__shared__ float fp[RFK_TOTAL_PARAMS + 1];
// The only slots that differ between temporal samples are the four rotated affine coefficients (a, b, c, d) of every
// xform (animate.tpl.glsl:42-49). They are kept a second time as one float4 per xform — [0] the final xform, [1 + i] xform i —
// so the kernel fetches the picked xform's coefficients with ONE 128-bit read whose address comes straight from the pick,
// issued before the switch; the generated text names them RFK_AFF(xform, component) (compile_flame_cuda).
__shared__ float4 rfk_aff[RFK_NUM_XFORMS + 1];
#define RFK_AFF(xform, component) rfk_A.component
// keeps the optimiser from reasoning across two tests of the same register (the weight-ordered if-chain of dispatch_a)
#define RFK_OPAQUE(v) asm volatile("" : "+r"(v))

// Packed FP32 (sm_100: FFMA2 / FMUL2 / FADD2, `fma.rn.f32x2`): one instruction does the x and the y lane of a vec2
// operation — the same FP32 rate as two scalar instructions but ONE issue slot, and the kernels are issue bound
// (tools/ffma2_probe.cu: FFMA2 mixed with integer work runs 24 % faster than the same flops as FFMA). A scalar operand
// is broadcast by the instruction itself (`R.F32`), so `s * v` costs no packing. Per-lane results are those of the
// scalar IEEE operations; a product that feeds a sum is fused (vprod below), as the compiler fuses a * b + c.
#define RFK_F2(v) make_float2((v).x, (v).y)
struct vprod;
struct vec2 {
    float x, y;
    vec2() = default;
    __device__ __forceinline__ vec2(float a, float b) : x(a), y(b) {}
    __device__ __forceinline__ explicit vec2(float a) : x(a), y(a) {}
    __device__ __forceinline__ explicit vec2(float2 f) : x(f.x), y(f.y) {}
    __device__ __forceinline__ vec2 yx() const { return vec2(y, x); }
    __device__ __forceinline__ vec2& operator+=(vec2 o) { return *this = vec2(__fadd2_rn(RFK_F2(*this), RFK_F2(o))); }
    __device__ __forceinline__ vec2& operator-=(vec2 o) { x -= o.x; y -= o.y; return *this; }
    __device__ __forceinline__ vec2& operator*=(vec2 o) { return *this = vec2(__fmul2_rn(RFK_F2(*this), RFK_F2(o))); }
    __device__ __forceinline__ vec2& operator/=(vec2 o) { x /= o.x; y /= o.y; return *this; }
    __device__ __forceinline__ vec2& operator+=(float s) { return *this = vec2(__fadd2_rn(RFK_F2(*this), make_float2(s, s))); }
    __device__ __forceinline__ vec2& operator-=(float s) { x -= s; y -= s; return *this; }
    __device__ __forceinline__ vec2& operator*=(float s) { return *this = vec2(__fmul2_rn(RFK_F2(*this), make_float2(s, s))); }
    __device__ __forceinline__ vec2& operator/=(float s);
    __device__ __forceinline__ vec2& operator+=(const vprod& p);
    __device__ __forceinline__ vec2& operator-=(const vprod& p);
};

// a * b not yet multiplied: converts to vec2 with one FMUL2, or is absorbed by a following sum as one FFMA2
struct vprod {
    vec2 a, b;
    __device__ __forceinline__ vprod(vec2 a_, vec2 b_) : a(a_), b(b_) {}
    __device__ __forceinline__ operator vec2() const { return vec2(__fmul2_rn(RFK_F2(a), RFK_F2(b))); }
    __device__ __forceinline__ vec2 yx() const { return vec2(*this).yx(); }
};
__device__ __forceinline__ vec2 rfk_fma2(vec2 a, vec2 b, vec2 c) { return vec2(__ffma2_rn(RFK_F2(a), RFK_F2(b), RFK_F2(c))); }
__device__ __forceinline__ vec2& vec2::operator+=(const vprod& p) { return *this = rfk_fma2(p.a, p.b, *this); }
__device__ __forceinline__ vec2& vec2::operator-=(const vprod& p) { return *this = rfk_fma2(vec2(-p.a.x, -p.a.y), p.b, *this); }

// x' = fma(a, x, fma(c, y, e)), y' = fma(b, x, fma(d, y, f)): the affine lines of the generated text, two FFMA2
__device__ __forceinline__ vec2 rfk_affine(float a, float b, float c, float d, float e, float f, float x, float y) {
    return rfk_fma2(vec2(a, b), vec2(x), rfk_fma2(vec2(c, d), vec2(y), vec2(e, f)));
}

struct ivec2 {
    int x, y;
    ivec2() = default;
    __device__ __forceinline__ ivec2(int a, int b) : x(a), y(b) {}
};

// v.xy must be assignable and alias v.x / v.y (the affine line of every xform writes it)
struct vec3 {
    union {
        struct { float x, y; };
        vec2 xy;
    };
    float z;
    vec3() = default;
    __device__ __forceinline__ vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    __device__ __forceinline__ vec3(vec2 a, float c) : x(a.x), y(a.y), z(c) {}
};

struct vec4 {
    float x, y, z, w;
    vec4() = default;
    __device__ __forceinline__ vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    __device__ __forceinline__ vec4(vec2 a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
};

__device__ __forceinline__ vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
__device__ __forceinline__ vprod operator-(vprod p) { return vprod(-p.a, p.b); }
// sums
__device__ __forceinline__ vec2 operator+(vec2 a, vec2 b) { return vec2(__fadd2_rn(RFK_F2(a), RFK_F2(b))); }
__device__ __forceinline__ vec2 operator+(vec2 a, float s) { return vec2(__fadd2_rn(RFK_F2(a), make_float2(s, s))); }
__device__ __forceinline__ vec2 operator+(float s, vec2 a) { return vec2(__fadd2_rn(make_float2(s, s), RFK_F2(a))); }
__device__ __forceinline__ vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
__device__ __forceinline__ vec2 operator-(float s, vec2 a) { return vec2(s - a.x, s - a.y); }
// products stay lazy until a sum or a conversion needs them
__device__ __forceinline__ vprod operator*(vec2 a, vec2 b) { return vprod(a, b); }
__device__ __forceinline__ vprod operator*(vec2 a, float s) { return vprod(a, vec2(s)); }
__device__ __forceinline__ vprod operator*(float s, vec2 a) { return vprod(vec2(s), a); }
__device__ __forceinline__ vprod operator*(vprod p, float s) { return vprod(vec2(p), vec2(s)); }
__device__ __forceinline__ vprod operator*(float s, vprod p) { return vprod(vec2(s), vec2(p)); }
__device__ __forceinline__ vprod operator*(vprod p, vec2 b) { return vprod(vec2(p), b); }
__device__ __forceinline__ vprod operator*(vec2 a, vprod p) { return vprod(a, vec2(p)); }
__device__ __forceinline__ vprod operator*(vprod p, vprod q) { return vprod(vec2(p), vec2(q)); }
// product + addend: one FFMA2
__device__ __forceinline__ vec2 operator+(vprod p, vec2 c) { return rfk_fma2(p.a, p.b, c); }
__device__ __forceinline__ vec2 operator+(vec2 c, vprod p) { return rfk_fma2(p.a, p.b, c); }
__device__ __forceinline__ vec2 operator+(vprod p, vprod q) { return rfk_fma2(p.a, p.b, vec2(q)); }
__device__ __forceinline__ vec2 operator+(vprod p, float s) { return rfk_fma2(p.a, p.b, vec2(s)); }
__device__ __forceinline__ vec2 operator+(float s, vprod p) { return rfk_fma2(p.a, p.b, vec2(s)); }
__device__ __forceinline__ vec2 operator-(vprod p, vec2 c) { return rfk_fma2(p.a, p.b, -c); }
__device__ __forceinline__ vec2 operator-(vec2 c, vprod p) { return rfk_fma2(-p.a, p.b, c); }
__device__ __forceinline__ vec2 operator-(vprod p, vprod q) { return rfk_fma2(p.a, p.b, -vec2(q)); }
__device__ __forceinline__ vec2 operator-(vprod p, float s) { return rfk_fma2(p.a, p.b, vec2(-s)); }
__device__ __forceinline__ vec2 operator-(float s, vprod p) { return rfk_fma2(-p.a, p.b, vec2(s)); }
#if RFK_MATH_MODE == 0
__device__ __forceinline__ vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
__device__ __forceinline__ vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
__device__ __forceinline__ vec2 operator/(float s, vec2 a) { return vec2(s / a.x, s / a.y); }
__device__ __forceinline__ vec2 operator/(vprod p, float s) { return vec2(p) / s; }
#else
// two quotients by the same scalar share one reciprocal (each within 2 ulp of the IEEE quotient); still a lazy product
__device__ __forceinline__ vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
__device__ __forceinline__ vprod operator/(vec2 a, float s) { return vprod(a, vec2(1.0f / s)); }
__device__ __forceinline__ vec2 operator/(float s, vec2 a) { return vec2(s / a.x, s / a.y); }
__device__ __forceinline__ vprod operator/(vprod p, float s) { return vprod(vec2(p), vec2(1.0f / s)); }
#endif
__device__ __forceinline__ vec2 operator/(vprod p, vec2 b) { return vec2(p) / b; }
__device__ __forceinline__ vec2 operator/(vec2 a, vprod p) { return a / vec2(p); }
__device__ __forceinline__ vec2 operator/(float s, vprod p) { return s / vec2(p); }
__device__ __forceinline__ vec2& vec2::operator/=(float s) { return *this = vec2(*this / s); }
// `e / rfk_cfp[n]` in the generated text (a divisor that is one warp-uniform parameter slot) is emitted as
// `e RFK_DIVC(n, r)`: the quotient in mode 0, a product with the reciprocal the host stored in rfk_cfp[r] otherwise.
// RFK_MIXC(z, colour, speed, 1 - speed, colour * speed): the colour blend of every xform, mix(z, colour, speed), from
// the two host-derived constants — one FFMA.
#if RFK_MATH_MODE == 0
#define RFK_DIVC(n, r) / rfk_cfp[n]
#define RFK_MIXC(z, c, t, one_minus_t, c_times_t) mix(z, c, t)
#else
#define RFK_DIVC(n, r) * rfk_cfp[r]
#define RFK_MIXC(z, c, t, one_minus_t, c_times_t) ::fmaf(z, one_minus_t, c_times_t)
#endif

// math.glsl:1-4
static constexpr float PI = 3.141592653589793f;
static constexpr float PI_2 = PI / 2.0f;  // used by some variations only
static constexpr float EPS = (1e-10f);

// RFK_MATH_MODE 0: libdevice functions (1-2 ulp), IEEE division and square root.
// RFK_MATH_MODE 1: the same accuracy class against the 1e-5 parity contract at a fraction of the
//   instructions: 2-ulp division / square root (compiler flags), sine and cosine on the SFU after a
//   two-constant Cody-Waite reduction to [-pi, pi] (absolute error ~5e-7), denormals flushed, pow through lg2/ex2 while |y| <= 16 and x is not negative (relative error < 3e-6, libdevice otherwise), a polynomial atan2
//   (absolute error 1.3e-7), vec2 / scalar through one reciprocal.
// RFK_MATH_MODE 2: --use_fast_math (SFU intrinsics with no range reduction; outside the parity contract).
#if RFK_MATH_MODE == 1
// Mode 1 is compiled with --use_fast_math for ONE of its effects: `a / b` becomes div.approx (one MUFU.RCP and a
// multiply, 2 ulp for |b| in [2^-126, 2^126], 0 beyond) instead of the range-scaled div.full of --prec-div=false
// (four more instructions per quotient). The flag would also turn expf / logf / tanf / powf into their SFU
// intrinsics, which are outside the 1e-5 contract, so mode 1 calls the libdevice entry points by name.
extern "C" __device__ float __nv_expf(float);
extern "C" __device__ float __nv_logf(float);
extern "C" __device__ float __nv_tanf(float);
extern "C" __device__ float __nv_powf(float, float);
#define RFK_EXPF(x) __nv_expf(x)
#define RFK_LOGF(x) __nv_logf(x)
#define RFK_TANF(x) __nv_tanf(x)
#define RFK_POWF(x, y) __nv_powf(x, y)
#else
#define RFK_EXPF(x) ::expf(x)
#define RFK_LOGF(x) ::logf(x)
#define RFK_TANF(x) ::tanf(x)
#define RFK_POWF(x, y) ::powf(x, y)
#endif
#if RFK_MATH_MODE == 1
// x - rint(x / 2pi) * 2pi with 2pi split in two binary32 constants: the product k * hi is exact inside the
// fma, so the reduced argument is good to ~3e-6 rad even at |x| = 1e9 (where binary32 itself resolves 64 rad);
// no large-argument fallback branch is needed. Inf / NaN give NaN like sinf / cosf.
__device__ __forceinline__ float rfk_reduce_2pi(float v) {
    float k = ::rintf(v * 0.15915494309189535f);
    float r = ::fmaf(k, -6.2831854820251465f, v);
    return ::fmaf(k, 1.7484555e-7f, r);
}
__device__ __forceinline__ float sin(float v) { return ::__sinf(rfk_reduce_2pi(v)); }
__device__ __forceinline__ float cos(float v) { return ::__cosf(rfk_reduce_2pi(v)); }
__device__ __forceinline__ void rfk_sincos(float v, float* s, float* c) {
    float r = rfk_reduce_2pi(v);
    *s = ::__sinf(r);
    *c = ::__cosf(r);
}
// `!(a < 0)` rather than `a >= 0`: a NaN base takes the two-SFU path as well (NaN in, NaN out). The reference never resets
// a particle that went non-finite (the badval line of flame.glsl:70 is commented out), so on genomes that overflow — a
// quarter of the stress genome's particles — every warp had lanes in libdevice's 65-instruction powf on every iteration.
__device__ __forceinline__ float pow(float a, float b) {
    if (!(a < 0.0f) && ::fabsf(b) <= 16.0f) return ::exp2f(b * ::__log2f(a));
    return RFK_POWF(a, b);
}
// atan2 from a degree-15 odd polynomial on [0, 1] (least-squares fit on Chebyshev nodes; absolute error 1.3e-7 in
// binary32 Horner form) plus octant fix-ups: ~20 instructions instead of libdevice's ~35. atan2(0, 0) = 0.
__device__ __forceinline__ float rfk_atan2(float y, float x) {
    const float ax = ::fabsf(x), ay = ::fabsf(y);
    const float mx = ::fmaxf(ax, ay), mn = ::fminf(ax, ay);
    const float t = mx > 0.0f ? mn / mx : 0.0f;
    const float s = t * t;
    float p = -0.003960257396101952f;
    p = ::fmaf(p, s, 0.021509254351258278f);
    p = ::fmaf(p, s, -0.05538169667124748f);
    p = ::fmaf(p, s, 0.09601656347513199f);
    p = ::fmaf(p, s, -0.13892041146755219f);
    p = ::fmaf(p, s, 0.19943080842494965f);
    p = ::fmaf(p, s, -0.33329537510871887f);
    p = ::fmaf(p, s, 0.9999992251396179f);
    float r = p * t;
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.0f) r = 3.141592653589793f - r;
    return ::copysignf(r, y);
}
__device__ __forceinline__ float atan(float a, float b) { return rfk_atan2(a, b); }
#else
__device__ __forceinline__ float atan(float a, float b) { return ::atan2f(a, b); }
__device__ __forceinline__ float sin(float v) { return ::sinf(v); }
__device__ __forceinline__ float cos(float v) { return ::cosf(v); }
__device__ __forceinline__ void rfk_sincos(float v, float* s, float* c) { ::sincosf(v, s, c); }
__device__ __forceinline__ float pow(float a, float b) { return ::powf(a, b); }
#endif
#if RFK_MATH_MODE == 1
// tan stays on libdevice: next to its poles the SFU sine / cosine quotient loses too much (popcorn, sin(tan(3y)))
__device__ __forceinline__ float tan(float v) { return RFK_TANF(v); }
// sinh / cosh from one exponential and its reciprocal: relative error ~|x| * 1e-7 for |x| >= 1, absolute error ~1e-7
// below (sinh(x) loses relative accuracy to cancellation there; the parity contract is absolute for small outputs)
__device__ __forceinline__ void rfk_sinhcosh(float v, float* sh, float* ch) {
    float e = RFK_EXPF(v);
    float inv = 1.0f / e;
    *sh = ::fabsf(v) < 0.125f ? v * ::fmaf(v * v, 0.16666667f, 1.0f) : 0.5f * (e - inv);
    *ch = 0.5f * (e + inv);
}
__device__ __forceinline__ float sinh(float v) { float s, c; rfk_sinhcosh(v, &s, &c); return s; }
__device__ __forceinline__ float cosh(float v) { float s, c; rfk_sinhcosh(v, &s, &c); return c; }
#else
__device__ __forceinline__ float tan(float v) { return ::tanf(v); }
__device__ __forceinline__ float sinh(float v) { return ::sinhf(v); }
__device__ __forceinline__ float cosh(float v) { return ::coshf(v); }
__device__ __forceinline__ void rfk_sinhcosh(float v, float* sh, float* ch) { *sh = ::sinhf(v); *ch = ::coshf(v); }
#endif
__device__ __forceinline__ float exp(float v) { return RFK_EXPF(v); }
__device__ __forceinline__ float log(float v) { return RFK_LOGF(v); }
__device__ __forceinline__ float sqrt(float v) { return ::sqrtf(v); }
__device__ __forceinline__ float atan(float a) { return ::atanf(a); }
__device__ __forceinline__ float acos(float v) { return ::acosf(v); }
__device__ __forceinline__ float asin(float v) { return ::asinf(v); }
__device__ __forceinline__ float floor(float v) { return ::floorf(v); }
__device__ __forceinline__ float ceil(float v) { return ::ceilf(v); }
__device__ __forceinline__ float trunc(float v) { return ::truncf(v); }
// GLSL round(): the tie direction is the implementation's choice; nearest-even (one FRND)
__device__ __forceinline__ float round(float v) { return ::rintf(v); }
__device__ __forceinline__ float abs(float v) { return ::fabsf(v); }
__device__ __forceinline__ int abs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
__device__ __forceinline__ float min(float a, float b) { return ::fminf(a, b); }
__device__ __forceinline__ float max(float a, float b) { return ::fmaxf(a, b); }
__device__ __forceinline__ int min(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int max(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ float clamp(float v, float lo, float hi) { return ::fminf(::fmaxf(v, lo), hi); }
__device__ __forceinline__ float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float sign(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
__device__ __forceinline__ float fract(float v) { return v - ::floorf(v); }
__device__ __forceinline__ float mod(float a, float b) { return a - b * ::floorf(a / b); }
__device__ __forceinline__ float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float length(vec2 v) { return ::sqrtf(v.x * v.x + v.y * v.y); }

__device__ __forceinline__ vec2 sin(vec2 v) { return vec2(sin(v.x), sin(v.y)); }
__device__ __forceinline__ vec2 cos(vec2 v) { return vec2(cos(v.x), cos(v.y)); }
__device__ __forceinline__ vec2 abs(vec2 v) { return vec2(::fabsf(v.x), ::fabsf(v.y)); }
__device__ __forceinline__ vec2 mix(vec2 a, vec2 b, float t) { return vec2(mix(a.x, b.x, t), mix(a.y, b.y, t)); }

// math.glsl:6-22
__device__ __forceinline__ vec2 sincos(float v) {
    float s, c;
    rfk_sincos(v, &s, &c);
    return vec2(s, c);
}
__device__ __forceinline__ vec2 sinhcosh(float v) {
    float s, c;
    rfk_sinhcosh(v, &s, &c);
    return vec2(s, c);
}
__device__ __forceinline__ float mod2(float x, float y) { return x - y * ::truncf(x / y); }
__device__ __forceinline__ float log10(float x) { return RFK_LOGF(x) * 0.434294481903251827651128918916f; }
__device__ __forceinline__ bool badval(float x) { return (x != x) || (x > 1e10f) || (x < -1e10f); }

// random.glsl:29-41. The device generator returns `a` after the update (the host
// one in src/util.hpp:81-88 returns `d`); randf is float(u)/4294967295.0f, and that
// divisor rounds to 2^32 in binary32, so the quotient is an exact scaling.
__device__ __forceinline__ uint rfk_rot32(uint x, int k) { return (x << k) | (x >> (32 - k)); }
__device__ __forceinline__ uint rfk_ranval(rfk_rng& s) {
    uint e = s.x - rfk_rot32(s.y, 27);
    s.x = s.y ^ rfk_rot32(s.z, 17);
    s.y = s.z + s.w;
    s.z = s.w + e;
    s.w = e + s.x;
    return s.x;
}
__device__ __forceinline__ float rfk_randf(rfk_rng& s) {
    // clamp(float(u) / 4294967295.0f, 0, 1): the quotient is float(u) * 2^-32, already inside [0, 1]
    return __uint2float_rn(rfk_ranval(s)) * 2.3283064365386963e-10f;
}

}  // namespace rfk_glsl
