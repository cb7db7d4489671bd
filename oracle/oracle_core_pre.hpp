// TEST INFRASTRUCTURE — declarations the generated GLSL needs before its own definitions
// (the globals and randf() of shaders/include/random.glsl, shaders/flame.glsl:13,27).
#pragma once
#include "glsl_shim.hpp"
namespace glsl {
static thread_local uvec4 local_random_state;  // random.glsl:19
static thread_local const float* fp;           // flame.glsl:27 (shared float fp[1024])
static thread_local bool first_run;            // flame.glsl:13
inline uint rot32(uint x, int k) { return (x << k) | (x >> (32 - k)); }
// random.glsl:29-41
inline uint ranval() {
    uint e = local_random_state.x - rot32(local_random_state.y, 27);
    local_random_state.x = local_random_state.y ^ rot32(local_random_state.z, 17);
    local_random_state.y = local_random_state.z + local_random_state.w;
    local_random_state.z = local_random_state.w + e;
    local_random_state.w = e + local_random_state.x;
    return local_random_state.x;
}
inline float randf() { return clamp(float(ranval()) / 4294967295.0f, 0.0f, 1.0f); }
}  // namespace glsl
