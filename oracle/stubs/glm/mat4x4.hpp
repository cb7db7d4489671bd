// TEST INFRASTRUCTURE. Type names only (see ../glad/glad.h).
#pragma once
namespace glm {
struct vec2 { float v[2]; }; struct vec3 { float v[3]; }; struct vec4 { float v[4]; };
struct uvec2 { unsigned v[2]; uvec2() = default; template <typename A, typename B> uvec2(A a, B b) : v{(unsigned)a, (unsigned)b} {} }; struct uvec3 { unsigned v[3]; }; struct uvec4 { unsigned v[4]; };
struct ivec2 { int v[2]; }; struct ivec3 { int v[3]; }; struct ivec4 { int v[4]; };
struct mat4 { float v[16]; };
}
