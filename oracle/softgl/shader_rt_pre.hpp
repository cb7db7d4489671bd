// TEST INFRASTRUCTURE — first half of the run time that lets a GLSL 4.60 shader of the reference, transliterated to C++
// by glsl_to_cpp.py, execute on the CPU: the built-in variables of the three shader stages the path uses and the image
// type of tonemap.glsl. Included BEFORE the shader text (inside no namespace; it opens `namespace glsl` itself).
// The second half (shader_rt_post.hpp) holds the invocation scheduler. Nothing here is part of the product.
#pragma once
#define GLSL_SHIM_NO_MATH_GLSL  // the shader text brings its own math.glsl
#include "../glsl_shim.hpp"
#include <cstddef>
#include <cstring>

namespace glsl {

// One shader invocation: the per-invocation built-ins, plus a pointer to the shader's own unqualified globals
// (GLSL gives every invocation its own copy; the transliteration gathers them in `struct rfk_private`).
struct rfk_invocation {
    uvec3 local_id, global_id;   // compute
    int vertex_id = 0;           // vertex
    vec4 position;               // vertex out
    float point_size = 1.0f;     // vertex out
    vec2 point_coord;            // fragment in
    bool discarded = false;      // fragment
    void* priv = nullptr;
};
static rfk_invocation* rfk_cur = nullptr;

#ifndef RFK_LOCAL_X
#define RFK_LOCAL_X 1
#define RFK_LOCAL_Y 1
#define RFK_LOCAL_Z 1
#endif
static const uvec3 gl_WorkGroupSize(RFK_LOCAL_X, RFK_LOCAL_Y, RFK_LOCAL_Z);
static uvec3 gl_NumWorkGroups, gl_WorkGroupID;
#define gl_LocalInvocationID (rfk_cur->local_id)
#define gl_GlobalInvocationID (rfk_cur->global_id)
#define gl_VertexID (rfk_cur->vertex_id)
#define gl_Position (rfk_cur->position)
#define gl_PointSize (rfk_cur->point_size)
#define gl_PointCoord (rfk_cur->point_coord)

void barrier();                                  // shader_rt_post.hpp
inline void rfk_discard() { rfk_cur->discarded = true; }

// layout(rgba32f) image2D: a W x H array of vec4 owned by the soft GL
struct image2D { vec4* texels = nullptr; int width = 0, height = 0; };
inline vec4 imageLoad(const image2D& img, const ivec2& p) {
    if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height) return vec4();
    return img.texels[(size_t)p.y * img.width + p.x];
}
inline void imageStore(image2D& img, const ivec2& p, const vec4& v) {
    if (p.x < 0 || p.y < 0 || p.x >= img.width || p.y >= img.height) return;
    img.texels[(size_t)p.y * img.width + p.x] = v;
}

struct rfk_binding { int binding; void** pointer; };
struct rfk_uniform { const char* name; void* pointer; size_t bytes; };

}  // namespace glsl
