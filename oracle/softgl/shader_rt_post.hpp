// TEST INFRASTRUCTURE — second half of the shader run time (see shader_rt_pre.hpp): the invocation scheduler and the C
// entry points the soft GL (softgl.cpp) resolves with dlsym. Included AFTER the transliterated shader text, which has
// defined, inside namespace glsl: struct rfk_private, void rfk_shader_main(), rfk_bindings[], rfk_uniforms[],
// rfk_varyings[] and the macros RFK_STAGE (0 compute, 1 vertex, 2 fragment) and RFK_USES_BARRIER.
//
// Compute work groups run one after another in x-fastest order; inside a work group every invocation is a ucontext
// fiber, so that barrier() has its GLSL meaning (all invocations of the group reach it before any continues) and
// `shared` variables are plain statics. Invocations run in local-index order between barriers.
#pragma once
#include <ucontext.h>
#include <vector>
#include <cstdlib>

namespace glsl {

#if RFK_USES_BARRIER
namespace rfk_fibers {
static const size_t STACK_BYTES = 256 * 1024;
static ucontext_t scheduler;
static std::vector<ucontext_t> contexts;
static std::vector<char*> stacks;
static std::vector<char> finished;
static std::vector<rfk_invocation> invocations;
static std::vector<rfk_private> privates;
static void trampoline() {
    rfk_shader_main();
    finished[rfk_cur - invocations.data()] = 1;
}
}  // namespace rfk_fibers
void barrier() {
    rfk_invocation* me = rfk_cur;
    swapcontext(&rfk_fibers::contexts[me - rfk_fibers::invocations.data()], &rfk_fibers::scheduler);
    rfk_cur = me;
}
#else
void barrier() {}
#endif

static void rfk_run_group(unsigned gx, unsigned gy, unsigned gz) {
    const unsigned n = RFK_LOCAL_X * RFK_LOCAL_Y * RFK_LOCAL_Z;
    gl_WorkGroupID = uvec3(gx, gy, gz);
#if RFK_USES_BARRIER
    using namespace rfk_fibers;
    if (contexts.size() != n) {
        contexts.resize(n); finished.resize(n); invocations.resize(n); privates.resize(n);
        for (unsigned i = stacks.size(); i < n; i++) stacks.push_back((char*)std::malloc(STACK_BYTES));
    }
    for (unsigned i = 0; i < n; i++) {
        unsigned lx = i % RFK_LOCAL_X, ly = (i / RFK_LOCAL_X) % RFK_LOCAL_Y, lz = i / (RFK_LOCAL_X * RFK_LOCAL_Y);
        invocations[i] = rfk_invocation();
        invocations[i].local_id = uvec3(lx, ly, lz);
        invocations[i].global_id = uvec3(gx * RFK_LOCAL_X + lx, gy * RFK_LOCAL_Y + ly, gz * RFK_LOCAL_Z + lz);
        privates[i] = rfk_private();
        invocations[i].priv = &privates[i];
        finished[i] = 0;
        getcontext(&contexts[i]);
        contexts[i].uc_stack.ss_sp = stacks[i];
        contexts[i].uc_stack.ss_size = STACK_BYTES;
        contexts[i].uc_link = &scheduler;
        makecontext(&contexts[i], trampoline, 0);
    }
    for (bool any = true; any;) {
        any = false;
        for (unsigned i = 0; i < n; i++) {
            if (finished[i]) continue;
            rfk_cur = &invocations[i];
            swapcontext(&scheduler, &contexts[i]);
            any = true;
        }
    }
#else
    rfk_invocation inv;
    rfk_private priv;
    for (unsigned i = 0; i < n; i++) {
        unsigned lx = i % RFK_LOCAL_X, ly = (i / RFK_LOCAL_X) % RFK_LOCAL_Y, lz = i / (RFK_LOCAL_X * RFK_LOCAL_Y);
        inv = rfk_invocation();
        inv.local_id = uvec3(lx, ly, lz);
        inv.global_id = uvec3(gx * RFK_LOCAL_X + lx, gy * RFK_LOCAL_Y + ly, gz * RFK_LOCAL_Z + lz);
        priv = rfk_private();
        inv.priv = &priv;
        rfk_cur = &inv;
        rfk_shader_main();
    }
#endif
    rfk_cur = nullptr;
}

}  // namespace glsl

extern "C" {
int rfk_sh_stage() { return RFK_STAGE; }
void rfk_sh_local_size(unsigned* xyz) { xyz[0] = RFK_LOCAL_X; xyz[1] = RFK_LOCAL_Y; xyz[2] = RFK_LOCAL_Z; }
int rfk_sh_bind(int binding, void* pointer) {
    int hit = 0;
    for (size_t i = 0; i < sizeof(glsl::rfk_bindings) / sizeof(glsl::rfk_bindings[0]); i++)
        if (glsl::rfk_bindings[i].pointer && glsl::rfk_bindings[i].binding == binding) { *glsl::rfk_bindings[i].pointer = pointer; hit = 1; }
    return hit;
}
int rfk_sh_uniform_location(const char* name) {
    for (size_t i = 0; i < sizeof(glsl::rfk_uniforms) / sizeof(glsl::rfk_uniforms[0]); i++)
        if (glsl::rfk_uniforms[i].name && !std::strcmp(glsl::rfk_uniforms[i].name, name)) return (int)i;
    return -1;
}
int rfk_sh_set_uniform(int location, const void* data, size_t bytes) {
    if (location < 0 || (size_t)location >= sizeof(glsl::rfk_uniforms) / sizeof(glsl::rfk_uniforms[0]) || !glsl::rfk_uniforms[location].name) return 0;
    if (bytes > glsl::rfk_uniforms[location].bytes) bytes = glsl::rfk_uniforms[location].bytes;
    std::memcpy(glsl::rfk_uniforms[location].pointer, data, bytes);
    return 1;
}
void rfk_sh_dispatch(unsigned nx, unsigned ny, unsigned nz) {
    glsl::gl_NumWorkGroups = glsl::uvec3(nx, ny, nz);
    for (unsigned z = 0; z < nz; z++)
        for (unsigned y = 0; y < ny; y++)
            for (unsigned x = 0; x < nx; x++) glsl::rfk_run_group(x, y, z);
}
size_t rfk_sh_private_bytes() { return sizeof(glsl::rfk_private); }
int rfk_sh_varyings(const char** names, size_t* offsets, size_t* bytes, int* is_out, int cap) {
    int n = 0;
    for (size_t i = 0; i < sizeof(glsl::rfk_varyings) / sizeof(glsl::rfk_varyings[0]); i++) {
        if (!glsl::rfk_varyings[i].name) continue;
        if (n < cap) { names[n] = glsl::rfk_varyings[i].name; offsets[n] = glsl::rfk_varyings[i].offset; bytes[n] = glsl::rfk_varyings[i].bytes; is_out[n] = glsl::rfk_varyings[i].is_out; }
        n++;
    }
    return n;
}
// vertex stage: runs main() for one vertex; `priv` (rfk_sh_private_bytes() bytes) receives the flat outputs
void rfk_sh_run_vertex(int vertex_id, void* priv, float* position4, float* point_size) {
    glsl::rfk_invocation inv;
    new (priv) glsl::rfk_private();
    inv.vertex_id = vertex_id;
    inv.priv = priv;
    glsl::rfk_cur = &inv;
    glsl::rfk_shader_main();
    glsl::rfk_cur = nullptr;
    position4[0] = inv.position.x; position4[1] = inv.position.y; position4[2] = inv.position.z; position4[3] = inv.position.w;
    *point_size = inv.point_size;
}
// fragment stage: `priv` holds the flat inputs (copied from the vertex outputs by name); returns 0 when discarded
int rfk_sh_run_fragment(float point_coord_x, float point_coord_y, void* priv) {
    glsl::rfk_invocation inv;
    inv.point_coord = glsl::vec2(point_coord_x, point_coord_y);
    inv.priv = priv;
    glsl::rfk_cur = &inv;
    glsl::rfk_shader_main();
    glsl::rfk_cur = nullptr;
    return inv.discarded ? 0 : 1;
}
}
