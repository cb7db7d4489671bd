// TEST INFRASTRUCTURE — a software stand-in for the slice of OpenGL 4.6 the reference's render path drives
// (src/flame.cpp, src/shaders.hpp, src/buffer_objects.hpp): named buffers, shader-storage bindings, compute / vertex /
// fragment programs, uniforms, glDispatchCompute and a GL_POINTS draw with additive blending into an RGBA32F target.
// Shaders are the reference's own GLSL text, turned into shared libraries by glsl_to_cpp.py at glLinkProgram time.
// With it the reference's host code (set_sim_parameters, load_flame, warmup, draw_to_bins) runs UNMODIFIED on the CPU and
// its buffers can be read back: that is what pins the oracle's restatement of the device half of the path.
// oracle/stubs/glad/glad.h forwards the GL entry points here when RFK_SOFTGL is defined. Never part of the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace softgl {
typedef unsigned int uint;

// --- buffers
void create_buffers(int n, uint* names);
void buffer_storage(uint name, std::ptrdiff_t bytes, const void* data);
void buffer_sub_data(uint name, std::ptrdiff_t offset, std::ptrdiff_t bytes, const void* data);
void get_buffer_sub_data(uint name, std::ptrdiff_t offset, std::ptrdiff_t bytes, void* out);
void clear_buffer(uint name);
void delete_buffers(int n, const uint* names);
void* map_buffer(uint name);
void bind_buffer_base(uint index, uint name);
std::size_t buffer_size(uint name);
void* buffer_data(uint name);
uint bound_buffer(uint index);

// --- shaders and programs
uint create_shader(uint type);
void shader_source(uint shader, const char* text);
uint create_program();
void attach_shader(uint program, uint shader);
void link_program(uint program);
int link_status(uint program);
std::string program_log(uint program);
void use_program(uint program);
int uniform_location(uint program, const char* name);
void set_uniform(int location, const void* data, std::size_t bytes);  // on the program in use
void dispatch_compute(uint nx, uint ny, uint nz);

// --- GL_POINTS into an RGBA32F colour target with glBlendFunc(GL_ONE, GL_ONE) (density estimation, main.cpp:490-519)
void set_color_target(float* rgba, int width, int height);
void draw_points(int first, int count);

// --- what the reference handed over, for the harness
struct uniform_record { uint program; std::string name; std::vector<unsigned char> bytes; };
struct dispatch_record { uint program; uint nx, ny, nz; };
std::vector<std::string>& shader_sources();
std::vector<uniform_record>& uniform_log();
std::vector<dispatch_record>& dispatch_log();
void reset_logs();
}  // namespace softgl
