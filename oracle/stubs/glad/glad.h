// TEST INFRASTRUCTURE. Declarations-only stand-in for glad (OpenGL loader), just enough for the reference's
// headers (src/buffer_objects.hpp, src/shaders.hpp) to parse so that the PURE inline helpers of src/flame.hpp
// (rotate_affine / scale_affine / translate_affine) can be compiled from the reference source and used as pins.
// Every GL entry point is an empty inline template: nothing here touches a GL context, and the wrappers in
// ref_pins_flame.cpp only call the reference's pure arithmetic helpers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <array>
#include <iostream>
#include <cmath>
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef int GLsizei; typedef float GLfloat;
typedef unsigned char GLboolean; typedef char GLchar; typedef std::ptrdiff_t GLsizeiptr; typedef std::ptrdiff_t GLintptr; typedef unsigned int GLbitfield;
#define RFK_GLC(name, value) static const GLenum name = value;
RFK_GLC(GL_DYNAMIC_STORAGE_BIT, 0x0100) RFK_GLC(GL_DYNAMIC_COPY, 0x88EA) RFK_GLC(GL_R32UI, 0x8236) RFK_GLC(GL_RED_INTEGER, 0x8D94) RFK_GLC(GL_UNSIGNED_INT, 0x1405)
RFK_GLC(GL_UNIFORM_BUFFER, 0x8A11) RFK_GLC(GL_TEXTURE_2D, 0x0DE1) RFK_GLC(GL_RGBA32F, 0x8814) RFK_GLC(GL_RGBA32UI, 0x8D70) RFK_GLC(GL_RGBA32I, 0x8D82)
RFK_GLC(GL_RGBA, 0x1908) RFK_GLC(GL_UNSIGNED_BYTE, 0x1401) RFK_GLC(GL_FRAMEBUFFER, 0x8D40) RFK_GLC(GL_COLOR_ATTACHMENT0, 0x8CE0) RFK_GLC(GL_FLOAT, 0x1406)
RFK_GLC(GL_TEXTURE_MIN_FILTER, 0x2801) RFK_GLC(GL_TEXTURE_MAG_FILTER, 0x2800) RFK_GLC(GL_LINEAR, 0x2601) RFK_GLC(GL_NEAREST, 0x2600)
RFK_GLC(GL_TEXTURE_WRAP_S, 0x2802) RFK_GLC(GL_TEXTURE_WRAP_T, 0x2803) RFK_GLC(GL_CLAMP_TO_EDGE, 0x812F) RFK_GLC(GL_CLAMP_TO_BORDER, 0x812D)
RFK_GLC(GL_COMPUTE_SHADER, 0x91B9) RFK_GLC(GL_VERTEX_SHADER, 0x8B31) RFK_GLC(GL_FRAGMENT_SHADER, 0x8B30) RFK_GLC(GL_COMPILE_STATUS, 0x8B81)
RFK_GLC(GL_LINK_STATUS, 0x8B82) RFK_GLC(GL_INFO_LOG_LENGTH, 0x8B84) RFK_GLC(GL_FALSE, 0) RFK_GLC(GL_TRUE, 1) RFK_GLC(GL_INT, 0x1404)
RFK_GLC(GL_SHADER_STORAGE_BUFFER, 0x90D2) RFK_GLC(GL_DRAW_FRAMEBUFFER, 0x8CA9) RFK_GLC(GL_READ_FRAMEBUFFER, 0x8CA8) RFK_GLC(GL_RGBA_INTEGER, 0x8D99)
RFK_GLC(GL_TEXTURE_BORDER_COLOR, 0x1004) RFK_GLC(GL_REPEAT, 0x2901) RFK_GLC(GL_FRAMEBUFFER_COMPLETE, 0x8CD5)
// any GL call in an inline body of the reference's headers resolves to an empty variadic template
#define RFK_GLF(name) template <typename... A> void name(A...) {}
RFK_GLF(glCreateTextures) RFK_GLF(glTextureStorage2D)
RFK_GLF(glDeleteTextures) RFK_GLF(glGetTextureImage) RFK_GLF(glTextureParameteri) RFK_GLF(glTextureParameterfv) RFK_GLF(glCreateFramebuffers) RFK_GLF(glDeleteFramebuffers)
RFK_GLF(glBindFramebuffer) RFK_GLF(glNamedFramebufferTexture) RFK_GLF(glFramebufferTexture) RFK_GLF(glFramebufferTexture2D) RFK_GLF(glCompileShader) RFK_GLF(glGetShaderInfoLog) RFK_GLF(glDeleteShader)
RFK_GLF(glDeleteProgram) RFK_GLF(glDetachShader)
RFK_GLF(glDrawBuffers) RFK_GLF(glBindTexture) RFK_GLF(glClearTexImage) RFK_GLF(glBindBuffer) RFK_GLF(glGenFramebuffers) RFK_GLF(glTexImage2D) RFK_GLF(glGenTextures)
template <typename... A> GLenum glCheckFramebufferStatus(A...) { return 0; }
template <typename... A> GLenum glCheckNamedFramebufferStatus(A...) { return 0; }
// further names the reference's headers mention (values are never used)
RFK_GLC(GL_DYNAMIC_DRAW, 0)
RFK_GLC(GL_R8, 0)
RFK_GLC(GL_READ_WRITE, 0)
RFK_GLC(GL_RED, 0)
RFK_GLC(GL_RGB, 0)
RFK_GLC(GL_TEXTURE_BINDING_2D, 0)
RFK_GLF(glGenerateMipmap)
RFK_GLF(glGetIntegerv)
RFK_GLF(glTexParameteri)
RFK_GLF(glUnmapNamedBuffer)

RFK_GLC(GL_ALL_BARRIER_BITS, 0)
RFK_GLF(glFinish)
RFK_GLF(glMemoryBarrier)

// ---------------------------------------------------------------------------------------------------------------
// The entry points of the render path. Two builds:
//  * default: empty / capturing templates — what the reference hands to GL is recorded (shader text, parameter upload),
//    nothing executes (libref_pins_flame.so);
//  * RFK_SOFTGL: forwarded to the software GL of oracle/softgl/, which executes the reference's shaders on the CPU
//    (libref_host.so).
#include <string>
#include <cstring>
#ifndef RFK_SOFTGL
RFK_GLF(glCreateBuffers) RFK_GLF(glNamedBufferStorage) RFK_GLF(glNamedBufferData) RFK_GLF(glDeleteBuffers) RFK_GLF(glGetNamedBufferSubData) RFK_GLF(glClearNamedBufferData) 
RFK_GLF(glBindBufferBase) RFK_GLF(glUniform1i) RFK_GLF(glUniform1ui) RFK_GLF(glUniform1f) RFK_GLF(glUniform2fv) RFK_GLF(glUniform3fv) 
RFK_GLF(glUniform4fv) RFK_GLF(glUniform2uiv) RFK_GLF(glUniform3uiv) RFK_GLF(glUniform4uiv) RFK_GLF(glUniform2iv) RFK_GLF(glUniform3iv) 
RFK_GLF(glUniform4iv) RFK_GLF(glUniformMatrix4fv) RFK_GLF(glUseProgram) RFK_GLF(glAttachShader) RFK_GLF(glLinkProgram) RFK_GLF(glGetProgramInfoLog) 
RFK_GLF(glMapNamedBuffer) RFK_GLF(glDispatchCompute) 
template <typename... A> GLuint glCreateShader(A...) { return 0; }
template <typename... A> GLuint glCreateProgram(A...) { return 0; }
template <typename... A> GLint glGetUniformLocation(A...) { return 0; }
namespace rfk_gl_capture {
inline std::vector<std::string> shader_sources;            // every glShaderSource string, in call order
inline std::vector<std::vector<unsigned char>> uploads;    // every glNamedBufferSubData payload
inline std::vector<std::vector<float>> uniform_arrays;     // every glUniform1fv payload
}
template <typename A, typename B, typename C, typename D> void glShaderSource(A, B, C strings, D) { rfk_gl_capture::shader_sources.emplace_back(strings[0]); }
template <typename A, typename B, typename C> void glGetShaderiv(A, B, C out) { *out = 1; }
template <typename A, typename B, typename C> void glGetProgramiv(A, B, C out) { *out = 1; }
template <typename A, typename B, typename C, typename D> void glNamedBufferSubData(A, B, C size, D ptr) {
    const unsigned char* p = static_cast<const unsigned char*>(static_cast<const void*>(ptr));
    rfk_gl_capture::uploads.emplace_back(p, p + (std::size_t)size);
}
template <typename A, typename B, typename C> void glUniform1fv(A, B n, C ptr) { rfk_gl_capture::uniform_arrays.emplace_back(ptr, ptr + n); }

#else
#include "../../softgl/softgl.hpp"
template <typename N> void glCreateBuffers(int n, N* names) { softgl::create_buffers(n, names); }
template <typename S, typename D, typename F> void glNamedBufferStorage(GLuint name, S bytes, D data, F) { softgl::buffer_storage(name, (std::ptrdiff_t)bytes, (const void*)data); }
template <typename S, typename D, typename F> void glNamedBufferData(GLuint name, S bytes, D data, F) { softgl::buffer_storage(name, (std::ptrdiff_t)bytes, (const void*)data); }
template <typename O, typename S, typename D> void glNamedBufferSubData(GLuint name, O offset, S bytes, D data) { softgl::buffer_sub_data(name, (std::ptrdiff_t)offset, (std::ptrdiff_t)bytes, (const void*)data); }
template <typename O, typename S, typename D> void glGetNamedBufferSubData(GLuint name, O offset, S bytes, D out) { softgl::get_buffer_sub_data(name, (std::ptrdiff_t)offset, (std::ptrdiff_t)bytes, (void*)out); }
template <typename... A> void glClearNamedBufferData(GLuint name, A...) { softgl::clear_buffer(name); }
template <typename N> void glDeleteBuffers(int n, N* names) { softgl::delete_buffers(n, names); }
template <typename A> void* glMapNamedBuffer(GLuint name, A) { return softgl::map_buffer(name); }
inline void glBindBufferBase(GLenum, GLuint index, GLuint name) { softgl::bind_buffer_base(index, name); }
inline GLuint glCreateShader(GLenum type) { return softgl::create_shader(type); }
template <typename C, typename D> void glShaderSource(GLuint shader, int, C strings, D) { softgl::shader_source(shader, strings[0]); }
inline void glGetShaderiv(GLuint, GLenum, int* out) { *out = 1; }  // compilation happens at link time
inline GLuint glCreateProgram() { return softgl::create_program(); }
inline void glAttachShader(GLuint program, GLuint shader) { softgl::attach_shader(program, shader); }
inline void glLinkProgram(GLuint program) { softgl::link_program(program); }
inline void glGetProgramiv(GLuint program, GLenum, int* out) { *out = softgl::link_status(program); }
template <typename L> void glGetProgramInfoLog(GLuint program, int cap, L, char* out) { std::string l = softgl::program_log(program); std::strncpy(out, l.c_str(), cap - 1); out[cap - 1] = 0; }
inline void glUseProgram(GLuint program) { softgl::use_program(program); }
inline GLint glGetUniformLocation(GLuint program, const char* name) { return softgl::uniform_location(program, name); }
inline void glUniform1i(GLint l, int v) { softgl::set_uniform(l, &v, 4); }
inline void glUniform1ui(GLint l, unsigned v) { softgl::set_uniform(l, &v, 4); }
inline void glUniform1f(GLint l, float v) { softgl::set_uniform(l, &v, 4); }
template <typename P> void glUniform1fv(GLint l, int n, P v) { softgl::set_uniform(l, v, 4 * (std::size_t)n); }
template <typename P> void glUniform2fv(GLint l, int n, P v) { softgl::set_uniform(l, v, 8 * (std::size_t)n); }
template <typename P> void glUniform3fv(GLint l, int n, P v) { softgl::set_uniform(l, v, 12 * (std::size_t)n); }
template <typename P> void glUniform4fv(GLint l, int n, P v) { softgl::set_uniform(l, v, 16 * (std::size_t)n); }
template <typename P> void glUniform2uiv(GLint l, int n, P v) { softgl::set_uniform(l, v, 8 * (std::size_t)n); }
template <typename P> void glUniform3uiv(GLint l, int n, P v) { softgl::set_uniform(l, v, 12 * (std::size_t)n); }
template <typename P> void glUniform4uiv(GLint l, int n, P v) { softgl::set_uniform(l, v, 16 * (std::size_t)n); }
template <typename P> void glUniform2iv(GLint l, int n, P v) { softgl::set_uniform(l, v, 8 * (std::size_t)n); }
template <typename P> void glUniform3iv(GLint l, int n, P v) { softgl::set_uniform(l, v, 12 * (std::size_t)n); }
template <typename P> void glUniform4iv(GLint l, int n, P v) { softgl::set_uniform(l, v, 16 * (std::size_t)n); }
template <typename T, typename P> void glUniformMatrix4fv(GLint l, int n, T, P v) { softgl::set_uniform(l, v, 64 * (std::size_t)n); }
inline void glDispatchCompute(GLuint x, GLuint y, GLuint z) { softgl::dispatch_compute(x, y, z); }
#endif
