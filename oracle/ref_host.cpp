// TEST INFRASTRUCTURE. The reference's HOST code — src/flame.cpp, src/variation_table.cpp, src/util.cpp (genome parser,
// buffer map, parameter buffer, flame compiler) — compiled from the sources where they lie and driven to the point where
// it would hand its generated shader and its parameter buffer to OpenGL. Its third-party dependencies are replaced by
// stand-ins under oracle/stubs/: glad / glm (empty or capturing inline templates), pugixml and yaml-cpp (thin adapters
// over the product's own XML / YAML readers), fmt (a "{}" formatter), inja (returns a marker instead of rendering the two
// templates), nlohmann/json (the real header, from the image's cudnn_frontend). What this pins, with the reference's own
// code doing the work: attribute handling of load_flame, make_shader_buffer_map, copy_flame_data_to_buffer, the
// optimisers and precalc ordering of compile_flame_xforms, `$include_` resolution — i.e. the complete GLSL text of the
// iterate shader and the fp[] upload. Used by tests/golden/make_reference_golden.py to write committed fixtures (the GPU
// box has no /root/reference), never by the product.
#include <any>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <regex>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <vector>
#include <nlohmann/json.hpp>
#include <glad/glad.h>
#include <glm/mat4x4.hpp>
#include <glm/gtc/type_ptr.hpp>

#define private public  // flame's constructor, buffer_map_ and copy_flame_data_to_buffer() are private
#define protected public
#include "util.hpp"
#include "flame.hpp"
#include "variation_table.hpp"
#undef private
#undef protected

#include "buffer_cache.hpp"
void buffer_cache::buffer_group::write_buffer_impl(const char*, std::size_t, std::size_t, std::string) const {}  // never reached here

namespace {
std::string g_out;
int finish(const std::string& s, char* out, long cap) {
    if ((long)s.size() + 1 > cap) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
std::string hex_float(float v) {
    unsigned u;
    std::memcpy(&u, &v, 4);
    char b[16];
    std::snprintf(b, sizeof b, "%08x", u);
    return b;
}
}  // namespace

extern "C" {
// Loads `genome_path` with the reference's flame::load_flame (working directory must be the reference root: it reads
// variations.yaml and shaders/ relative to it) and returns a JSON document with everything it produced.
long ref_host_describe(const char* genome_path, unsigned long W, unsigned long H, char* out, long cap) {
    rfk_gl_capture::shader_sources.clear();
    rfk_gl_capture::uploads.clear();
    flame_compiler vt{};
    auto f = flame::load_flame(genome_path, vt);
    if (!f) return finish("{\"loaded\": false}", out, cap);
    nlohmann::json j;
    j["loaded"] = true;
    j["buffer_map"] = f->buffer_map_;
    j["iterate_shader"] = rfk_gl_capture::shader_sources.empty() ? "" : rfk_gl_capture::shader_sources.front();
    j["compile_flame_xforms"] = vt.compile_flame_xforms(*f);
    rfk_gl_capture::uploads.clear();
    f->copy_flame_data_to_buffer();
    std::vector<std::string> fp;
    if (!rfk_gl_capture::uploads.empty()) {
        const auto& bytes = rfk_gl_capture::uploads.back();
        for (std::size_t i = 0; i + 4 <= bytes.size(); i += 4) { float v; std::memcpy(&v, bytes.data() + i, 4); fp.push_back(hex_float(v)); }
    }
    fp.resize(f->buffer_map_["size"].get<int>());  // the rest of the 1024-float stack buffer is uninitialised in the reference
    j["fp"] = fp;
    auto arr = [](const auto& a) { std::vector<std::string> v; for (auto x : a) v.push_back(hex_float(x)); return v; };
    j["center"] = arr(f->center); j["size"] = {f->size[0], f->size[1]};
    j["scale"] = hex_float(f->scale); j["rotate"] = hex_float(f->rotate);
    j["estimator_min"] = f->estimator_min; j["estimator_radius"] = f->estimator_radius; j["estimator_curve"] = hex_float(f->estimator_curve);
    j["gamma"] = hex_float(f->gamma); j["vibrancy"] = hex_float(f->vibrancy); j["brightness"] = hex_float(f->brightness);
    nlohmann::json xs = nlohmann::json::array();
    f->for_each_xform([&](int idx, flame_xform& x) {
        nlohmann::json o;
        o["index"] = idx; o["affine"] = arr(x.affine);
        if (x.post) o["post"] = arr(*x.post);
        for (auto& [k, v] : x.variations) o["variations"][k] = hex_float(v);
        for (auto& [k, v] : x.var_param) o["var_param"][k] = hex_float(v);
        o["weight"] = hex_float(x.weight); o["color"] = hex_float(x.color); o["color_speed"] = hex_float(x.color_speed);
        o["rotation_frequency"] = hex_float(x.rotation_frequency); o["opacity"] = hex_float(x.opacity);
        xs.push_back(o);
    });
    j["xforms"] = xs;
    std::vector<std::string> pal;
    for (auto& c : f->palette) for (float v : c) pal.push_back(hex_float(v));
    j["palette"] = pal;
    // src/flame.cpp:289-296 (the first statements of draw_to_bins), with the reference's helpers
    std::size_t target_dims[2] = {W, H};
    flame_xform::affine_t base{1, 0, 0, 1, 0, 0};
    base = flame::translate_affine(base, {target_dims[0] / 2.0f, target_dims[1] / 2.0f});
    base = flame::scale_affine(base, f->scale * float(target_dims[1]) / float(f->size[1]));
    base = flame::rotate_affine(base, f->rotate);
    base = flame::translate_affine(base, {-f->center[0], -f->center[1]});
    j["ss_affine"] = arr(base);
    return finish(j.dump(), out, cap);
}
}
