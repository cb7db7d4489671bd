// TEST INFRASTRUCTURE. The reference's HOST code — src/flame.cpp, src/variation_table.cpp, src/util.cpp, src/hammersley.cpp,
// src/shuffle_buffers.cpp (genome parser, buffer map, parameter buffer, flame compiler, set_sim_parameters, warmup,
// draw_to_bins) — compiled UNMODIFIED from the sources where they lie and run against a software GL (oracle/softgl/) that
// executes the reference's own GLSL text on the CPU. Third-party dependencies are replaced by stand-ins under
// oracle/stubs/: glad (forwards to the soft GL), glm (type names), pugixml and yaml-cpp (thin adapters over the product's
// XML / YAML tokenisers), fmt (a "{}" formatter), inja (hands template + data to oracle/softgl/glsl_to_cpp.py, which
// evaluates the inja subset the two templates use), nlohmann/json (the real header, from the image's cudnn_frontend).
//   ref_host_describe : load_flame only — parsed fields, buffer map, fp[], generated shader text (host-half pins)
//   ref_host_run      : set_sim_parameters + load_flame + warmup + draw_to_bins — shuffle tables, per-pass shuffle ids,
//                       RNG states, particle buffers, fp_inflated, bins, binned counter (device-half pins)
//   ref_host_post     : the density-estimation draw call and the tonemap dispatch of src/main.cpp:490-535 (that host
//                       sequence sits inside main() and is restated here call by call; the shaders are the reference's)
// Used by tests/golden/make_reference_golden.py to write committed fixtures (the GPU box has no /root/reference), never
// by the product. The working directory must hold shaders/ and variations.yaml (links into the reference) and a writable
// cache/ (buffer_cache.hpp).
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
#include "buffer_cache.hpp"
#undef private
#undef protected


// src/buffer_cache.cpp needs xxhash + fmt's width specifiers; this writes the same file layout (size_t byte count, then
// the payload; buffer_cache.cpp:7-21) under a content-hash name of its own.
void buffer_cache::buffer_group::write_buffer_impl(const char* buf, std::size_t total_size, std::size_t, std::string filename) const {
    if (filename.empty()) {
        std::uint64_t h1 = 1469598103934665603ull, h2 = 0x9E3779B97F4A7C15ull;
        for (std::size_t i = 0; i < total_size; i++) { h1 = (h1 ^ (unsigned char)buf[i]) * 1099511628211ull; h2 = (h2 + (unsigned char)buf[i] + i) * 0xD6E8FEB86659FD93ull; }
        char b[40];
        std::snprintf(b, sizeof b, "%016llX%016llX", (unsigned long long)h1, (unsigned long long)h2);
        filename = b;
    }
    std::ofstream f(path_ + filename + ".bin", std::ios::binary);
    f.write((const char*)&total_size, sizeof total_size);
    f.write(buf, total_size);
}

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
    softgl::reset_logs();
    flame_compiler vt{};
    auto f = flame::load_flame(genome_path, vt);
    if (!f) return finish("{\"loaded\": false}", out, cap);
    nlohmann::json j;
    j["loaded"] = true;
    j["buffer_map"] = f->buffer_map_;
    j["iterate_shader"] = softgl::shader_sources().empty() ? "" : softgl::shader_sources().front();
    j["animate_shader"] = softgl::shader_sources().size() < 2 ? "" : softgl::shader_sources()[1];
    j["compile_flame_xforms"] = vt.compile_flame_xforms(*f);
    f->copy_flame_data_to_buffer();
    std::vector<std::string> fp;
    {
        const float* p = static_cast<const float*>(softgl::buffer_data(f->param_buffer_.name()));
        for (std::size_t i = 0; p && i < 1024; i++) fp.push_back(hex_float(p[i]));
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

// ---------------------------------------------------------------------------------------------------------------
// Full passes: the reference's own set_sim_parameters / load_flame / warmup / draw_to_bins on the soft GL.
namespace {
std::unique_ptr<flame_compiler> g_vt;
std::unique_ptr<flame> g_flame;
std::unique_ptr<flame::bin_t> g_bins;
std::size_t g_W = 0, g_H = 0;
std::vector<float> g_de, g_tm;
}

long ref_host_run(const char* genome_path, unsigned long P, unsigned long TS, unsigned long n_shuffle, unsigned long warmup_passes, float tss_width,
                  unsigned long W, unsigned long H, int draw_passes, char* out, long cap) {
    softgl::reset_logs();
    g_flame.reset(); g_bins.reset();
    flame::set_sim_parameters(P, TS, n_shuffle);
    if (!g_vt) g_vt = std::make_unique<flame_compiler>();
    g_flame = flame::load_flame(genome_path, *g_vt);
    if (!g_flame) return finish("{\"loaded\": false}", out, cap);
    g_flame->warmup(warmup_passes, tss_width);
    g_W = W; g_H = H;
    g_bins = std::make_unique<flame::bin_t>(W * H);
    g_bins->zero_out();
    std::size_t binned = draw_passes > 0 ? g_flame->draw_to_bins(*g_bins, W, draw_passes) : 0;
    nlohmann::json j;
    j["loaded"] = true;
    j["binned"] = binned;
    j["total_params"] = g_flame->buffer_map_["size"];
    nlohmann::json ids_in = nlohmann::json::array(), ids_out = nlohmann::json::array(), tsw = nlohmann::json::array();
    for (auto& u : softgl::uniform_log()) {
        if (u.bytes.size() != 4) continue;
        unsigned v; std::memcpy(&v, u.bytes.data(), 4);
        if (u.name == "shuf_buf_idx_in") ids_in.push_back(v);
        if (u.name == "shuf_buf_idx_out") ids_out.push_back(v);
        if (u.name == "temporal_sample_width") tsw.push_back(hex_float(*reinterpret_cast<float*>(&v)));
    }
    j["shuf_buf_idx_in"] = ids_in; j["shuf_buf_idx_out"] = ids_out; j["temporal_sample_width"] = tsw;
    nlohmann::json d = nlohmann::json::array();
    for (auto& r : softgl::dispatch_log()) d.push_back({r.program, r.nx, r.ny, r.nz});
    j["dispatches"] = d;
    return finish(j.dump(), out, cap);
}

// A further flame::draw_to_bins call on the flame of the last ref_host_run (fresh, zeroed bins of W x H).
long ref_host_draw(unsigned long W, unsigned long H, int draw_passes, char* out, long cap) {
    if (!g_flame) return -1;
    softgl::reset_logs();
    g_W = W; g_H = H;
    g_bins = std::make_unique<flame::bin_t>(W * H);
    g_bins->zero_out();
    std::size_t binned = g_flame->draw_to_bins(*g_bins, W, draw_passes);
    nlohmann::json j;
    j["binned"] = binned;
    nlohmann::json ids_in = nlohmann::json::array(), ids_out = nlohmann::json::array(), ss = nlohmann::json::array();
    for (auto& u : softgl::uniform_log()) {
        if (u.name == "ss_affine" && u.bytes.size() == 24) { ss = nlohmann::json::array(); for (int k = 0; k < 6; k++) { float v; std::memcpy(&v, u.bytes.data() + 4 * k, 4); ss.push_back(hex_float(v)); } }
        if (u.bytes.size() != 4) continue;
        unsigned v; std::memcpy(&v, u.bytes.data(), 4);
        if (u.name == "shuf_buf_idx_in") ids_in.push_back(v);
        if (u.name == "shuf_buf_idx_out") ids_out.push_back(v);
    }
    j["shuf_buf_idx_in"] = ids_in; j["shuf_buf_idx_out"] = ids_out; j["ss_affine"] = ss;
    return finish(j.dump(), out, cap);
}

// Copies one of the reference's GL buffers: "shuffle", "rand_states", "particles" (the buffer the next pass would read),
// "particles_other", "samples", "fp_inflated", "bins", "palette", "density" / "tonemapped" (after ref_host_post).
// Returns the byte count (call with cap = 0 to query it).
long ref_host_buffer(const char* which, void* out, long cap) {
    std::string w = which;
    unsigned name = 0;
    if (w == "shuffle" && flame::shuffle_buffers_) name = flame::shuffle_buffers_->name();
    else if (w == "rand_states" && flame::rand_states_) name = flame::rand_states_->name();
    else if (w == "samples" && flame::sample_buffer_) name = flame::sample_buffer_->name();
    else if (w == "particles" && g_flame && g_flame->local_buffer_) name = g_flame->local_buffer_->name();
    else if (w == "particles_other" && flame::swap_buffer_) name = flame::swap_buffer_->name();
    else if (w == "fp_inflated" && g_flame && g_flame->inflated_buffer_) name = g_flame->inflated_buffer_->name();
    else if (w == "bins" && g_bins) name = g_bins->name();
    else if (w == "palette" && g_flame) name = g_flame->palette_.name();
    else if (w == "density" || w == "tonemapped") {
        auto& v = w == "density" ? g_de : g_tm;
        long n = (long)(v.size() * 4);
        if (cap >= n && n) std::memcpy(out, v.data(), n);
        return n;
    } else return -1;
    long n = (long)softgl::buffer_size(name);
    if (cap >= n && n) std::memcpy(out, softgl::buffer_data(name), n);
    return n;
}

// Replaces the contents of the bins buffer (to post-process a histogram of the caller's choosing).
long ref_host_set_bins(const float* rgba, unsigned long W, unsigned long H) {
    g_W = W; g_H = H;
    g_bins = std::make_unique<flame::bin_t>(W * H);
    g_bins->update_all(reinterpret_cast<const std::array<float, 4>*>(rgba));
    return 0;
}

// src/main.cpp:490-535: density estimation as a GL_POINTS draw with additive blending, then the tonemap dispatch.
// estimator_* / gamma / brightness / vibrancy < 0 take the loaded flame's values (main.cpp reads flame_def->...).
long ref_host_post(int estimator_radius, int estimator_min, float estimator_curve, float gamma, float brightness, float vibrancy, double scale_constant) {
    if (!g_bins) return -1;
    static std::unique_ptr<vf_shader> density_vf;
    static std::unique_ptr<compute_shader> tonemap_cs;
    if (!density_vf) density_vf = std::make_unique<vf_shader>(read_file("shaders/density_vert.glsl"), read_file("shaders/density_frag.glsl"));  // main.cpp:209
    if (!tonemap_cs) tonemap_cs = std::make_unique<compute_shader>(read_file("shaders/tonemap.glsl"));                                          // main.cpp:208
    std::size_t target_dims[2] = {g_W, g_H};
    g_de.assign(g_W * g_H * 4, 0.0f);  // glClearColor(0,0,0,0); glClear
    g_tm.assign(g_W * g_H * 4, 0.0f);
    softgl::set_color_target(g_de.data(), (int)g_W, (int)g_H);
    glBindBufferBase(GL_SHADER_STORAGE_BUFFER, 8, g_bins->name());
    {
        // glm::ortho(0, W, H, 0, -1, 1), column-major
        glm::mat4 proj{};
        float l = 0.0f, r = float(target_dims[0]), b = float(target_dims[1]), t = 0.0f, n = -1.0f, f = 1.0f;
        proj.v[0] = 2.0f / (r - l); proj.v[5] = 2.0f / (t - b); proj.v[10] = -2.0f / (f - n);
        proj.v[12] = -(r + l) / (r - l); proj.v[13] = -(t + b) / (t - b); proj.v[14] = -(f + n) / (f - n); proj.v[15] = 1.0f;
        glUseProgram(density_vf->name());
        if (estimator_radius > 100) estimator_radius = 100;
        density_vf->set_uniform<int>("estimator_radius", estimator_radius);
        density_vf->set_uniform<int>("estimator_min", estimator_min);
        density_vf->set_uniform<float>("estimator_curve", estimator_curve);
        density_vf->set_uniform<int>("row_width", target_dims[0]);
        density_vf->set_uniform<float>("scale_constant", 1.0 / pow(10.0, scale_constant));
        density_vf->set_uniform<float>("gamma", gamma);
        density_vf->set_uniform<float>("brightness", brightness);
        density_vf->set_uniform<float>("vibrancy", vibrancy);
        density_vf->set_uniform<glm::mat4>("projection", proj);
        softgl::draw_points(0, (int)(target_dims[0] * target_dims[1]));  // glDrawArrays(GL_POINTS, 0, W*H)
    }
    {
        struct image { float* texels; int w, h; } in{g_de.data(), (int)g_W, (int)g_H}, outi{g_tm.data(), (int)g_W, (int)g_H};
        glUseProgram(tonemap_cs->name());
        softgl::set_uniform(glGetUniformLocation(tonemap_cs->name(), "image_in"), &in, sizeof in);    // glBindImageTexture(0, ...) + uniform 0
        softgl::set_uniform(glGetUniformLocation(tonemap_cs->name(), "image_out"), &outi, sizeof outi);
        tonemap_cs->set_uniform<float>("scale_constant", 1.0 / pow(10.0, scale_constant));
        tonemap_cs->set_uniform<float>("gamma", gamma);
        tonemap_cs->set_uniform<float>("brightness", brightness);
        tonemap_cs->set_uniform<float>("vibrancy", vibrancy);
        glDispatchCompute(target_dims[0] / 8, target_dims[1] / 8, 1);
    }
    return 0;
}
}
