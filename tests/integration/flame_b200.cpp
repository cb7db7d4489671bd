// The reference-side binding of INTEGRATION.md as a real translation unit: the member functions of refrakt's own
// `struct flame` (declared in /root/reference/src/flame.hpp, included from where it lies) implemented over the C ABI of
// librefrakt_b200.so instead of OpenGL. A refrakt maintainer would add this file as src/flame_b200.cpp in place of
// src/flame.cpp. tests/integration/Makefile compiles it against the reference's headers (with the stand-in glad / glm of
// oracle/stubs: the GL objects that `flame` still declares as members become inert) together with binding_demo.cpp;
// tests/test_integration_binding.py builds and runs the result. Nothing here is linked into the product.
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <regex>
#include <set>
#include <string>
#include <vector>

#include "refrakt_b200.h"

#include "util.hpp"
#include "flame.hpp"
#include "variation_table.hpp"

namespace b200 {
// state the reference keeps in GL objects; here: handles into the library, keyed by the reference's flame object
struct bound { rfk_flame* handle = nullptr; };
static rfk_compiler* compiler = nullptr;
static std::map<const flame*, bound> flames;
static float* bins_dev = nullptr;
static std::size_t bins_len = 0;

static void report(const char* what) { std::cout << what << ": " << rfk_last_error() << std::endl; }

// the caller's histogram lives on the device; `storage_buffer` (a GL buffer name) is only its stand-in on the host side
float* device_bins(std::size_t len) {
    if (len != bins_len) {
        if (bins_dev) rfk_device_free(bins_dev);
        bins_dev = static_cast<float*>(rfk_device_alloc(len * 16));
        bins_len = bins_dev ? len : 0;
    }
    return bins_dev;
}

// copies the public fields the UI edits (main.cpp:324-381) into the library's flame
static bool push_fields(flame& f, rfk_flame* h) {
    rfk_flame_info info;
    if (rfk_flame_get_info(h, &info) != RFK_OK) return false;
    info.size[0] = f.size[0]; info.size[1] = f.size[1];
    info.center[0] = f.center[0]; info.center[1] = f.center[1];
    info.scale = f.scale; info.rotate = f.rotate;
    info.estimator_min = f.estimator_min; info.estimator_radius = f.estimator_radius; info.estimator_curve = f.estimator_curve;
    info.gamma = f.gamma; info.vibrancy = f.vibrancy; info.brightness = f.brightness;
    if (rfk_flame_set_info(h, &info) != RFK_OK) return false;
    bool ok = true;
    f.for_each_xform([&](int idx, flame_xform& x) {
        rfk_xform_info xi;
        if (rfk_flame_get_xform(h, idx, &xi) != RFK_OK) { ok = false; return; }
        for (int a = 0; a < 6; a++) xi.affine[a] = x.affine[a];
        if (x.post && xi.has_post) for (int a = 0; a < 6; a++) xi.post[a] = (*x.post)[a];
        xi.weight = x.weight; xi.color = x.color; xi.color_speed = x.color_speed;
        xi.rotation_frequency = x.rotation_frequency; xi.opacity = x.opacity;
        ok = ok && rfk_flame_set_xform(h, idx, &xi) == RFK_OK;
        for (auto& [name, w] : x.variations) ok = ok && rfk_flame_set_variation(h, idx, name.c_str(), w) == RFK_OK;
        for (auto& [name, v] : x.var_param) ok = ok && rfk_flame_set_param(h, idx, name.c_str(), v) == RFK_OK;
    });
    return ok && rfk_flame_set_palette(h, &f.palette[0][0]) == RFK_OK;
}

// fills the reference's public fields from the library's parse of the genome
static void pull_fields(flame& f, rfk_flame* h) {
    rfk_flame_info info;
    rfk_flame_get_info(h, &info);
    f.size = {info.size[0], info.size[1]};
    f.center = {info.center[0], info.center[1]};
    f.scale = info.scale; f.rotate = info.rotate;
    f.estimator_min = info.estimator_min; f.estimator_radius = info.estimator_radius; f.estimator_curve = info.estimator_curve;
    f.gamma = info.gamma; f.vibrancy = info.vibrancy; f.brightness = info.brightness;
    auto read_xform = [&](int idx) {
        flame_xform x{};
        rfk_xform_info xi;
        rfk_flame_get_xform(h, idx, &xi);
        for (int a = 0; a < 6; a++) x.affine[a] = xi.affine[a];
        if (xi.has_post) { x.post = flame_xform::affine_t{}; for (int a = 0; a < 6; a++) (*x.post)[a] = xi.post[a]; }
        x.weight = xi.weight; x.color = xi.color; x.color_speed = xi.color_speed;
        x.rotation_frequency = xi.rotation_frequency; x.opacity = xi.opacity;
        for (int k = 0; k < xi.num_variations; k++) { const char* n = rfk_flame_variation_name(h, idx, k); float v = 0; rfk_flame_get_variation(h, idx, n, &v); x.variations[n] = v; }
        for (int k = 0; k < xi.num_params; k++) { const char* n = rfk_flame_param_name(h, idx, k); float v = 0; rfk_flame_get_param(h, idx, n, &v); x.var_param[n] = v; }
        return x;
    };
    for (int i = 0; i < info.num_xforms; i++) f.xforms.push_back(read_xform(i));
    if (info.has_final_xform) f.final_xform = read_xform(-1);
    rfk_flame_get_palette(h, &f.palette[0][0]);
}

rfk_flame* handle_of(const flame* f) { auto it = flames.find(f); return it == flames.end() ? nullptr : it->second.handle; }
}  // namespace b200

// ---- src/flame.cpp:105-158
void flame::set_sim_parameters(std::size_t total_particles, std::size_t temporal_samples, std::size_t shuffle_count) {
    num_temporal_samples_ = temporal_samples;
    num_shuffle_buffers_ = shuffle_count;
    if (rfk_set_sim_parameters(total_particles, temporal_samples, shuffle_count, /*seed=*/0) != RFK_OK) b200::report("set_sim_parameters");
    for (auto f : active_flames_) f->needs_update = true;  // the reference drops the flames' local buffers (:153-157)
}

// ---- src/flame.cpp:160-226: nullptr on an unknown attribute or a kernel that does not compile, message on stdout
std::unique_ptr<flame> flame::load_flame(const std::string& path, const flame_compiler&) {
    if (!b200::compiler) b200::compiler = rfk_compiler_create("variations.yaml");  // variation_table.cpp:183 reads the same file
    rfk_flame* h = b200::compiler ? rfk_flame_load(path.c_str(), b200::compiler) : nullptr;
    if (!h) { b200::report("load_flame"); return nullptr; }
    auto f = std::unique_ptr<flame>(new flame{});
    b200::pull_fields(*f, h);
    b200::flames[f.get()].handle = h;
    f->needs_update = true;
    return f;
}

// ---- src/flame.cpp:228-281
void flame::warmup(std::size_t num_passes, float tss_width) {
    rfk_flame* h = b200::handle_of(this);
    if (!h || !b200::push_fields(*this, h)) { b200::report("warmup (fields)"); return; }
    if (rfk_flame_warmup(h, num_passes, tss_width) != RFK_OK) { b200::report("warmup"); return; }
    needs_update = false;
    // needs_warmup() (flame.hpp:86) also tests the two GL buffers the reference allocates here; give it inert ones
    if (!local_buffer_) local_buffer_ = std::make_unique<pos_buffer_t>(1);
    if (!inflated_buffer_) inflated_buffer_ = std::make_unique<storage_buffer<float>>(1);
}

// ---- src/flame.cpp:283-330: accumulates into the device histogram that stands behind `bins`
std::size_t flame::draw_to_bins(bin_t& bins, std::size_t bins_width, int num_iter) {
    rfk_flame* h = b200::handle_of(this);
    float* dev = b200::device_bins(bins.size());
    if (!h || !dev) { b200::report("draw_to_bins"); return 0; }
    int64_t binned = rfk_flame_draw_to_bins(h, dev, bins.size(), bins_width, num_iter);
    if (binned < 0) { b200::report("draw_to_bins"); return 0; }
    return std::size_t(binned);
}

// ---- src/flame.cpp:332-336
void flame::reset_animation() {
    if (rfk_flame* h = b200::handle_of(this)) rfk_flame_reset_animation(h);
    needs_update = true;
}
