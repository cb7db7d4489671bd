// extern "C" surface of include/refrakt_b200.h over the C++ host classes.
#include "../../include/refrakt_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "flame.hpp"
#include "flame_device.hpp"
#include "png_writer.hpp"
#include "buffer_cache.hpp"
#include "comm.hpp"
#include "static_kernels.cuh"
#include "textutil.hpp"
#include "variation_table.hpp"

using namespace rfk;

namespace {

thread_local std::string t_error;
thread_local std::string t_scratch;

int fail(int code, const std::string& msg) {
    t_error = msg;
    return code;
}

template <typename F>
int guarded(F&& body) {
    try {
        return (int)body();
    } catch (const std::invalid_argument& e) {
        return fail(RFK_E_INVALID, e.what());
    } catch (const std::out_of_range& e) {
        return fail(RFK_E_NOTFOUND, e.what());
    } catch (const std::exception& e) {
        return fail(RFK_E_CUDA, e.what());
    }
}

flame* F(rfk_flame* f) { return reinterpret_cast<flame*>(f); }
const flame* F(const rfk_flame* f) { return reinterpret_cast<const flame*>(f); }
flame_compiler* C(rfk_compiler* c) { return reinterpret_cast<flame_compiler*>(c); }
const flame_compiler* C(const rfk_compiler* c) { return reinterpret_cast<const flame_compiler*>(c); }

void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

flame_xform* xform_at(flame* f, int index) {
    if (index == -1) return f->final_xform ? &f->final_xform.value() : nullptr;
    if (index < 0 || index >= (int)f->xforms.size()) return nullptr;
    return &f->xforms[index];
}
const flame_xform* xform_at(const flame* f, int index) { return xform_at(const_cast<flame*>(f), index); }

kernels::density_params to_density(const rfk_post_params& p, size_t W, size_t H) {
    kernels::density_params d{};
    d.W = (int)W;
    d.H = (int)H;
    d.estimator_radius = p.estimator_radius > 100 ? 100 : (p.estimator_radius < 0 ? 0 : p.estimator_radius);  // main.cpp:502
    d.estimator_min = p.estimator_min < 0 ? 0 : p.estimator_min;
    d.estimator_curve = p.estimator_curve;
    d.gamma = p.gamma;
    d.brightness = p.brightness;
    d.vibrancy = p.vibrancy;
    d.scale_constant = p.scale_constant;
    return d;
}

int run_post(const float* in, float* out_f4, uint8_t* out_rgba8, size_t W, size_t H, const rfk_post_params* p, bool density, bool tonemap) {
    if (!in || !p || W == 0 || H == 0 || W > 0x3fffffff || H > 0x3fffffff) throw std::invalid_argument("post: bad image arguments");
    if (!out_f4 && !out_rgba8) throw std::invalid_argument("post: no output buffer");
    auto d = to_density(*p, W, H);
    kernels::density_tonemap(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out_f4), reinterpret_cast<uchar4*>(out_rgba8), d, density, tonemap, current_stream());
    count_launch(1);
    cuda_ok(cudaGetLastError(), "density_tonemap launch");
    return RFK_OK;  // stream-ordered: rfk_memcpy_to_host / rfk_synchronize wait for it
}

}  // namespace

extern "C" {

int rfk_abi_version(void) { return RFK_ABI_VERSION; }
const char* rfk_last_error(void) { return t_error.c_str(); }

int rfk_set_device(int ordinal) {
    return guarded([&]() -> int {
        int before = -1;
        const bool had = cudaGetDevice(&before) == cudaSuccess;
        cuda_ok(cudaSetDevice(ordinal), "cudaSetDevice");
        // the buffers the library owns (simulation state, frame buffers) live on the device they were allocated on: a process
        // that moves to another device starts over there (set_sim_parameters + warmup again)
        if (had && before != ordinal) rfk_release_buffers();
        return RFK_OK;
    });
}
int rfk_set_stream(void* cuda_stream) { set_current_stream(reinterpret_cast<cudaStream_t>(cuda_stream)); return RFK_OK; }
int rfk_synchronize(void) {
    return guarded([&]() -> int { cuda_ok(cudaStreamSynchronize(current_stream()), "cudaStreamSynchronize"); return RFK_OK; });
}
void* rfk_device_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(RFK_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}
int rfk_device_free(void* p) {
    return guarded([&]() -> int { cuda_ok(cudaFree(p), "cudaFree"); return RFK_OK; });
}
int rfk_device_zero(void* p, size_t bytes) {
    return guarded([&]() -> int { cuda_ok(cudaMemsetAsync(p, 0, bytes, current_stream()), "cudaMemsetAsync"); return RFK_OK; });
}
int rfk_memcpy_to_device(void* dst, const void* src, size_t bytes) {
    return guarded([&]() -> int {
        cuda_ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, current_stream()), "cudaMemcpy H2D");
        cuda_ok(cudaStreamSynchronize(current_stream()), "cudaMemcpy H2D");
        return RFK_OK;
    });
}
int rfk_memcpy_to_host(void* dst, const void* src, size_t bytes) {
    return guarded([&]() -> int {
        cuda_ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, current_stream()), "cudaMemcpy D2H");
        cuda_ok(cudaStreamSynchronize(current_stream()), "cudaMemcpy D2H");
        return RFK_OK;
    });
}
uint64_t rfk_kernel_launch_count(void) { return kernel_launch_count(); }

int rfk_set_sim_parameters(size_t total_particles, size_t temporal_samples, size_t shuffle_count, uint64_t seed) {
    return guarded([&]() -> int { flame::set_sim_parameters(total_particles, temporal_samples, shuffle_count, seed); return RFK_OK; });
}

// ---- compiler ----
rfk_compiler* rfk_compiler_create(const char* path) {
    try {
        if (!path) throw std::invalid_argument("rfk_compiler_create: null path");
        return reinterpret_cast<rfk_compiler*>(new flame_compiler(path));
    } catch (const std::exception& e) { fail(RFK_E_INVALID, e.what()); return nullptr; }
}
rfk_compiler* rfk_compiler_create_from_text(const char* text) {
    try {
        if (!text) throw std::invalid_argument("rfk_compiler_create_from_text: null text");
        return reinterpret_cast<rfk_compiler*>(new flame_compiler(flame_compiler::from_text(text)));
    } catch (const std::exception& e) { fail(RFK_E_INVALID, e.what()); return nullptr; }
}
int rfk_compiler_load_overlay(rfk_compiler* c, const char* path) {
    return guarded([&]() -> int {
        if (!c || !path) throw std::invalid_argument("rfk_compiler_load_overlay: null argument");
        bool ok = false;
        std::string text = read_file(path, &ok);
        if (!ok) throw std::invalid_argument(std::string("cannot read ") + path);
        C(c)->load_overlay_text(text);
        return RFK_OK;
    });
}
void rfk_compiler_destroy(rfk_compiler* c) { delete C(c); }
int rfk_compiler_is_param(const rfk_compiler* c, const char* n) { return c && n && C(c)->is_param(n); }
int rfk_compiler_is_variation(const rfk_compiler* c, const char* n) { return c && n && C(c)->is_variation(n); }
int rfk_compiler_is_common(const rfk_compiler* c, const char* n) { return c && n && C(c)->is_common(n); }
int rfk_compiler_variation_count(const rfk_compiler* c) { return c ? (int)C(c)->variations().size() : 0; }
const char* rfk_compiler_variation_name(const rfk_compiler* c, int index) {
    if (!c || index < 0 || index >= (int)C(c)->variations().size()) return nullptr;
    auto it = C(c)->variations().begin();
    std::advance(it, index);
    return it->first.c_str();
}
int rfk_compiler_get_parameters_for_variation(const rfk_compiler* c, const char* name, char* buf, size_t buf_len) {
    return guarded([&]() -> int {
        if (!c || !name) throw std::invalid_argument("null argument");
        if (!C(c)->is_variation(name)) return fail(RFK_E_NOTFOUND, std::string("unknown variation ") + name);
        const auto& params = C(c)->get_parameters_for_variation(name);
        std::string joined;
        for (size_t i = 0; i < params.size(); i++) joined += (i ? "\n" : "") + params[i];
        if (buf && buf_len) {
            std::strncpy(buf, joined.c_str(), buf_len - 1);
            buf[buf_len - 1] = '\0';
        }
        return (int)params.size();
    });
}
const char* rfk_compiler_param_owner(const rfk_compiler* c, const char* param) {
    if (!c || !param || !C(c)->is_param(param)) return nullptr;
    t_scratch = C(c)->param_owner(param);
    return t_scratch.c_str();
}

// ---- flame ----
rfk_flame* rfk_flame_load(const char* path, const rfk_compiler* c) {
    if (!path || !c) { fail(RFK_E_INVALID, "rfk_flame_load: null argument"); return nullptr; }
    try {
        auto f = flame::load_flame(path, *C(c));
        if (!f) { fail(RFK_E_INVALID, flame::last_error()); return nullptr; }
        return reinterpret_cast<rfk_flame*>(f.release());
    } catch (const std::exception& e) { fail(RFK_E_INVALID, e.what()); return nullptr; }
}
rfk_flame* rfk_flame_load_string(const char* xml_text, const rfk_compiler* c) {
    if (!xml_text || !c) { fail(RFK_E_INVALID, "rfk_flame_load_string: null argument"); return nullptr; }
    try {
        auto f = flame::load_flame_string(xml_text, "<string>", *C(c));
        if (!f) { fail(RFK_E_INVALID, flame::last_error()); return nullptr; }
        return reinterpret_cast<rfk_flame*>(f.release());
    } catch (const std::exception& e) { fail(RFK_E_INVALID, e.what()); return nullptr; }
}
void rfk_flame_destroy(rfk_flame* f) { delete F(f); }

int rfk_flame_get_info(const rfk_flame* f, rfk_flame_info* o) {
    if (!f || !o) return fail(RFK_E_INVALID, "null argument");
    const flame* fl = F(f);
    o->size[0] = fl->size[0]; o->size[1] = fl->size[1];
    o->center[0] = fl->center[0]; o->center[1] = fl->center[1];
    o->scale = fl->scale; o->rotate = fl->rotate;
    o->estimator_min = fl->estimator_min; o->estimator_radius = fl->estimator_radius; o->estimator_curve = fl->estimator_curve;
    o->gamma = fl->gamma; o->vibrancy = fl->vibrancy; o->brightness = fl->brightness;
    o->num_xforms = (int)fl->xforms.size();
    o->has_final_xform = fl->final_xform ? 1 : 0;
    o->param_count = fl->buffer_map().size;
    return RFK_OK;
}
int rfk_flame_set_info(rfk_flame* f, const rfk_flame_info* i) {
    if (!f || !i) return fail(RFK_E_INVALID, "null argument");
    flame* fl = F(f);
    fl->size = {i->size[0], i->size[1]};
    fl->center = {i->center[0], i->center[1]};
    fl->scale = i->scale; fl->rotate = i->rotate;
    fl->estimator_min = i->estimator_min; fl->estimator_radius = i->estimator_radius; fl->estimator_curve = i->estimator_curve;
    fl->gamma = i->gamma; fl->vibrancy = i->vibrancy; fl->brightness = i->brightness;
    return RFK_OK;  // none of these feed the particle state: no re-warmup (main.cpp:324-333)
}
int rfk_flame_get_xform(const rfk_flame* f, int index, rfk_xform_info* o) {
    if (!f || !o) return fail(RFK_E_INVALID, "null argument");
    const flame_xform* x = xform_at(F(f), index);
    if (!x) return fail(RFK_E_NOTFOUND, "no xform " + std::to_string(index));
    for (int k = 0; k < 6; k++) { o->affine[k] = x->affine[k]; o->post[k] = x->post ? (*x->post)[k] : 0.0f; }
    o->has_post = x->post ? 1 : 0;
    o->weight = x->weight; o->color = x->color; o->color_speed = x->color_speed;
    o->rotation_frequency = x->rotation_frequency; o->opacity = x->opacity;
    o->num_variations = (int)x->variations.size();
    o->num_params = (int)x->var_param.size();
    return RFK_OK;
}
int rfk_flame_set_xform(rfk_flame* f, int index, const rfk_xform_info* i) {
    if (!f || !i) return fail(RFK_E_INVALID, "null argument");
    flame_xform* x = xform_at(F(f), index);
    if (!x) return fail(RFK_E_NOTFOUND, "no xform " + std::to_string(index));
    if ((i->has_post != 0) != x->post.has_value()) return fail(RFK_E_INVALID, "has_post cannot change after load (the generated kernel depends on it)");
    for (int k = 0; k < 6; k++) x->affine[k] = i->affine[k];
    if (x->post) for (int k = 0; k < 6; k++) (*x->post)[k] = i->post[k];
    x->weight = i->weight; x->color = i->color; x->color_speed = i->color_speed;
    x->rotation_frequency = i->rotation_frequency; x->opacity = i->opacity;
    F(f)->mark_dirty();
    return RFK_OK;
}
static const char* nth_key(const std::map<std::string, float>& m, int k) {
    if (k < 0 || k >= (int)m.size()) return nullptr;
    auto it = m.begin();
    std::advance(it, k);
    return it->first.c_str();
}
const char* rfk_flame_variation_name(const rfk_flame* f, int xform, int k) {
    const flame_xform* x = f ? xform_at(F(f), xform) : nullptr;
    return x ? nth_key(x->variations, k) : nullptr;
}
const char* rfk_flame_param_name(const rfk_flame* f, int xform, int k) {
    const flame_xform* x = f ? xform_at(F(f), xform) : nullptr;
    return x ? nth_key(x->var_param, k) : nullptr;
}
static int map_get(const std::map<std::string, float>& m, const char* name, float* out) {
    auto it = m.find(name);
    if (it == m.end()) return fail(RFK_E_NOTFOUND, std::string("no entry named ") + name);
    *out = it->second;
    return RFK_OK;
}
static int map_set(std::map<std::string, float>& m, const char* name, float v) {
    auto it = m.find(name);
    if (it == m.end()) return fail(RFK_E_NOTFOUND, std::string("no entry named ") + name + " (the structure of a flame is fixed after load)");
    it->second = v;
    return RFK_OK;
}
int rfk_flame_get_variation(const rfk_flame* f, int xform, const char* name, float* out) {
    const flame_xform* x = (f && name && out) ? xform_at(F(f), xform) : nullptr;
    return x ? map_get(x->variations, name, out) : fail(RFK_E_NOTFOUND, "no such xform");
}
int rfk_flame_set_variation(rfk_flame* f, int xform, const char* name, float v) {
    flame_xform* x = (f && name) ? xform_at(F(f), xform) : nullptr;
    if (!x) return fail(RFK_E_NOTFOUND, "no such xform");
    int r = map_set(x->variations, name, v);
    if (r == RFK_OK) F(f)->mark_dirty();
    return r;
}
int rfk_flame_get_param(const rfk_flame* f, int xform, const char* name, float* out) {
    const flame_xform* x = (f && name && out) ? xform_at(F(f), xform) : nullptr;
    return x ? map_get(x->var_param, name, out) : fail(RFK_E_NOTFOUND, "no such xform");
}
int rfk_flame_set_param(rfk_flame* f, int xform, const char* name, float v) {
    flame_xform* x = (f && name) ? xform_at(F(f), xform) : nullptr;
    if (!x) return fail(RFK_E_NOTFOUND, "no such xform");
    int r = map_set(x->var_param, name, v);
    if (r == RFK_OK) F(f)->mark_dirty();
    return r;
}
int rfk_flame_get_palette(const rfk_flame* f, float* rgba) {
    if (!f || !rgba) return fail(RFK_E_INVALID, "null argument");
    std::memcpy(rgba, F(f)->palette.data(), 256 * 4 * sizeof(float));
    return RFK_OK;
}
int rfk_flame_set_palette(rfk_flame* f, const float* rgba) {
    if (!f || !rgba) return fail(RFK_E_INVALID, "null argument");
    std::memcpy(F(f)->palette.data(), rgba, 256 * 4 * sizeof(float));
    F(f)->mark_dirty();
    return RFK_OK;
}

const char* rfk_flame_buffer_map_json(const rfk_flame* f) {
    if (!f) return nullptr;
    t_scratch = F(f)->buffer_map().dump_json();
    return t_scratch.c_str();
}
int rfk_flame_copy_params(const rfk_flame* f, float* out) {
    if (!f || !out) return fail(RFK_E_INVALID, "null argument");
    auto buf = F(f)->copy_flame_data_to_buffer();
    std::memcpy(out, buf.data(), sizeof(float) * flame::PARAM_BUFFER);
    return RFK_OK;
}
const char* rfk_flame_glsl_source(const rfk_flame* f) { return f ? F(f)->glsl_source().c_str() : nullptr; }
const char* rfk_flame_cuda_source(const rfk_flame* f) { return f ? F(f)->cuda_source().c_str() : nullptr; }
int rfk_flame_get_cubin(rfk_flame* f, void* buf, size_t buf_len, size_t* size) {
    return guarded([&]() -> int {
        if (!f || !size) throw std::invalid_argument("null argument");
        const auto& image = F(f)->cubin();
        *size = image.size();
        if (buf) {
            if (buf_len < image.size()) throw std::invalid_argument("cubin buffer too small");
            std::memcpy(buf, image.data(), image.size());
        }
        return RFK_OK;
    });
}

int rfk_flame_get_variant_cubin(rfk_flame* f, int staged, int specialised, void* buf, size_t buf_len, size_t* size) {
    return guarded([&]() -> int {
        if (!f || !size) throw std::invalid_argument("null argument");
        flame* fl = F(f);
        std::vector<float> table;
        if (specialised) {
            auto fp = fl->copy_flame_data_to_buffer();
            table = fl->constant_table(fp.data());
        }
        const auto image = fl->variant_cubin(staged != 0, specialised ? &table : nullptr, specialised == 2);
        *size = image.size();
        if (buf) {
            if (buf_len < image.size()) throw std::invalid_argument("cubin buffer too small");
            std::memcpy(buf, image.data(), image.size());
        }
        return RFK_OK;
    });
}
const char* rfk_flame_variant_source(rfk_flame* f, int staged, int specialised) {
    if (!f) return nullptr;
    try {
        flame* fl = F(f);
        std::vector<float> table;
        if (specialised) {
            auto fp = fl->copy_flame_data_to_buffer();
            table = fl->constant_table(fp.data());
        }
        t_scratch = fl->variant_source(staged != 0, specialised ? &table : nullptr, specialised == 2);
        return t_scratch.c_str();
    } catch (const std::exception& e) { fail(RFK_E_INVALID, e.what()); return nullptr; }
}

int rfk_flame_pair_particles_state(const rfk_flame* f, float probe_ms_out[2]) {
    if (!f) return fail(RFK_E_INVALID, "null flame");
    return flame_pairs_state(*F(f), probe_ms_out);
}
int rfk_flame_uses_specialised(const rfk_flame* f) { return f ? (flame_uses_baked(*F(f)) ? 1 : 0) : fail(RFK_E_INVALID, "null flame"); }

int rfk_flame_get_options(const rfk_flame* f, rfk_kernel_options* o) {
    if (!f || !o) return fail(RFK_E_INVALID, "null argument");
    const auto& k = F(f)->options();
    *o = rfk_kernel_options{k.math_mode, k.fmad, k.per_lane_xform, k.warp_aggregate, k.deterministic, k.count_xforms, k.min_blocks, k.block_width, k.deal_period, k.l2_hints, k.staged_bins, k.specialize, k.pair_particles};
    return RFK_OK;
}
int rfk_flame_set_options(rfk_flame* f, const rfk_kernel_options* i) {
    if (!f || !i) return fail(RFK_E_INVALID, "null argument");
    kernel_options k;
    k.math_mode = i->math_mode; k.fmad = i->fmad != 0; k.per_lane_xform = i->per_lane_xform != 0;
    k.warp_aggregate = i->warp_aggregate != 0; k.deterministic = i->deterministic != 0; k.count_xforms = i->count_xforms != 0;
    k.min_blocks = i->min_blocks;
    k.block_width = i->block_width;
    k.deal_period = i->deal_period;
    k.l2_hints = i->l2_hints ? 1 : 0;
    k.staged_bins = i->staged_bins;
    k.specialize = i->specialize;
    k.pair_particles = i->pair_particles;
    if (k.pair_particles < 0 || k.pair_particles > 2) return fail(RFK_E_INVALID, "pair_particles must be 0 (never), 1 (measured) or 2 (always)");
    if (k.specialize < 0 || k.specialize > 2) return fail(RFK_E_INVALID, "specialize must be 0 (off), 1 (always) or 2 (automatic)");
    if (k.staged_bins != 0 && k.staged_bins != -1 && (k.staged_bins < 8 || k.staged_bins > 24))
        return fail(RFK_E_INVALID, "staged_bins must be -1 (automatic), 0 (off) or the log2 of the bins per region, 8 to 24");
    if (k.staged_bins > 0 && (k.deterministic || k.warp_aggregate || k.l2_hints)) return fail(RFK_E_INVALID, "staged_bins excludes deterministic, warp_aggregate and l2_hints");
    if (k.block_width != 128 && k.block_width != 256 && k.block_width != 512) return fail(RFK_E_INVALID, "block_width must be 128, 256 or 512");
    if (k.deal_period < 1) return fail(RFK_E_INVALID, "deal_period must be >= 1");
    if (k.min_blocks < -1 || k.min_blocks * k.block_width > 2048) return fail(RFK_E_INVALID, "min_blocks must be -1, 0 or at most 2048 / block_width");
    if (k.math_mode < 0 || k.math_mode > 2) return fail(RFK_E_INVALID, "math_mode must be 0, 1 or 2");
    if (k.count_xforms && F(f)->xforms.size() > 62) return fail(RFK_E_INVALID, "count_xforms supports at most 62 xforms");
    if (!F(f)->set_options(k)) return fail(RFK_E_CUDA, flame::last_error());
    return RFK_OK;
}
int rfk_flame_kernel_info(rfk_flame* f, const char* kernel, int* regs, int* smem, int* blocks) {
    return guarded([&]() -> int {
        if (!f || !kernel || !regs || !smem || !blocks) throw std::invalid_argument("null argument");
        flame_kernel_info(*F(f), kernel, regs, smem, blocks);
        return RFK_OK;
    });
}

int rfk_flame_needs_warmup(const rfk_flame* f) { return f ? (F(f)->needs_warmup() ? 1 : 0) : fail(RFK_E_INVALID, "null flame"); }
int rfk_flame_warmup(rfk_flame* f, size_t num_passes, float tss_width) {
    return guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        if (num_passes > 0x7fffffff) throw std::invalid_argument("too many passes");
        F(f)->warmup(num_passes, tss_width);
        return RFK_OK;
    });
}
int64_t rfk_flame_draw_to_bins(rfk_flame* f, float* bins, size_t bins_len, size_t bins_width, int num_iter) {
    int64_t result = 0;
    int rc = guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        if (num_iter < 0) throw std::invalid_argument("negative iteration count");
        if (F(f)->needs_warmup()) return fail(RFK_E_STATE, "draw_to_bins: warmup() has not been run for the current parameters");
        result = (int64_t)F(f)->draw_to_bins(bins, bins_len, bins_width, num_iter);
        return RFK_OK;
    });
    return rc == RFK_OK ? result : rc;
}
int rfk_flame_draw_to_bins_async(rfk_flame* f, float* bins, size_t bins_len, size_t bins_width, int num_iter) {
    return guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        if (num_iter < 0) throw std::invalid_argument("negative iteration count");
        if (F(f)->needs_warmup()) return fail(RFK_E_STATE, "draw_to_bins: warmup() has not been run for the current parameters");
        F(f)->draw_to_bins_async(bins, bins_len, bins_width, num_iter);
        return RFK_OK;
    });
}
int rfk_flame_build_hot_map(rfk_flame* f, const float* bins, size_t bins_len, size_t bins_width, uint64_t budget_bytes, rfk_hot_map_info* info_out) {
    return guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        auto i = F(f)->build_hot_map(bins, bins_len, bins_width, budget_bytes);
        if (info_out) *info_out = rfk_hot_map_info{i.tiles_x, i.tiles_y, i.hot_tiles, i.threshold_bucket, i.budget_bytes};
        return RFK_OK;
    });
}
int rfk_flame_clear_hot_map(rfk_flame* f) {
    return guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        F(f)->clear_hot_map();
        return RFK_OK;
    });
}
int64_t rfk_flame_copy_hot_map(rfk_flame* f, uint32_t* out, size_t n_words) {
    int64_t result = 0;
    int rc = guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        auto m = F(f)->copy_hot_map();
        if (out) std::memcpy(out, m.data(), std::min(n_words, m.size()) * sizeof(uint32_t));
        result = (int64_t)m.size();
        return RFK_OK;
    });
    return rc == RFK_OK ? result : rc;
}
int64_t rfk_flame_binned_total(rfk_flame* f) {
    int64_t result = 0;
    int rc = guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        result = (int64_t)F(f)->binned_total();
        return RFK_OK;
    });
    return rc == RFK_OK ? result : rc;
}
int rfk_flame_reset_animation(rfk_flame* f) {
    if (!f) return fail(RFK_E_INVALID, "null flame");
    F(f)->reset_animation();
    return RFK_OK;
}
int rfk_flame_xform_counts(rfk_flame* f, uint64_t* out, int n) {
    return guarded([&]() -> int {
        if (!f || !out || n < 0 || n > 62) throw std::invalid_argument("bad argument");
        unsigned long long raw[64] = {0};
        flame_read_counters(*F(f), raw, 64);
        for (int i = 0; i < n; i++) out[i] = raw[1 + i];
        return RFK_OK;
    });
}
int rfk_flame_screen_affine(const rfk_flame* f, size_t W, size_t H, float out[6]) {
    if (!f || !out) return fail(RFK_E_INVALID, "null argument");
    auto a = F(f)->screen_space_affine(W, H);
    for (int i = 0; i < 6; i++) out[i] = a[i];
    return RFK_OK;
}
int rfk_flame_rotate_xforms(rfk_flame* f, float degrees) {
    if (!f) return fail(RFK_E_INVALID, "null flame");
    flame* fl = F(f);
    for (auto& x : fl->xforms)
        if (x.rotation_frequency != 0.0f) x.affine = flame::rotate_affine(x.affine, degrees * x.rotation_frequency);
    if (fl->final_xform) {
        auto& fx = fl->final_xform.value();
        if (fx.rotation_frequency != 0.0f) fx.affine = flame::rotate_affine(fx.affine, degrees * fx.rotation_frequency);
    }
    fl->mark_dirty();
    return RFK_OK;
}

int rfk_flame_motion_count(const rfk_flame* f, int xform) {
    const flame_xform* x = f ? xform_at(F(f), xform) : nullptr;
    return x ? (int)x->motion.size() : fail(RFK_E_NOTFOUND, "no such xform");
}
int rfk_flame_get_motion(const rfk_flame* f, int xform, int k, rfk_motion_info* out) {
    const flame_xform* x = (f && out) ? xform_at(F(f), xform) : nullptr;
    if (!x) return fail(RFK_E_NOTFOUND, "no such xform");
    if (k < 0 || k >= (int)x->motion.size()) return fail(RFK_E_NOTFOUND, "no such motion entry");
    auto it = x->motion.begin();
    std::advance(it, k);
    std::memset(out, 0, sizeof *out);
    out->freq = it->second.freq;
    out->amplitude = it->second.amplitude;
    std::strncpy(out->function, it->second.function.c_str(), sizeof out->function - 1);
    std::strncpy(out->target, it->first.c_str(), sizeof out->target - 1);
    return RFK_OK;
}
int rfk_flame_apply_motion(rfk_flame* f, float time_seconds) {
    if (!f) return fail(RFK_E_INVALID, "null flame");
    return F(f)->apply_motion(time_seconds);
}
float rfk_motion_function(const char* name, float x) { return name ? flame::motion_function(name, x) : 0.0f; }

void rfk_rotate_affine(const float a[6], float deg, float out[6]) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::rotate_affine(in, deg);
    for (int i = 0; i < 6; i++) out[i] = r[i];
}
void rfk_scale_affine(const float a[6], float scale, float out[6]) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::scale_affine(in, scale);
    for (int i = 0; i < 6; i++) out[i] = r[i];
}
void rfk_translate_affine(const float a[6], const float t[2], float out[6]) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::translate_affine(in, {t[0], t[1]});
    for (int i = 0; i < 6; i++) out[i] = r[i];
}

// ---- post ----
int rfk_flame_post_params(const rfk_flame* f, rfk_post_params* o) {
    if (!f || !o) return fail(RFK_E_INVALID, "null argument");
    const flame* fl = F(f);
    o->estimator_radius = fl->estimator_radius > 100 ? 100 : fl->estimator_radius;
    o->estimator_min = fl->estimator_min;
    o->estimator_curve = fl->estimator_curve;
    o->gamma = fl->gamma; o->brightness = fl->brightness; o->vibrancy = fl->vibrancy;
    o->scale_constant = (float)(1.0 / std::pow(10.0, 4.0));  // main.cpp:228, :528
    return RFK_OK;
}
int rfk_density_estimate(const float* bins, float* image, size_t W, size_t H, const rfk_post_params* p) {
    return guarded([&]() -> int { return run_post(bins, image, nullptr, W, H, p, true, false); });
}
int rfk_tonemap(const float* in, float* out, uint8_t* rgba8, size_t W, size_t H, const rfk_post_params* p) {
    return guarded([&]() -> int { return run_post(in, out, rgba8, W, H, p, false, true); });
}
int rfk_density_tonemap(const float* bins, float* out, uint8_t* rgba8, size_t W, size_t H, const rfk_post_params* p) {
    return guarded([&]() -> int { return run_post(bins, out, rgba8, W, H, p, true, true); });
}
int rfk_density_tonemap_rows(const float* bins_rows, float* out, uint8_t* rgba8, size_t W, size_t H, const rfk_post_params* p,
                             uint32_t y0, uint32_t y1, uint32_t src_y0, uint32_t src_y1, uint32_t out_y0) {
    return guarded([&]() -> int {
        if (!bins_rows || !p || W == 0 || H == 0 || W > 0x3fffffff || H > 0x3fffffff) throw std::invalid_argument("post: bad image arguments");
        if (!out && !rgba8) throw std::invalid_argument("post: no output buffer");
        auto d = to_density(*p, W, H);
        const uint32_t R = (uint32_t)std::max(d.estimator_radius, d.estimator_min);
        if (y0 >= y1 || y1 > H || src_y0 >= src_y1 || src_y1 > H || out_y0 > y0) throw std::invalid_argument("rfk_density_tonemap_rows: bad row range");
        if (src_y0 > (y0 > R ? y0 - R : 0) || src_y1 < std::min<uint32_t>((uint32_t)H, y1 + R)) throw std::invalid_argument("rfk_density_tonemap_rows: the slab does not cover the estimator radius around the output rows");
        d.y0 = (int)y0; d.y1 = (int)y1; d.src_y0 = (int)src_y0; d.src_y1 = (int)src_y1; d.out_y0 = (int)out_y0;
        kernels::density_tonemap(reinterpret_cast<const float4*>(bins_rows), reinterpret_cast<float4*>(out), reinterpret_cast<uchar4*>(rgba8), d, true, true, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "density_tonemap launch");
        return RFK_OK;
    });
}
int rfk_downsample2x(const float* in, float* out, size_t W, size_t H) {
    return guarded([&]() -> int {
        if (!in || !out || !W || !H) throw std::invalid_argument("bad argument");
        kernels::downsample2x(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), (int)W, (int)H, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "downsample2x");
        return RFK_OK;
    });
}

int rfk_spatial_downsample(const float* in, float* out, size_t W, size_t H, int ss, float filter_radius) {
    return guarded([&]() -> int {
        if (!in || !out || !W || !H || ss < 1 || ss > 16 || !(filter_radius >= 0.0f)) throw std::invalid_argument("rfk_spatial_downsample: bad argument");
        kernels::spatial_downsample(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), (int)W, (int)H, ss, filter_radius, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "spatial_downsample");
        return RFK_OK;
    });
}
int rfk_spatial_filter_taps(int ss, float filter_radius, float* taps) {
    if (ss < 1 || ss > 16 || !(filter_radius >= 0.0f)) return fail(RFK_E_INVALID, "rfk_spatial_filter_taps: bad argument");
    return kernels::spatial_filter_taps(ss, filter_radius, taps);
}

// ---- seeding ----
int rfk_seed_rng_states(uint32_t* states, size_t count, uint32_t seed_base) {
    return guarded([&]() -> int {
        if (!states) throw std::invalid_argument("null argument");
        kernels::seed_rng_states(reinterpret_cast<uint4*>(states), count, seed_base, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "seed_rng_states");
        return RFK_OK;
    });
}
int rfk_make_sample_points(float* points, uint32_t count) {
    return guarded([&]() -> int {
        if (!points) throw std::invalid_argument("null argument");
        kernels::make_sample_points(reinterpret_cast<float4*>(points), count, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "make_sample_points");
        return RFK_OK;
    });
}
int rfk_make_shuffle_buffers(uint32_t* out, uint32_t size, uint32_t count, uint64_t seed) {
    return guarded([&]() -> int {
        if (!out) throw std::invalid_argument("null argument");
        kernels::make_shuffle_buffers(out, size, count, seed, current_stream());
        count_launch(1);
        cuda_ok(cudaGetLastError(), "make_shuffle_buffers");
        return RFK_OK;
    });
}
int rfk_copy_rng_states(uint32_t* out, size_t first, size_t count) {
    return guarded([&]() -> int {
        if (!out) throw std::invalid_argument("null argument");
        if (!sim_rng_states() || first + count > sim_total_particles()) throw std::invalid_argument("range outside the particle RNG states");
        cuda_ok(cudaMemcpyAsync(out, sim_rng_states() + first, count * sizeof(uint4), cudaMemcpyDeviceToHost, current_stream()), "copy rng states");
        cuda_ok(cudaStreamSynchronize(current_stream()), "copy rng states");
        return RFK_OK;
    });
}

// ---- end to end ----
namespace {
// library-owned frame buffers of rfk_render_frame: kept between calls and regrown only when the frame gets larger
struct frame_buffers {
    float4* bins = nullptr; float4* image = nullptr; uchar4* rgba8 = nullptr; float4* small = nullptr;
    size_t bins_n = 0, image_n = 0, rgba8_n = 0, small_n = 0;
    cudaEvent_t ev[5] = {};
    unsigned long long* host_counter = nullptr;  // pinned: the binned counter of a frame read back together with its image
    void release() {
        cudaFree(bins); cudaFree(image); cudaFree(rgba8); cudaFree(small);
        bins = image = small = nullptr; rgba8 = nullptr;
        bins_n = image_n = rgba8_n = small_n = 0;
    }
};
frame_buffers g_frame;

// device buffers of rfk_render_frame_sharded; every rank (re)allocates them in lockstep (the sizes follow from the request)
struct sharded_buffers {
    float4* bins = nullptr; float4* slab = nullptr; float4* image = nullptr; float4* small = nullptr; uchar4* rows8 = nullptr;
    uchar4* full8 = nullptr; float4* full_image = nullptr;
    size_t bins_n = 0, slab_n = 0, image_n = 0, small_n = 0, rows8_n = 0, full8_n = 0, full_image_n = 0;
    unsigned long long* counters = nullptr;  // [0] binned samples of all ranks, [1] barrier token
    uint64_t generation = 0;
    cudaEvent_t ev[6] = {};
    cudaEvent_t counted = nullptr;                // the first call's count has reached host_counter
    unsigned long long* host_counter = nullptr;   // pinned
    void release() {
        comm::release_peers();
        cudaFree(bins); cudaFree(slab); cudaFree(image); cudaFree(small); cudaFree(rows8); cudaFree(full8); cudaFree(full_image); cudaFree(counters);
        if (host_counter) cudaFreeHost(host_counter);
        host_counter = nullptr;
        bins = slab = image = small = full_image = nullptr; rows8 = full8 = nullptr; counters = nullptr;
        bins_n = slab_n = image_n = small_n = rows8_n = full8_n = full_image_n = 0;
        generation++;
    }
};
sharded_buffers g_sharded;
}  // namespace

int rfk_release_buffers(void) {
    return guarded([&]() -> int {
        g_frame.release();
        g_sharded.release();
        release_sim_buffers();
        return RFK_OK;
    });
}

int rfk_render_frame(rfk_flame* f, const rfk_frame_request* req, uint8_t* rgba8_out, float* image_out, rfk_frame_stats* stats) {
    return guarded([&]() -> int {
        if (!f || !req || (!rgba8_out && !image_out)) throw std::invalid_argument("rfk_render_frame: null argument");
        if (!req->width || !req->height) throw std::invalid_argument("rfk_render_frame: empty image");
        if (!req->target_binned && !req->max_draw_calls) throw std::invalid_argument("rfk_render_frame: neither target_binned nor max_draw_calls given");
        flame* fl = F(f);
        const size_t ss = req->supersample > 1 ? req->supersample : 1;
        if (ss > 16) throw std::invalid_argument("rfk_render_frame: supersample > 16");
        const size_t OW = req->width, OH = req->height, on = OW * OH;  // output image
        const size_t W = OW * ss, H = OH * ss, n = W * H;              // histogram (and full-resolution image)
        cudaStream_t s = current_stream();

        frame_buffers& b = g_frame;
        auto grow = [](auto*& ptr, size_t& have, size_t want, const char* what) {
            if (have >= want) return;
            cudaFree(ptr); ptr = nullptr; have = 0;
            cuda_ok(cudaMalloc(&ptr, want * sizeof(*ptr)), what);
            have = want;
        };
        grow(b.bins, b.bins_n, n, "cudaMalloc(bins)");
        if (image_out || ss > 1) grow(b.image, b.image_n, n, "cudaMalloc(image)");
        if (rgba8_out) grow(b.rgba8, b.rgba8_n, on, "cudaMalloc(rgba8)");
        if (ss > 1) grow(b.small, b.small_n, on, "cudaMalloc(filtered image)");
        for (auto& e : b.ev) if (!e) cuda_ok(cudaEventCreate(&e), "cudaEventCreate");

        cuda_ok(cudaEventRecord(b.ev[0], s), "event");
        fl->warmup(req->warmup_passes, req->tss_width);
        cuda_ok(cudaMemsetAsync(b.bins, 0, n * sizeof(float4), s), "clear bins");  // bins.zero_out(), main.cpp:406
        cuda_ok(cudaEventRecord(b.ev[1], s), "event");

        uint64_t binned = 0, iterations = 0;
        uint32_t calls = 0;
        // kernel option l2_hints: a histogram more than twice the L2 gets its hot map after the first draw call
        bool want_hot_map = false;
        fl->clear_hot_map();
        if (fl->options().l2_hints) {
            int dev = 0, l2 = 0;
            cuda_ok(cudaGetDevice(&dev), "cudaGetDevice");
            cuda_ok(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev), "query L2 size");
            want_hot_map = n * sizeof(float4) > 2 * (size_t)l2;
        }
        // The reference reads the binned counter back after every draw_to_bins call (flame.cpp:329) and stops when
        // `accumulated >= quality * W * H` (main.cpp:411). Same stopping rule, same whole calls of `drawing_passes` passes, but
        // the read-back blocks only where the answer is needed: after the first call (it gives the samples a call lands), then
        // once per batch of calls enqueued back to back, sized to stop just short of the target so that the count of calls is
        // the one the call-by-call loop would make.
        auto more = [&]() { return (req->max_draw_calls == 0 || calls < req->max_draw_calls) && (req->target_binned == 0 || binned < req->target_binned); };
        uint64_t per_call = 0, first_call = 0;
        const bool count_at_the_end = !req->target_binned && !want_hot_map;  // a fixed number of calls: no read-back in between
        if (count_at_the_end) {
            for (uint32_t k = 0; k < req->max_draw_calls; k++) fl->draw_to_bins_async(reinterpret_cast<float*>(b.bins), n, W, (int)req->drawing_passes);
            calls = req->max_draw_calls;
            iterations = (uint64_t)calls * req->drawing_passes * sim_total_particles();
        }
        while (!count_at_the_end && more()) {
            uint32_t batch = 1;
            if (per_call && req->target_binned) {
                const uint64_t left = req->target_binned - binned;
                const uint64_t safe = (uint64_t)((double)left / ((double)per_call * 1.01));  // calls that cannot reach the target yet
                batch = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(safe, 1), 1u << 20);
            } else if (!req->target_binned) {
                batch = 1u << 20;  // a fixed number of calls: all of them at once, the counter is read once at the end
            }
            if (req->max_draw_calls) batch = std::min(batch, req->max_draw_calls - calls);
            for (uint32_t k = 0; k < batch; k++) {
                fl->draw_to_bins_async(reinterpret_cast<float*>(b.bins), n, W, (int)req->drawing_passes);
                if (want_hot_map && calls + k == 0) fl->build_hot_map(reinterpret_cast<const float*>(b.bins), n, W, 0);
            }
            const uint64_t total = fl->binned_total();  // blocks, like counters_.get_one(0)
            per_call = (total - binned) / batch;
            binned = total;
            calls += batch;
            iterations += (uint64_t)batch * req->drawing_passes * sim_total_particles();
            if (req->target_binned && binned == 0 && calls >= 4) throw std::runtime_error("rfk_render_frame: nothing lands in the histogram");
            // a genome whose particles overflow one after the other (the reference never resets them, flame.glsl:70) stops
            // landing samples: the target would never be reached
            if (first_call == 0) first_call = per_call;
            if (req->target_binned && binned < req->target_binned && per_call * 1000 < first_call)
                throw std::runtime_error("rfk_render_frame: the samples stopped landing before the target was reached (the genome's particles overflow and are never reset, as in the reference); render with max_draw_calls instead");
        }
        cuda_ok(cudaEventRecord(b.ev[2], s), "event");

        rfk_post_params pp;
        rfk_flame_post_params(f, &pp);
        pp.scale_constant = (float)(1.0 / std::pow(10.0, (double)req->scale_constant_exp));
        auto d = to_density(pp, W, H);
        const float4* final_image = b.image;
        if (ss == 1) {
            kernels::density_tonemap(b.bins, image_out ? b.image : nullptr, rgba8_out ? b.rgba8 : nullptr, d, true, true, s);
            count_launch(1);
        } else {
            kernels::density_tonemap(b.bins, b.image, nullptr, d, true, true, s);
            kernels::spatial_downsample(b.image, b.small, (int)OW, (int)OH, (int)ss, req->filter_radius, s);
            count_launch(2);
            if (rgba8_out) { kernels::pack_rgba8(b.small, b.rgba8, on, s); count_launch(1); }
            final_image = b.small;
        }
        cuda_ok(cudaGetLastError(), "density_tonemap launch");
        cuda_ok(cudaEventRecord(b.ev[3], s), "event");
        if (rgba8_out) cuda_ok(cudaMemcpyAsync(rgba8_out, b.rgba8, on * sizeof(uchar4), cudaMemcpyDeviceToHost, s), "read back rgba8");
        if (image_out) cuda_ok(cudaMemcpyAsync(image_out, final_image, on * sizeof(float4), cudaMemcpyDeviceToHost, s), "read back image");
        if (count_at_the_end) {
            if (!b.host_counter) cuda_ok(cudaMallocHost(&b.host_counter, sizeof(unsigned long long)), "cudaMallocHost(counter)");
            cuda_ok(cudaMemcpyAsync(b.host_counter, flame_binned_counter_dev(*fl), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "read binned counter");
        }
        cuda_ok(cudaEventRecord(b.ev[4], s), "event");
        cuda_ok(cudaStreamSynchronize(s), "rfk_render_frame");
        if (count_at_the_end) binned = *b.host_counter;
        if (stats) {
            stats->iterations = iterations; stats->binned = binned; stats->draw_calls = calls;
            cudaEventElapsedTime(&stats->ms_warmup, b.ev[0], b.ev[1]);
            cudaEventElapsedTime(&stats->ms_draw, b.ev[1], b.ev[2]);
            cudaEventElapsedTime(&stats->ms_post, b.ev[2], b.ev[3]);
            cudaEventElapsedTime(&stats->ms_readback, b.ev[3], b.ev[4]);
        }
        return RFK_OK;
    });
}

// ---- several GPUs: one process per GPU (csrc/comm.cpp) ----

int rfk_comm_unique_id(uint8_t id_out[RFK_COMM_ID_BYTES]) {
    return guarded([&]() -> int {
        if (!id_out) throw std::invalid_argument("rfk_comm_unique_id: null argument");
        comm::unique_id(id_out);
        return RFK_OK;
    });
}
int rfk_comm_init(const uint8_t id[RFK_COMM_ID_BYTES], int rank, int world) {
    return guarded([&]() -> int {
        if (!id) throw std::invalid_argument("rfk_comm_init: null id");
        g_sharded.release();
        comm::init(id, rank, world);
        return RFK_OK;
    });
}
int rfk_comm_destroy(void) {
    return guarded([&]() -> int {
        if (current_stream()) cudaStreamSynchronize(current_stream()); else cudaDeviceSynchronize();
        g_sharded.release();
        comm::destroy();
        return RFK_OK;
    });
}
int rfk_comm_rank(void) { return comm::rank(); }
int rfk_comm_world(void) { return comm::world(); }
int rfk_comm_p2p(void) { return comm::p2p() ? 1 : 0; }
int rfk_comm_barrier(void) {
    return guarded([&]() -> int {
        if (!comm::active()) throw std::runtime_error("rfk_comm_barrier: rfk_comm_init has not been called");
        cudaStream_t s = current_stream();
        if (!g_sharded.counters) {
            cuda_ok(cudaMalloc(&g_sharded.counters, 8 * sizeof(unsigned long long)), "cudaMalloc(comm counters)");
            cuda_ok(cudaMemsetAsync(g_sharded.counters, 0, 8 * sizeof(unsigned long long), s), "clear comm counters");
        }
        comm::barrier_sum(reinterpret_cast<std::uint64_t*>(g_sharded.counters + 1), 1, s);
        cuda_ok(cudaStreamSynchronize(s), "rfk_comm_barrier");
        return RFK_OK;
    });
}
int rfk_comm_reduce_histogram(float* bins, size_t bins_len, int root) {
    return guarded([&]() -> int {
        if (!bins) throw std::invalid_argument("rfk_comm_reduce_histogram: null histogram");
        comm::reduce_histogram(reinterpret_cast<float4*>(bins), bins_len, root, current_stream());
        return RFK_OK;  // stream-ordered
    });
}
int rfk_comm_row_slab(uint32_t height, uint32_t halo, int rank, int world, rfk_row_slab* out) {
    return guarded([&]() -> int {
        if (!out || height > 0x3fffffffu || halo > 0x3fffffffu) throw std::invalid_argument("rfk_comm_row_slab: bad argument");
        const comm::row_slab s = comm::slab_of((int)height, (int)halo, rank, world);
        *out = rfk_row_slab{(uint32_t)s.y0, (uint32_t)s.y1, (uint32_t)s.src_y0, (uint32_t)s.src_y1};
        return RFK_OK;
    });
}

int rfk_render_frame_sharded(rfk_flame* f, const rfk_sharded_request* sreq, uint8_t* rgba8_out, float* image_out, rfk_sharded_stats* stats) {
    return guarded([&]() -> int {
        if (!f || !sreq) throw std::invalid_argument("rfk_render_frame_sharded: null argument");
        if (!comm::active()) throw std::runtime_error("rfk_render_frame_sharded: rfk_comm_init has not been called");
        const rfk_frame_request* req = &sreq->frame;
        const int rank = comm::rank(), world = comm::world();
        const bool want8 = sreq->want_rgba8 != 0, wantf = sreq->want_image != 0;
        if (!want8 && !wantf) throw std::invalid_argument("rfk_render_frame_sharded: neither want_rgba8 nor want_image");
        if (rank == 0 && ((want8 && !rgba8_out) || (wantf && !image_out))) throw std::invalid_argument("rfk_render_frame_sharded: rank 0 needs the output buffers it asked for");
        if (!req->width || !req->height) throw std::invalid_argument("rfk_render_frame_sharded: empty image");
        if (!req->target_binned && !req->max_draw_calls) throw std::invalid_argument("rfk_render_frame_sharded: neither target_binned nor max_draw_calls given");
        if (!req->drawing_passes) throw std::invalid_argument("rfk_render_frame_sharded: drawing_passes is 0");
        flame* fl = F(f);
        if (fl->options().deterministic) throw std::invalid_argument("rfk_render_frame_sharded: not with the deterministic kernel option (use rfk_flame_draw_to_bins + rfk_comm_reduce_histogram)");
        const int ss = req->supersample > 1 ? (int)req->supersample : 1;
        if (ss > 16) throw std::invalid_argument("rfk_render_frame_sharded: supersample > 16");
        const int OW = (int)req->width, OH = (int)req->height, W = OW * ss, H = OH * ss;
        const size_t n = (size_t)W * H;
        cudaStream_t s = current_stream();

        rfk_post_params pp;
        rfk_flame_post_params(f, &pp);
        pp.scale_constant = (float)(1.0 / std::pow(10.0, (double)req->scale_constant_exp));
        auto d = to_density(pp, W, H);
        const int halo = std::max(d.estimator_radius, d.estimator_min);

        // geometry of every rank: output rows, the rows density estimation produces for them (more than the output rows when
        // the spatial filter reads across the slab border), and the source rows of those
        std::vector<comm::row_slab> out_slabs(world), de_slabs(world);
        for (int r = 0; r < world; r++) {
            out_slabs[r] = comm::slab_of(OH, 0, r, world);
            comm::row_slab de = out_slabs[r];
            if (ss > 1 && de.y1 > de.y0) kernels::spatial_filter_rows(ss, req->filter_radius, out_slabs[r].y0, out_slabs[r].y1, H, &de.y0, &de.y1);
            de.src_y0 = std::max(0, de.y0 - halo);
            de.src_y1 = std::min(H, de.y1 + halo);
            if (de.y1 == de.y0) de.src_y0 = de.src_y1 = de.y0;
            de_slabs[r] = de;
        }
        const comm::row_slab mine_out = out_slabs[rank], mine = de_slabs[rank];
        const size_t out_rows = (size_t)(mine_out.y1 - mine_out.y0), de_rows = (size_t)(mine.y1 - mine.y0), src_rows = (size_t)(mine.src_y1 - mine.src_y0);
        size_t max_out_rows = 0, max_de_rows = 0, max_src_rows = 0;
        for (int r = 0; r < world; r++) {
            max_out_rows = std::max(max_out_rows, (size_t)(out_slabs[r].y1 - out_slabs[r].y0));
            max_de_rows = std::max(max_de_rows, (size_t)(de_slabs[r].y1 - de_slabs[r].y0));
            max_src_rows = std::max(max_src_rows, (size_t)(de_slabs[r].src_y1 - de_slabs[r].src_y0));
        }

        sharded_buffers& b = g_sharded;
        bool reallocated = false;
        auto grow = [&](auto*& ptr, size_t& have, size_t want, const char* what) {
            if (have >= want) return;
            comm::release_peers();  // the peers' mappings of the old buffers die with them
            cudaFree(ptr); ptr = nullptr; have = 0;
            cuda_ok(cudaMalloc(&ptr, want * sizeof(*ptr)), what);
            have = want;
            reallocated = true;
        };
        // the same sizes on every rank (maxima over the ranks), so that all ranks re-allocate in the same call
        grow(b.bins, b.bins_n, n, "cudaMalloc(bins)");
        grow(b.slab, b.slab_n, std::max<size_t>(1, max_src_rows * W), "cudaMalloc(row slab)");
        if (ss > 1) grow(b.image, b.image_n, std::max<size_t>(1, max_de_rows * W), "cudaMalloc(slab image)");
        grow(b.small, b.small_n, std::max<size_t>(1, max_out_rows * OW), "cudaMalloc(slab rows)");
        grow(b.rows8, b.rows8_n, std::max<size_t>(1, max_out_rows * OW), "cudaMalloc(slab rows rgba8)");
        if (want8) grow(b.full8, b.full8_n, (size_t)OW * OH, "cudaMalloc(rgba8)");
        if (wantf) grow(b.full_image, b.full_image_n, (size_t)OW * OH, "cudaMalloc(image)");
        if (!b.counters) { cuda_ok(cudaMalloc(&b.counters, 8 * sizeof(unsigned long long)), "cudaMalloc(comm counters)"); reallocated = true; }
        if (reallocated) b.generation++;
        for (auto& e : b.ev) if (!e) cuda_ok(cudaEventCreate(&e), "cudaEventCreate");
        const comm::peer_buffers& peers = comm::exchange(b.bins, b.full8, b.full_image, b.generation, s);
        const bool p2p = comm::p2p();

        cuda_ok(cudaEventRecord(b.ev[0], s), "event");
        fl->warmup(req->warmup_passes, req->tss_width);
        cuda_ok(cudaMemsetAsync(b.bins, 0, n * sizeof(float4), s), "clear bins");
        cuda_ok(cudaMemsetAsync(b.counters, 0, 8 * sizeof(unsigned long long), s), "clear counters");
        cuda_ok(cudaEventRecord(b.ev[1], s), "event");

        // ---- draw: every rank the same number of passes ----
        const unsigned long long* binned_dev = flame_binned_counter_dev(*fl);
        uint64_t passes = 0, binned_global = 0;
        uint32_t calls = 0;
        auto draw_passes = [&](uint64_t count) {
            while (count) {
                if (req->max_draw_calls && calls >= req->max_draw_calls) return;
                const int now = (int)std::min<uint64_t>(count, req->drawing_passes);
                fl->draw_to_bins_async(reinterpret_cast<float*>(b.bins), n, W, now);
                passes += now; count -= now; calls++;
            }
        };
        double per_pass = 0.0;  // samples one pass of ALL ranks lands
        if (!req->target_binned) {
            draw_passes((uint64_t)req->max_draw_calls * req->drawing_passes);
        } else {
            // a first call on every rank measures what a pass lands. Its count travels to the host behind an event while the GPU
            // already works on the passes that cannot overshoot whatever the count turns out to be (a pass lands at most one
            // sample per particle, so target / (world * particles) passes are always safe); the remainder is sized from the count
            // and enqueued behind them: no bubble on the device, and no host round trip after the last draw call either.
            const uint64_t even_share = (req->target_binned / ((uint64_t)world * sim_total_particles())) + 1;  // passes if every iteration binned
            draw_passes(std::min<uint64_t>(req->drawing_passes, std::max<uint64_t>(1, even_share / 2)));
            const uint64_t first_passes = passes;
            if (!b.host_counter) cuda_ok(cudaMallocHost(&b.host_counter, sizeof(unsigned long long)), "cudaMallocHost(counter)");
            if (!b.counted) cuda_ok(cudaEventCreateWithFlags(&b.counted, cudaEventDisableTiming), "cudaEventCreate");
            cuda_ok(cudaMemcpyAsync(b.counters + 2, binned_dev, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s), "copy binned counter");
            comm::barrier_sum(reinterpret_cast<std::uint64_t*>(b.counters + 2), 1, s);
            cuda_ok(cudaMemcpyAsync(b.host_counter, b.counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "read binned counter");
            cuda_ok(cudaEventRecord(b.counted, s), "event");
            if (even_share - 1 > passes) draw_passes(even_share - 1 - passes);
            cuda_ok(cudaEventSynchronize(b.counted), "read binned counter");
            binned_global = *b.host_counter;
            per_pass = first_passes ? (double)binned_global / (double)first_passes : 0.0;
            if (per_pass > 0.0) {
                // 0.2 % margin; a shortfall (rare) is found when the count is read together with the image and costs one more round
                const uint64_t wanted = (uint64_t)std::ceil((double)req->target_binned / per_pass * 1.002);
                if (wanted > passes) draw_passes(wanted - passes);
            } else if (binned_global < req->target_binned) {
                draw_passes(req->drawing_passes);
            }
        }
        cuda_ok(cudaEventRecord(b.ev[2], s), "event");

        uint64_t final_binned = 0, binned_before_topup = 0;
        for (int attempt = 0;; attempt++) {
            // ---- reduce-scatter over row slabs with halo; the all-reduce of the counters is also the barrier in front of it ----
            cuda_ok(cudaMemcpyAsync(b.counters, binned_dev, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s), "copy binned counter");
            comm::barrier_sum(reinterpret_cast<std::uint64_t*>(b.counters), 1, s);
            if (src_rows) {
                if (p2p) {
                    kernels::slab_reduce(peers.bins, world, (size_t)(H - mine.src_y1) * W, src_rows * W, b.slab, s);
                    count_launch(1);
                }
            }
            if (!p2p) comm::reduce_scatter_slabs_nccl(b.bins, b.slab, W, H, de_slabs, s);
            cuda_ok(cudaGetLastError(), "row slab reduce");
            if (attempt == 0) cuda_ok(cudaEventRecord(b.ev[3], s), "event");

            // ---- density estimation + tonemap (+ spatial filter) on this rank's rows; finished rows go to rank 0 ----
            if (de_rows) {
                auto ds = d;
                ds.y0 = mine.y0; ds.y1 = mine.y1; ds.src_y0 = mine.src_y0; ds.src_y1 = mine.src_y1;
                if (ss == 1) {
                    if (p2p) {  // straight into rank 0's images, rows addressed by their image row
                        ds.out_y0 = 0;
                        kernels::density_tonemap(b.slab, wantf ? peers.root_image : nullptr, want8 ? peers.root_rgba8 : nullptr, ds, true, true, s);
                    } else {
                        ds.out_y0 = mine.y0;
                        kernels::density_tonemap(b.slab, wantf ? b.small : nullptr, want8 ? b.rows8 : nullptr, ds, true, true, s);
                    }
                    count_launch(1);
                } else {
                    ds.out_y0 = mine.y0;
                    kernels::density_tonemap(b.slab, b.image, nullptr, ds, true, true, s);
                    kernels::spatial_downsample_rows(b.image, b.small, OW, OH, ss, req->filter_radius, mine_out.y0, mine_out.y1, mine.y0, s);
                    count_launch(2);
                    if (want8) {
                        kernels::pack_rgba8(b.small, p2p ? peers.root_rgba8 + (size_t)mine_out.y0 * OW : b.rows8, out_rows * OW, s);
                        count_launch(1);
                    }
                    if (wantf && p2p)
                        cuda_ok(cudaMemcpyAsync(peers.root_image + (size_t)mine_out.y0 * OW, b.small, out_rows * OW * sizeof(float4), cudaMemcpyDefault, s), "rows to rank 0");
                }
                cuda_ok(cudaGetLastError(), "density_tonemap launch");
            }
            if (!p2p) {
                if (want8) comm::gather_slabs_nccl(rank == 0 ? nullptr : b.rows8, b.full8, (size_t)OW * sizeof(uchar4), out_slabs, s);
                if (wantf) comm::gather_slabs_nccl(rank == 0 ? nullptr : b.small, b.full_image, (size_t)OW * sizeof(float4), out_slabs, s);
                if (rank == 0 && out_rows) {  // rank 0's own rows
                    if (want8) cuda_ok(cudaMemcpyAsync(b.full8 + (size_t)mine_out.y0 * OW, b.rows8, out_rows * OW * sizeof(uchar4), cudaMemcpyDeviceToDevice, s), "own rows");
                    if (wantf) cuda_ok(cudaMemcpyAsync(b.full_image + (size_t)mine_out.y0 * OW, b.small, out_rows * OW * sizeof(float4), cudaMemcpyDeviceToDevice, s), "own rows");
                }
            }
            if (attempt == 0) cuda_ok(cudaEventRecord(b.ev[4], s), "event");

            // ---- barrier: all rows have landed on rank 0, and no rank clears its histogram before its peers have read it ----
            comm::barrier_sum(reinterpret_cast<std::uint64_t*>(b.counters + 1), 1, s);
            unsigned long long counted = 0;
            cuda_ok(cudaMemcpyAsync(&counted, b.counters, sizeof counted, cudaMemcpyDeviceToHost, s), "read binned counter");
            if (rank == 0) {
                if (want8) cuda_ok(cudaMemcpyAsync(rgba8_out, b.full8, (size_t)OW * OH * sizeof(uchar4), cudaMemcpyDeviceToHost, s), "read back rgba8");
                if (wantf) cuda_ok(cudaMemcpyAsync(image_out, b.full_image, (size_t)OW * OH * sizeof(float4), cudaMemcpyDeviceToHost, s), "read back image");
            }
            cuda_ok(cudaEventRecord(b.ev[5], s), "event");
            cuda_ok(cudaStreamSynchronize(s), "rfk_render_frame_sharded");
            final_binned = counted;
            const bool capped = req->max_draw_calls && calls >= req->max_draw_calls;
            if (!req->target_binned || final_binned >= req->target_binned || capped) break;
            if (attempt >= 8 || (attempt >= 1 && final_binned - binned_before_topup < (req->target_binned - final_binned) / 1000))
                throw std::runtime_error("rfk_render_frame_sharded: the samples stopped landing before the target was reached (the genome's particles overflow and are never reset, as in the reference); render with max_draw_calls instead");
            binned_before_topup = final_binned;
            // short of the target (the in-bounds fraction drifted by more than the margin): top up and redo the exchange
            per_pass = passes ? (double)final_binned / (double)passes : 0.0;
            const double left = (double)(req->target_binned - final_binned);
            draw_passes(std::max<uint64_t>(1, per_pass > 0.0 ? (uint64_t)std::ceil(left / per_pass * 1.01) : req->drawing_passes));
        }
        if (stats) {
            stats->iterations_global = (uint64_t)world * passes * sim_total_particles();
            stats->binned_global = final_binned;
            stats->passes = passes; stats->draw_calls = calls; stats->p2p = p2p ? 1 : 0;
            stats->y0 = (uint32_t)mine_out.y0; stats->y1 = (uint32_t)mine_out.y1;
            cudaEventElapsedTime(&stats->ms_warmup, b.ev[0], b.ev[1]);
            cudaEventElapsedTime(&stats->ms_draw, b.ev[1], b.ev[2]);
            cudaEventElapsedTime(&stats->ms_reduce, b.ev[2], b.ev[3]);
            cudaEventElapsedTime(&stats->ms_post, b.ev[3], b.ev[4]);
            cudaEventElapsedTime(&stats->ms_readback, b.ev[4], b.ev[5]);
        }
        return RFK_OK;
    });
}

// ---- buffer cache ----
int rfk_cache_write_buffer(const char* root, const char* type, const char* group, const void* data, size_t bytes, const char* name, char* name_out, size_t name_out_len) {
    return guarded([&]() -> int {
        if (!root || !type || !group || (!data && bytes)) throw std::invalid_argument("rfk_cache_write_buffer: null argument");
        std::string written = buffer_cache::buffer_group(root, type, group).write_buffer(data, bytes, name ? name : "");
        if (name_out && name_out_len) {
            std::strncpy(name_out, written.c_str(), name_out_len - 1);
            name_out[name_out_len - 1] = '\0';
        }
        return RFK_OK;
    });
}
int64_t rfk_cache_read_buffer(const char* root, const char* type, const char* group, const char* name, void* out, size_t out_len) {
    int64_t result = 0;
    int rc = guarded([&]() -> int {
        if (!root || !type || !group || !name) throw std::invalid_argument("rfk_cache_read_buffer: null argument");
        auto data = buffer_cache::buffer_group(root, type, group).read_buffer(name);
        result = (int64_t)data.size();
        if (out) {
            if (out_len < data.size()) throw std::invalid_argument("rfk_cache_read_buffer: output buffer too small");
            std::memcpy(out, data.data(), data.size());
        }
        return RFK_OK;
    });
    return rc == RFK_OK ? result : rc;
}
int rfk_cache_list(const char* root, const char* type, const char* group, char* names_out, size_t names_out_len) {
    return guarded([&]() -> int {
        if (!root || !type || !group) throw std::invalid_argument("rfk_cache_list: null argument");
        auto names = buffer_cache::buffer_group(root, type, group).cached_buffers();
        std::string joined;
        for (size_t i = 0; i < names.size(); i++) joined += (i ? "\n" : "") + names[i];
        if (names_out && names_out_len) {
            if (joined.size() + 1 > names_out_len) throw std::invalid_argument("rfk_cache_list: output buffer too small");
            std::memcpy(names_out, joined.c_str(), joined.size() + 1);
        }
        return (int)names.size();
    });
}
int rfk_export_sim_cache(const char* root, uint64_t shuffle_seed) {
    return guarded([&]() -> int {
        if (!root) throw std::invalid_argument("rfk_export_sim_cache: null root");
        const size_t P = sim_total_particles(), TS = sim_temporal_samples();
        if (!P) throw std::runtime_error("set_sim_parameters has not been called");
        std::vector<uint32_t> states(P * 4);
        cuda_ok(cudaMemcpyAsync(states.data(), sim_rng_states(), P * sizeof(uint4), cudaMemcpyDeviceToHost, current_stream()), "read rng states");
        cuda_ok(cudaStreamSynchronize(current_stream()), "read rng states");
        buffer_cache::buffer_group("" + std::string(root), "rand_state", std::to_string(P)).write_buffer(states.data(), states.size() * 4);
        const size_t ppt = P / TS, count = sim_shuffle_count();
        if (count) {
            uint32_t* dev = nullptr;
            cuda_ok(cudaMalloc(&dev, ppt * count * sizeof(uint32_t)), "cudaMalloc(shuffle buffers)");
            kernels::make_shuffle_buffers(dev, (uint32_t)ppt, (uint32_t)count, shuffle_seed, current_stream());
            count_launch(1);
            std::vector<uint32_t> host(ppt * count);
            cudaError_t e = cudaMemcpyAsync(host.data(), dev, host.size() * 4, cudaMemcpyDeviceToHost, current_stream());
            if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
            cudaFree(dev);
            cuda_ok(e, "read shuffle buffers");
            buffer_cache::buffer_group group(root, "shuffle", std::to_string(ppt));
            for (size_t k = 0; k < count; k++) group.write_buffer(host.data() + k * ppt, ppt * 4);
        }
        return RFK_OK;
    });
}
int rfk_set_rng_states(const uint32_t* states, size_t first, size_t count) {
    return guarded([&]() -> int {
        if (!states) throw std::invalid_argument("rfk_set_rng_states: null argument");
        if (!sim_rng_states() || first + count > sim_total_particles()) throw std::invalid_argument("range outside the particle RNG states");
        cuda_ok(cudaMemcpyAsync(sim_rng_states_mutable() + first, states, count * sizeof(uint4), cudaMemcpyHostToDevice, current_stream()), "upload rng states");
        cuda_ok(cudaStreamSynchronize(current_stream()), "upload rng states");
        return RFK_OK;
    });
}
int rfk_import_rng_states(const char* root, const char* name) {
    return guarded([&]() -> int {
        if (!root) throw std::invalid_argument("rfk_import_rng_states: null root");
        const size_t P = sim_total_particles();
        if (!P) throw std::runtime_error("set_sim_parameters has not been called");
        buffer_cache::buffer_group group(root, "rand_state", std::to_string(P));
        std::string which = name ? name : "";
        if (which.empty()) {
            auto names = group.cached_buffers();
            if (names.empty()) return fail(RFK_E_NOTFOUND, "no cached rand_state buffer for " + std::to_string(P) + " particles under " + group.path());
            which = names.front();
        }
        auto data = group.read_buffer(which);
        if (data.size() != P * sizeof(uint4)) return fail(RFK_E_INVALID, "cached rand_state buffer has the wrong size");
        return rfk_set_rng_states(reinterpret_cast<const uint32_t*>(data.data()), 0, P);
    });
}

int rfk_write_png(const char* path, const uint8_t* rgba8, size_t width, size_t height) {
    try {
        if (!path || !rgba8) return fail(RFK_E_INVALID, "rfk_write_png: null argument");
        write_png_rgba8(path, rgba8, width, height);
        return RFK_OK;
    } catch (const std::exception& e) { return fail(RFK_E_INVALID, e.what()); }
}

int rfk_write_exr(const char* path, const float* rgba, size_t width, size_t height) {
    try {
        if (!path || !rgba) return fail(RFK_E_INVALID, "rfk_write_exr: null argument");
        write_exr_rgba32f(path, rgba, width, height);
        return RFK_OK;
    } catch (const std::exception& e) { return fail(RFK_E_INVALID, e.what()); }
}

// ---- reference pass mode ----
int rfk_set_shuffle_buffers(const uint32_t* tables, size_t count, uint64_t seed) {
    return guarded([&]() -> int { flame::set_shuffle_buffers(tables, count, seed); return RFK_OK; });
}
int rfk_flame_reference_warmup(rfk_flame* f, size_t num_passes, float tss_width, const uint32_t* ids) {
    return guarded([&]() -> int {
        if (!f) throw std::invalid_argument("null flame");
        F(f)->reference_warmup(num_passes, tss_width, ids);
        return RFK_OK;
    });
}
int64_t rfk_flame_reference_draw_to_bins(rfk_flame* f, float* bins, size_t bins_len, size_t bins_width, int num_iter, const uint32_t* ids) {
    int64_t result = 0;
    int rc = guarded([&]() -> int {
        if (!f || num_iter < 0) throw std::invalid_argument("bad argument");
        if (F(f)->needs_warmup()) return fail(RFK_E_STATE, "reference_draw_to_bins: warmup has not been run for the current parameters");
        result = (int64_t)F(f)->reference_draw_to_bins(bins, bins_len, bins_width, num_iter, ids);
        return RFK_OK;
    });
    return rc == RFK_OK ? result : rc;
}
int rfk_flame_copy_particles(rfk_flame* f, float* out) {
    return guarded([&]() -> int {
        if (!f || !out) throw std::invalid_argument("null argument");
        flame_copy_particles(*F(f), out);
        return RFK_OK;
    });
}

// ---- test hooks ----
int rfk_flame_single_step(rfk_flame* f, int n, const float* xyz, const int* xid, uint32_t* rng, const float* fp, int first_run, float* out) {
    return guarded([&]() -> int {
        if (!f || n < 0 || (n && (!xyz || !xid || !rng || !out))) throw std::invalid_argument("rfk_flame_single_step: bad argument");
        if (n == 0) return RFK_OK;
        flame_single_step(*F(f), n, xyz, xid, rng, fp, first_run, out);
        return RFK_OK;
    });
}
int rfk_flame_select_xform(rfk_flame* f, int n, const float* ratio, const float* fp, int* out) {
    return guarded([&]() -> int {
        if (!f || n < 0 || (n && (!ratio || !out))) throw std::invalid_argument("rfk_flame_select_xform: bad argument");
        if (n == 0) return RFK_OK;
        flame_select_xform(*F(f), n, ratio, fp, out);
        return RFK_OK;
    });
}
int rfk_flame_bucket_index(rfk_flame* f, int n, const float* xyzw, const float ss[6], int W, int H, int* idx, int* pal) {
    return guarded([&]() -> int {
        if (!f || n < 0 || !ss || W <= 0 || H <= 0 || (n && (!xyzw || !idx || !pal))) throw std::invalid_argument("rfk_flame_bucket_index: bad argument");
        if (n == 0) return RFK_OK;
        flame_bucket_index(*F(f), n, xyzw, ss, W, H, idx, pal);
        return RFK_OK;
    });
}
static int copy_out(const std::string& r, char* out, size_t out_len) {
    if (!out || r.size() + 1 > out_len) return fail(RFK_E_INVALID, "output buffer too small");
    std::memcpy(out, r.c_str(), r.size() + 1);
    return (int)r.size();
}
int rfk_text_replace_macro(const char* str, const char* name, const char* value, char* out, size_t out_len) {
    if (!str || !name || !value) return fail(RFK_E_INVALID, "null argument");
    return copy_out(replace_macro(str, name, value), out, out_len);
}
int rfk_text_find_macros(const char* str, char* out, size_t out_len) {
    if (!str) return fail(RFK_E_INVALID, "null argument");
    std::string joined;
    for (auto& m : find_macros(str)) joined += m + "\n";
    return copy_out(joined, out, out_len);
}
int rfk_flame_animate(rfk_flame* f, float tss_width, int temporal_samples, float* out) {
    return guarded([&]() -> int {
        if (!f || !out || temporal_samples <= 0) throw std::invalid_argument("rfk_flame_animate: bad argument");
        flame_animate_host(*F(f), tss_width, temporal_samples, out);
        return RFK_OK;
    });
}

}  // extern "C"
