// CUDA side of the flame object: NVRTC build of the generated kernels, device
// buffers, warmup / draw_to_bins launches. Replaces the GL half of src/flame.cpp
// (do_common_init :17-31, set_sim_parameters :105-158, warmup :228-281,
// draw_to_bins :283-330) — there is no CPU fallback: every entry point here throws
// when no CUDA device / driver is available.
#include "flame_device.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <mutex>
#include <random>
#include <set>
#include <stdexcept>
#include <unordered_map>

#include "static_kernels.cuh"
#include "variation_table.hpp"

namespace rfk {

namespace embedded {
extern const char* const device_prelude;
extern const char* const chaos_kernels;
}  // namespace embedded

// ---------------------------------------------------------------------------------
// driver entry points (resolved through the runtime, so the library has no link-time
// dependency on libcuda and loads on a machine without a GPU)
// ---------------------------------------------------------------------------------
namespace {

struct driver_api {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

const driver_api& driver() {
    static driver_api api = [] {
        driver_api a;
        cuda_check(cudaFree(nullptr), "CUDA initialisation (is a GPU visible?)");
        auto get = [](const char* name, auto& fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            cuda_check(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q), name);
            if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error(std::string("driver entry point not found: ") + name);
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
        };
        get("cuModuleLoadData", a.ModuleLoadData);
        get("cuModuleUnload", a.ModuleUnload);
        get("cuModuleGetFunction", a.ModuleGetFunction);
        get("cuModuleGetGlobal", a.ModuleGetGlobal);
        get("cuLaunchKernel", a.LaunchKernel);
        get("cuFuncGetAttribute", a.FuncGetAttribute);
        get("cuOccupancyMaxActiveBlocksPerMultiprocessor", a.OccupancyMaxActiveBlocksPerMultiprocessor);
        get("cuGetErrorString", a.GetErrorString);
        return a;
    }();
    return api;
}

void cu_check(CUresult r, const char* what) {
    if (r == CUDA_SUCCESS) return;
    const char* msg = nullptr;
    driver().GetErrorString(r, &msg);
    throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "unknown driver error"));
}

// global simulation state: class-static in the reference (src/flame.hpp:150-156)
struct sim_state {
    std::size_t total_particles = 0, temporal_samples = 0, shuffle_count = 0;
    std::uint64_t seed = 0;
    uint4* rng = nullptr;
    std::uint64_t generation = 0;  // bumped by set_sim_parameters: invalidates per-flame buffers (flame.cpp:153-157)
    cudaStream_t stream = nullptr;
    std::uint64_t launches = 0;
    // reference pass mode only
    std::uint32_t* shuffle = nullptr;      // [shuffle_tables][PPT]
    std::size_t shuffle_tables = 0;
    float4* samples = nullptr;             // [PPT] Hammersley points (sample_buffer_)
    std::uint64_t samples_generation = ~0ull, shuffle_generation = ~0ull;
    std::mt19937 pass_ids{0x5EED0001u};
};
sim_state g_sim;
std::set<flame*> g_active_flames;  // src/flame.hpp:173

// compiled modules by (hash of source text + options): cubin and compile log. Bounded: an animation whose frames change
// parameter VALUES builds a new value-specialised module per frame (kernel option specialize = 1); the oldest entries go.
struct cubin_entry { std::string key; std::vector<char> cubin; std::string log; };
std::mutex g_cache_mutex;
std::list<cubin_entry> g_cubin_cache;  // most recently used first
constexpr std::size_t kCubinCacheEntries = 24;

}  // namespace

cudaStream_t current_stream() { return g_sim.stream; }
void set_current_stream(cudaStream_t s) { g_sim.stream = s; }
std::uint64_t kernel_launch_count() { return g_sim.launches; }
void count_launch(unsigned n) { g_sim.launches += n; }

// ---------------------------------------------------------------------------------
// NVRTC
// ---------------------------------------------------------------------------------
std::vector<char> compile_cubin(const std::string& source, const kernel_options& opt, std::string* log_out, int auto_min_blocks) {
    const std::string key = source + "#" + std::to_string(opt.math_mode) + (opt.fmad ? "" : "#M") + "#" + std::to_string(auto_min_blocks);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        for (auto it = g_cubin_cache.begin(); it != g_cubin_cache.end(); ++it) {
            if (it->key != key) continue;
            g_cubin_cache.splice(g_cubin_cache.begin(), g_cubin_cache, it);
            if (log_out) *log_out = it->log;
            return it->cubin;
        }
    }
    // RFK_SOURCE_DUMP_DIR: keep the generated translation unit on disk under the name the line info refers to
    // (lets `ncu --import-source on` and cuobjdump map SASS back to the generated code)
    std::string unit_name = "rfk_chaos_game.cu";
    if (const char* dir = std::getenv("RFK_SOURCE_DUMP_DIR")) {
        std::size_t h = std::hash<std::string>{}(source);
        char tag[32];
        std::snprintf(tag, sizeof tag, "%016zx", h);
        unit_name = std::string(dir) + "/rfk_chaos_game_" + tag + ".cu";
        if (FILE* fh = std::fopen(unit_name.c_str(), "w")) {
            std::fwrite(source.data(), 1, source.size(), fh);
            std::fclose(fh);
        }
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, source.c_str(), unit_name.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)
        throw std::runtime_error("nvrtcCreateProgram failed");
    // --ptxas-options=-v: the log carries registers and spill bytes per kernel (read by flame::cubin's launch-bounds choice)
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "--generate-line-info", "--ptxas-options=-v"};
    const std::string auto_define = "-DRFK_AUTO_MIN_BLOCKS=" + std::to_string(auto_min_blocks);
    if (auto_min_blocks > 0) opts.push_back(auto_define.c_str());
    if (opt.math_mode == 2) opts.push_back("--use_fast_math");
    // mode 1: --use_fast_math for its div.approx lowering of `/`; the prelude keeps expf / logf / tanf / powf on libdevice by name
    if (opt.math_mode == 1) opts.push_back("--use_fast_math");
    if (!opt.fmad) opts.push_back("--fmad=false");
    nvrtcResult res = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    std::size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    std::string log(log_size, '\0');
    if (log_size) nvrtcGetProgramLog(prog, log.data());
    if (log_out) *log_out = log;
    if (res != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        throw std::runtime_error("kernel compilation failed:\n" + log);
    }
    std::size_t size = 0;
    nvrtcGetCUBINSize(prog, &size);
    std::vector<char> cubin(size);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cubin_cache.push_front(cubin_entry{key, cubin, log});
    while (g_cubin_cache.size() > kCubinCacheEntries) g_cubin_cache.pop_back();
    return cubin;
}

// ---------------------------------------------------------------------------------
// flame: construction, codegen products
// ---------------------------------------------------------------------------------
struct rfk_iter_params_host {  // must match rfk_iter_params in chaos_kernels.cuh
    float4* particles;
    uint4* rng;
    const float* fp_inflated;
    const float4* palette;
    float4* bins;
    unsigned long long* fixed_bins;
    unsigned long long* counters;
    float ss_affine[6];
    int bin_w, bin_h;
    float bin_wf, bin_hf;
    int num_iter;
    int ppt;
    int first_run;
    unsigned int deal_seed;
    int hammersley_bits;
    float hammersley_inv_max;
    const unsigned int* hot_map;
    int hot_tiles_x;
    uint2* stage_records;
    unsigned int* stage_cursors;
    unsigned int* stage_fill;
    unsigned int stage_capacity;
    int stage_region_shift;
    int stage_regions;
};

struct rfk_pass_params_host {  // must match rfk_pass_params in chaos_kernels.cuh
    const float4* pos_in;
    float4* pos_out;
    uint4* rng;
    const unsigned int* shuf_buf;
    const float* fp_inflated;
    const float4* palette;
    float4* bins;
    unsigned long long* counters;
    float ss_affine[6];
    int bin_w, bin_h;
    float bin_wf, bin_hf;
    int ppt;
    unsigned int shuf_buf_idx_in, shuf_buf_idx_out;
    int random_read, random_write, first_run, do_draw;
};

struct flame_device {
    CUmodule module = nullptr;
    CUfunction warm = nullptr, draw = nullptr, single_step = nullptr, select_xform = nullptr, bucket_index = nullptr, reference_pass = nullptr;
    float* cfp = nullptr;         // the module's __constant__ rfk_cfp[]: parameter slots that do not depend on the temporal sample
    std::size_t cfp_floats = 0;
    std::vector<float> cfp_staging;  // host copy of the last upload (slots, reciprocals, 1 - slot, slot products)
    // Further builds of the same kernels, made on first use. [staged][baked]:
    //  staged — rfk_draw compiled with RFK_STAGED_BINS (staged_bins = -1, automatic: the first time a histogram of 512 MiB or
    //           more is drawn into); holds rfk_draw alone;
    //  baked  — every rfk_cfp[k] of the text replaced by its current value (kernel option `specialize`); holds rfk_warm and
    //           rfk_draw (and rfk_single_step for the parity tests) and has no constant bank to upload.
    struct variant {
        CUmodule module = nullptr;
        CUfunction warm = nullptr, draw = nullptr, single_step = nullptr;
        float* cfp = nullptr;          // null for baked variants
        std::vector<float> values;     // baked variants: the constants compiled in
    };
    variant variants[2][2];            // [0][0] stays empty: that is `module` above
    variant paired;                    // baked, unstaged, two particles per thread (kernel option pair_particles)
    int pairs_state = 0;               // 0 = not measured for the current baked values, 1 = the paired build is the faster one, 2 = it is not
    float pairs_ms[2] = {0.0f, 0.0f};  // the measurement: ms of the probe launch, one particle per thread / two
    std::vector<float> previous_constants;  // rfk_cfp values of the warmup before the last one (specialize = 2)
    bool use_baked = false;                 // decided by warmup(): the baked variants match the uploaded parameters
    float* pinned = nullptr;                // pinned staging of warmup()'s uploads: fp[1024], constant table [4096], palette [1024]
    cudaEvent_t pinned_done = nullptr;
    float4* particles = nullptr;
    float4* swap = nullptr;  // reference pass mode: swap_buffer_
    float* fp = nullptr;
    float* fp_inflated = nullptr;
    float4* palette = nullptr;
    unsigned long long* counters = nullptr;  // [0] binned, [1..] per-xform picks
    unsigned long long* fixed_bins = nullptr;
    std::size_t fixed_len = 0;
    kernels::animate_xform* anim = nullptr;
    int anim_count = 0;
    std::uint64_t sim_generation = ~0ull;
    unsigned int deal_counter = 0x5EED0001u;
    std::uint64_t binned_reported = 0;
    bool warmed = false;
    // hot map (kernel option l2_hints)
    float* hot_sums = nullptr;
    unsigned int* hot_scratch = nullptr;
    unsigned int* hot_bitmap = nullptr;
    std::size_t hot_capacity_tiles = 0;
    int hot_W = 0, hot_H = 0, hot_tiles_x = 0;  // dimensions the current map was built for; 0 = no map
    // staging queues (kernel option staged_bins)
    uint2* stage_records = nullptr;
    unsigned int* stage_cursors = nullptr;
    unsigned int* stage_fill = nullptr;
    std::size_t stage_regions = 0, stage_capacity = 0;  // queues and chunks per queue the buffers were sized for
    std::size_t stage_requested = 0;                    // chunks per queue asked for (more than stage_capacity when memory was short)
    bool stage_unavailable = false;                     // automatic mode: no memory for the queues, draw directly

    ~flame_device() {
        cudaFree(hot_sums); cudaFree(hot_scratch); cudaFree(hot_bitmap);
        cudaFree(stage_records); cudaFree(stage_cursors); cudaFree(stage_fill);
        if (module) driver().ModuleUnload(module);
        for (auto& row : variants)
            for (auto& v : row)
                if (v.module) driver().ModuleUnload(v.module);
        if (paired.module) driver().ModuleUnload(paired.module);
        cudaFree(particles); cudaFree(swap); cudaFree(fp); cudaFree(fp_inflated); cudaFree(palette);
        cudaFree(counters); cudaFree(fixed_bins); cudaFree(anim);
        if (pinned_done) { cudaEventSynchronize(pinned_done); cudaEventDestroy(pinned_done); }
        cudaFreeHost(pinned);
    }
};

flame::flame() { g_active_flames.insert(this); }
flame::~flame() { g_active_flames.erase(this); }

void flame::rebuild_cuda_source() {
    const int n = (int)xforms.size();
    std::string s;
    s += "// generated by refrakt_b200 for one genome: options, prelude, dispatch, kernels\n";
    s += "#define RFK_BLOCK " + std::to_string(options_.block_width) + "\n";
    s += "#define RFK_DEAL_PERIOD " + std::to_string(options_.deal_period) + "\n";
    int log2b = 0;
    while ((1 << log2b) < options_.block_width) log2b++;
    s += "#define RFK_LOG2_BLOCK " + std::to_string(log2b) + "\n";
    s += "#define RFK_TOTAL_PARAMS " + std::to_string(buffer_map_.size) + "\n";
    s += "#define RFK_NUM_XFORMS " + std::to_string(n) + "\n";
    s += "#define RFK_HAS_FINAL " + std::to_string(final_xform ? 1 : 0) + "\n";
    s += "#define RFK_MATH_MODE " + std::to_string(options_.math_mode) + "\n";
    s += "#define RFK_PER_LANE_XFORM " + std::to_string(options_.per_lane_xform ? 1 : 0) + "\n";
    s += "#define RFK_WARP_AGGREGATE " + std::to_string(options_.warp_aggregate ? 1 : 0) + "\n";
    s += "#define RFK_DETERMINISTIC " + std::to_string(options_.deterministic ? 1 : 0) + "\n";
    s += "#define RFK_COUNT_XFORMS " + std::to_string(options_.count_xforms ? 1 : 0) + "\n";
    s += "#define RFK_L2_HINTS " + std::to_string(options_.l2_hints ? 1 : 0) + "\n";
    s += "#define RFK_STAGED_BINS " + std::to_string(options_.staged_bins > 0 ? 1 : 0) + "\n";
    if (const char* e = std::getenv("RFK_EXPERIMENT")) s += "#define RFK_EXPERIMENT " + std::string(e) + "\n";  // timing builds of tools/gpu_probe_l1tex.sh
    // min_blocks 0 = automatic (flame::cubin): 2048 resident threads per SM (32 registers) unless that spills, else 1536
    // (40 registers) — tools/probe_min_blocks.py; -1 = leave the register budget to the compiler
    if (options_.min_blocks > 0) s += "#define RFK_LAUNCH_BOUNDS __launch_bounds__(RFK_BLOCK, " + std::to_string(options_.min_blocks) + ")\n";
    else if (options_.min_blocks == 0) s += "#define RFK_LAUNCH_BOUNDS __launch_bounds__(RFK_BLOCK, RFK_AUTO_MIN_BLOCKS)\n";  // chosen by flame::cubin()
    else s += "#define RFK_LAUNCH_BOUNDS __launch_bounds__(RFK_BLOCK)\n";
    s += embedded::device_prelude;
    s += "\nnamespace rfk_glsl {\n#define randf() rfk_randf(rs)\n";
    s += cuda_body_;
    s += "#undef randf\n}  // namespace rfk_glsl\n\n";
    s += embedded::chaos_kernels;
    cuda_source_ = std::move(s);
    cubin_.clear();
    if (device_) device_.reset();  // module must be rebuilt
}

bool flame::do_common_init(const flame_compiler& fc) {
    make_shader_buffer_map();
    if (buffer_map_.size > PARAM_BUFFER) {
        set_last_error("flame needs " + std::to_string(buffer_map_.size) + " parameter slots; the parameter buffer holds " + std::to_string(PARAM_BUFFER));
        return false;
    }
    try {
        glsl_source_ = fc.compile_flame_xforms(*this);
        cuda_body_ = fc.compile_flame_cuda(*this);
    } catch (const std::exception& e) {
        set_last_error(std::string("flame compiler: ") + e.what());
        return false;
    }
    rebuild_cuda_source();
    reset_animation();
    try {
        cubin();  // the reference fails the load when the shader does not compile (flame.cpp:30)
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return false;
    }
    return true;
}

bool flame::set_options(const kernel_options& opt) {
    if (opt == options_) return true;
    {   // `specialize` alone does not change the generic module
        kernel_options same = opt;
        same.specialize = options_.specialize;
        if (same == options_) { options_.specialize = opt.specialize; needs_update_ = true; return true; }
    }
    kernel_options old = options_;
    options_ = opt;
    rebuild_cuda_source();
    try {
        cubin();
    } catch (const std::exception& e) {
        set_last_error(e.what());
        options_ = old;
        rebuild_cuda_source();
        return false;
    }
    needs_update_ = true;
    return true;
}

// bytes ptxas reports as spilled by the two hot kernels ("N bytes spill stores" after "Compiling entry function 'rfk_draw'")
static std::size_t spilled_bytes(const std::string& log) {
    std::size_t total = 0;
    for (const char* kernel : {"'rfk_draw'", "'rfk_warm'"}) {
        std::size_t at = log.find(std::string("Compiling entry function ") + kernel);
        if (at == std::string::npos) continue;
        std::size_t end = log.find("Compiling entry function", at + 10);
        std::size_t sp = log.find("bytes spill stores", at);
        if (sp == std::string::npos || (end != std::string::npos && sp > end)) continue;
        std::size_t b = log.rfind(',', sp);
        if (b == std::string::npos) continue;
        total += std::strtoull(log.c_str() + b + 1, nullptr, 10);
    }
    return total;
}

static constexpr std::size_t kSpillTolerance = 128;  // bytes, rfk_draw + rfk_warm together (stress genome: 48 bytes, +4 % at full occupancy)

// automatic launch bounds (min_blocks 0): full occupancy (2048 threads per SM, 32 registers per thread) when the genome's
// kernels fit without spilling (more than kSpillTolerance bytes), else 1536 threads per SM (40 registers)
static std::vector<char> compile_with_bounds(const std::string& source, const kernel_options& opt) {
    if (opt.min_blocks != 0) return compile_cubin(source, opt, nullptr, 0);
    std::string log;
    std::vector<char> tight = compile_cubin(source, opt, &log, 2048 / opt.block_width);
    // a few spilled bytes are loop-invariant values parked before the loop and re-read inside one xform's body (an L1 hit):
    // cheaper than giving up a quarter of the resident warps
    return spilled_bytes(log) <= kSpillTolerance ? std::move(tight) : compile_cubin(source, opt, nullptr, 1536 / opt.block_width);
}

const std::vector<char>& flame::cubin() {
    if (cubin_.empty()) cubin_ = compile_with_bounds(cuda_source_, options_);
    return cubin_;
}

// C literal of a binary32 value, exact (hexadecimal floating literal; non-finite values through the compiler's builtins)
static std::string float_literal(float v) {
    if (std::isnan(v)) return "__builtin_nanf(\"\")";
    if (std::isinf(v)) return v > 0 ? "__builtin_huge_valf()" : "(-__builtin_huge_valf())";
    char buf[64];
    std::snprintf(buf, sizeof buf, "(%af)", (double)v);
    return buf;
}

// The translation unit of a kernel variant, derived from the generic one by text:
//  staged: `#define RFK_STAGED_BINS 1` and rfk_draw alone (half the build time);
//  baked:  every `rfk_cfp[k]` (k is always a literal in the generated text) becomes the value of that slot, so the compiler
//          sees the genome's weights, variation amounts, parameters and their host-derived companions as immediates: the
//          constant-bank loads (LDC / LDCU, a tenth of the instructions of the shipped genome's rfk_draw) disappear, the
//          cumulative weights of get_xform_id() fold into compare immediates, and selects on parameters
//          (`rfk_cfp[n] == 0 ? ... : ...`) fold away. The four rotated affine coefficients of every xform differ between
//          temporal samples and stay in shared memory (fp[k]).
std::string flame::variant_source(bool staged, const std::vector<float>* baked, bool pairs) const {
    if (pairs && (!baked || staged)) throw std::invalid_argument("variant_source: the paired kernels exist in the value-specialised, unstaged build only");
    std::string source = cuda_source_;
    if (staged) {
        const std::string off = "#define RFK_STAGED_BINS 0\n", on = "#define RFK_STAGED_BINS 1\n#define RFK_DRAW_ONLY 1\n";
        const std::size_t at = source.find(off);
        if (at == std::string::npos) throw std::runtime_error("variant_source: the kernels are already compiled with staging");
        source.replace(at, off.size(), on);
    }
    if (baked) {
        const std::string decl_head = "extern \"C\" { __constant__ float rfk_cfp[";
        const std::size_t decl = source.find(decl_head);
        if (decl == std::string::npos) throw std::runtime_error("variant_source: no rfk_cfp declaration in the generated text");
        const std::size_t decl_end = source.find('\n', decl);
        source.replace(decl, decl_end - decl, std::string("#define RFK_BAKED 1  // rfk_cfp[] compiled in as literals\n#define RFK_HOT_ONLY 1") +
                                                  (pairs ? "\n#define RFK_PAIRS 1" + std::string(std::getenv("RFK_PAIRS_MIN_BLOCKS") ? std::string("\n#define RFK_PAIRS_MIN_BLOCKS ") + std::getenv("RFK_PAIRS_MIN_BLOCKS") : "") : ""));
        // `e RFK_DIVC(n, r)` (device_prelude.cuh) names its slots inside a macro: expanded here, so the pass below sees them
        for (std::size_t at = source.find("RFK_DIVC("); at != std::string::npos; at = source.find("RFK_DIVC(", at + 1)) {
            std::size_t j = at + 9, comma = source.find(',', j), close = source.find(')', j);
            if (comma == std::string::npos || close == std::string::npos || comma > close || !std::isdigit((unsigned char)source[j])) continue;  // the macro's own definition
            std::size_t r0 = comma + 1;
            while (r0 < close && source[r0] == ' ') r0++;
            const std::string n = source.substr(j, comma - j), r = source.substr(r0, close - r0);
            source.replace(at, close + 1 - at, options_.math_mode == 0 ? "/ rfk_cfp[" + n + "]" : "* rfk_cfp[" + r + "]");
        }
        // An xform that does not rotate (rotation_frequency 0: animate.tpl.glsl multiplies its angle by it) has the same four
        // coefficients in every temporal sample: RFK_AFF(i, c) of such an xform is a literal as well, and the kernel's
        // 128-bit read for it is dead code.
        auto bake_affine = [&](const xform_slots& m, int index) {
            if ((std::size_t)m.rotation_frequency >= baked->size() || (*baked)[m.rotation_frequency] != 0.0f) return;
            for (int a = 0; a < 4; a++) {
                const std::string token = "RFK_AFF(" + std::to_string(index) + ", " + std::string(1, "xyzw"[a]) + ")";
                const std::string value = float_literal((*baked)[m.affine[a]]);
                for (std::size_t at = source.find(token); at != std::string::npos; at = source.find(token, at + value.size())) source.replace(at, token.size(), value);
            }
        };
        for (std::size_t i = 0; i < buffer_map_.xforms.size(); i++) bake_affine(buffer_map_.xforms[i], (int)i);
        if (buffer_map_.final_xform) bake_affine(*buffer_map_.final_xform, -1);
        std::string out;
        out.reserve(source.size() + 16 * 1024);
        const std::string key = "rfk_cfp[";
        std::size_t pos = 0;
        for (;;) {
            const std::size_t at = source.find(key, pos);
            if (at == std::string::npos) break;
            const bool ident_before = at > 0 && (std::isalnum((unsigned char)source[at - 1]) || source[at - 1] == '_');
            std::size_t j = at + key.size(), k = 0;
            bool digits = false;
            while (j < source.size() && std::isdigit((unsigned char)source[j])) { k = k * 10 + (source[j] - '0'); j++; digits = true; }
            if (ident_before || !digits || j >= source.size() || source[j] != ']') {  // a comment or another identifier: leave it
                out.append(source, pos, at + key.size() - pos);
                pos = at + key.size();
                continue;
            }
            if (k >= baked->size()) throw std::runtime_error("variant_source: rfk_cfp index outside the parameter table");
            out.append(source, pos, at - pos);
            out += float_literal((*baked)[k]);
            pos = j + 1;
        }
        out.append(source, pos, std::string::npos);
        source = std::move(out);
    }
    return source;
}

bool flame::pairs_allowed() const {
    const kernel_options& o = options_;
    return o.pair_particles && !o.per_lane_xform && !o.warp_aggregate && !o.deterministic && !o.count_xforms && !o.l2_hints && o.staged_bins <= 0 &&
           o.deal_period == 1 && o.min_blocks == 0;
}

std::vector<char> flame::variant_cubin(bool staged, const std::vector<float>* baked, bool pairs) const {
    return compile_with_bounds(variant_source(staged, baked, pairs), options_);
}

void flame::reset_animation() { needs_update_ = true; }

// ---------------------------------------------------------------------------------
// simulation parameters
// ---------------------------------------------------------------------------------
void flame::set_sim_parameters(std::size_t total_particles, std::size_t temporal_samples, std::size_t shuffle_count, std::uint64_t seed) {
    if (temporal_samples == 0 || total_particles == 0) throw std::invalid_argument("set_sim_parameters: particle and temporal sample counts must be positive");
    if (total_particles % temporal_samples != 0 || (total_particles / temporal_samples) % BLOCK_WIDTH != 0)
        throw std::invalid_argument("set_sim_parameters: particles per temporal sample must be a multiple of " + std::to_string(BLOCK_WIDTH));
    if (total_particles / BLOCK_WIDTH > 0x7fffffffull) throw std::invalid_argument("set_sim_parameters: too many particles");
    driver();
    if (g_sim.rng) cuda_check(cudaFree(g_sim.rng), "cudaFree(rng)");
    g_sim.rng = nullptr;
    cuda_check(cudaMalloc(&g_sim.rng, total_particles * sizeof(uint4)), "cudaMalloc(rng states)");
    g_sim.total_particles = total_particles;
    g_sim.temporal_samples = temporal_samples;
    g_sim.shuffle_count = shuffle_count;
    g_sim.seed = seed;
    g_sim.generation++;
    kernels::seed_rng_states(g_sim.rng, total_particles, (std::uint32_t)seed, g_sim.stream);
    count_launch(1);
    cuda_check(cudaGetLastError(), "seed_rng_states");
    // invalidate existing flames
    for (auto* f : g_active_flames)
        if (f->device()) f->device()->sim_generation = ~0ull;
}

// Releases the class-static simulation buffers (the reference frees them when set_sim_parameters replaces them,
// flame.cpp:107-149, and at process exit); live flames must call set_sim_parameters + warmup again.
void release_sim_buffers() {
    if (g_sim.stream) cudaStreamSynchronize(g_sim.stream);
    cudaFree(g_sim.rng); g_sim.rng = nullptr;
    cudaFree(g_sim.shuffle); g_sim.shuffle = nullptr; g_sim.shuffle_tables = 0;
    cudaFree(g_sim.samples); g_sim.samples = nullptr;
    g_sim.total_particles = g_sim.temporal_samples = g_sim.shuffle_count = 0;
    g_sim.samples_generation = g_sim.shuffle_generation = ~0ull;
    g_sim.generation++;
    for (auto* f : g_active_flames)
        if (f->device()) f->device()->sim_generation = ~0ull;
}

std::size_t sim_total_particles() { return g_sim.total_particles; }
std::size_t sim_temporal_samples() { return g_sim.temporal_samples; }
const uint4* sim_rng_states() { return g_sim.rng; }
uint4* sim_rng_states_mutable() { return g_sim.rng; }
std::size_t sim_shuffle_count() { return g_sim.shuffle_count; }

// ---------------------------------------------------------------------------------
// device state
// ---------------------------------------------------------------------------------
static void ensure_module(flame& f) {
    if (!f.device_slot()) f.device_slot() = std::make_unique<flame_device>();
    flame_device& d = *f.device();
    if (d.module) return;
    const auto& api = driver();
    const auto& image = f.cubin();
    cu_check(api.ModuleLoadData(&d.module, image.data()), "cuModuleLoadData");
    cu_check(api.ModuleGetFunction(&d.warm, d.module, "rfk_warm"), "rfk_warm");
    cu_check(api.ModuleGetFunction(&d.draw, d.module, "rfk_draw"), "rfk_draw");
    cu_check(api.ModuleGetFunction(&d.single_step, d.module, "rfk_single_step"), "rfk_single_step");
    cu_check(api.ModuleGetFunction(&d.select_xform, d.module, "rfk_select_xform"), "rfk_select_xform");
    cu_check(api.ModuleGetFunction(&d.bucket_index, d.module, "rfk_bucket_index"), "rfk_bucket_index");
    cu_check(api.ModuleGetFunction(&d.reference_pass, d.module, "rfk_reference_pass"), "rfk_reference_pass");
    CUdeviceptr cfp = 0;
    std::size_t cfp_bytes = 0;
    cu_check(api.ModuleGetGlobal(&cfp, &cfp_bytes, d.module, "rfk_cfp"), "rfk_cfp");
    d.cfp = reinterpret_cast<float*>(cfp);
    d.cfp_floats = cfp_bytes / sizeof(float);
}

// builds (on first use) and returns a further variant of the kernels; a baked variant whose values are stale is rebuilt
static flame_device::variant& ensure_variant(flame& f, bool staged, bool baked, bool pairs = false) {
    flame_device& d = *f.device();
    flame_device::variant& v = pairs ? d.paired : d.variants[staged][baked];
    if (v.module && (!baked || v.values == d.cfp_staging)) return v;
    const auto& api = driver();
    if (v.module) { api.ModuleUnload(v.module); v = flame_device::variant{}; }
    if (baked && !staged) d.pairs_state = 0;  // new values: measure again
    const std::vector<char> image = f.variant_cubin(staged, baked ? &d.cfp_staging : nullptr, pairs);
    cu_check(api.ModuleLoadData(&v.module, image.data()), "cuModuleLoadData(variant)");
    cu_check(api.ModuleGetFunction(&v.draw, v.module, "rfk_draw"), "rfk_draw(variant)");
    if (!staged) {
        cu_check(api.ModuleGetFunction(&v.warm, v.module, "rfk_warm"), "rfk_warm(variant)");
        cu_check(api.ModuleGetFunction(&v.single_step, v.module, "rfk_single_step"), "rfk_single_step(variant)");
    }
    if (baked) {
        v.values = d.cfp_staging;
    } else {
        CUdeviceptr cfp = 0;
        std::size_t cfp_bytes = 0;
        cu_check(api.ModuleGetGlobal(&cfp, &cfp_bytes, v.module, "rfk_cfp"), "rfk_cfp(variant)");
        v.cfp = reinterpret_cast<float*>(cfp);
        if (!d.cfp_staging.empty())  // the parameters of the last warmup
            cuda_check(cudaMemcpyAsync(v.cfp, d.cfp_staging.data(), std::min(cfp_bytes, d.cfp_staging.size() * sizeof(float)), cudaMemcpyHostToDevice, g_sim.stream),
                       "upload constant parameters");
    }
    return v;
}

static void ensure_buffers(flame& f) {
    ensure_module(f);
    flame_device& d = *f.device();
    if (g_sim.total_particles == 0) throw std::runtime_error("set_sim_parameters has not been called");
    if ((g_sim.total_particles / g_sim.temporal_samples) % f.options().block_width != 0)
        throw std::invalid_argument("particles per temporal sample must be a multiple of the kernel's block_width");
    const int total_params = f.buffer_map().size;
    if (!d.fp) {
        cuda_check(cudaMalloc(&d.fp, flame::PARAM_BUFFER * sizeof(float)), "cudaMalloc(fp)");
        cuda_check(cudaMalloc(&d.palette, 256 * sizeof(float4)), "cudaMalloc(palette)");
        cuda_check(cudaMalloc(&d.counters, 64 * sizeof(unsigned long long)), "cudaMalloc(counters)");
        std::vector<kernels::animate_xform> ax;
        auto add = [&](const xform_slots& m) {
            kernels::animate_xform a;
            for (int i = 0; i < 6; i++) a.affine[i] = m.affine[i];
            a.rotation_frequency = m.rotation_frequency;
            ax.push_back(a);
        };
        for (auto& m : f.buffer_map().xforms) add(m);
        if (f.buffer_map().final_xform) add(*f.buffer_map().final_xform);
        d.anim_count = (int)ax.size();
        cuda_check(cudaMalloc(&d.anim, ax.size() * sizeof(kernels::animate_xform)), "cudaMalloc(animate table)");
        cuda_check(cudaMemcpyAsync(d.anim, ax.data(), ax.size() * sizeof(kernels::animate_xform), cudaMemcpyHostToDevice, g_sim.stream), "upload animate table");
        cuda_check(cudaStreamSynchronize(g_sim.stream), "upload animate table");
    }
    if (d.sim_generation != g_sim.generation) {
        cudaFree(d.particles); d.particles = nullptr;
        cudaFree(d.swap); d.swap = nullptr;
        cudaFree(d.fp_inflated); d.fp_inflated = nullptr;
        cuda_check(cudaMalloc(&d.particles, g_sim.total_particles * sizeof(float4)), "cudaMalloc(particles)");
        cuda_check(cudaMalloc(&d.fp_inflated, g_sim.temporal_samples * (std::size_t)total_params * sizeof(float)), "cudaMalloc(fp_inflated)");
        d.sim_generation = g_sim.generation;
        d.warmed = false;
    }
}

bool flame::needs_warmup() const {
    return needs_update_ || !device_ || !device_->particles || !device_->fp_inflated || device_->sim_generation != g_sim.generation || !device_->warmed;
}

static void launch(CUfunction fn, unsigned grid, unsigned block, void** args) {
    cu_check(driver().LaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, (CUstream)g_sim.stream, args, nullptr), "cuLaunchKernel");
    count_launch(1);
}

// The table behind `rfk_cfp[]` in the generated text (compile_flame_cuda), four quarters of max(1, buffer map size) floats:
// the slots; their reciprocals, read by `e / slot` (RFK_DIVC); 1 - slot and slot[i - 1] * slot[i], the two constants of
// the colour blend (RFK_MIXC, indexed by the colour-speed slot). Pure host arithmetic in binary32.
std::vector<float> flame::constant_table(const float* fp) const {
    const std::size_t size = std::min<std::size_t>(std::max(1, buffer_map_.size), PARAM_BUFFER);
    std::vector<float> t(4 * size);
    for (std::size_t i = 0; i < size; i++) {
        t[i] = fp[i];
        t[size + i] = 1.0f / fp[i];
    }
    // the blend constants exist at the colour-speed slots only (anywhere else they would tie the table — the key of the
    // value-specialised build — to neighbouring slots for nothing, e.g. the translation e to the rotating coefficient d)
    auto blend = [&](const xform_slots& m) {
        const std::size_t i = (std::size_t)m.color_speed;
        if (i == 0 || i >= size) return;
        volatile float prod = fp[i - 1] * fp[i];  // rounded to binary32, like the device's FMUL
        t[2 * size + i] = 1.0f - fp[i];
        t[3 * size + i] = prod;
    };
    for (const auto& m : buffer_map_.xforms) blend(m);
    if (buffer_map_.final_xform) blend(*buffer_map_.final_xform);
    // The rotated coefficients (a, b, c, d) of an xform that rotates are never read from this table — the kernels take them
    // from the per-temporal-sample rows (RFK_AFF) — and they change with every frame of an animation (src/main.cpp:383-395):
    // left out, so that the table, which is also the key of the value-specialised build, stays the same from frame to frame.
    auto blank = [&](const xform_slots& m) {
        if (fp[m.rotation_frequency] == 0.0f) return;  // does not rotate: its coefficients are constants like any other slot
        for (int a = 0; a < 4; a++)
            for (int q = 0; q < 4; q++)
                if ((std::size_t)m.affine[a] < size) t[q * size + m.affine[a]] = 0.0f;
    };
    for (const auto& m : buffer_map_.xforms) blank(m);
    if (buffer_map_.final_xform) blank(*buffer_map_.final_xform);
    return t;
}

// the generic kernels read every slot that is the same for all temporal samples from constant memory (rfk_cfp)
static void upload_constant_params(flame& f, const float* fp) {
    flame_device& d = *f.device();
    if (!d.cfp || !d.cfp_floats) return;
    d.cfp_staging = f.constant_table(fp);
    const std::size_t bytes = std::min(d.cfp_floats, d.cfp_staging.size()) * sizeof(float);
    cuda_check(cudaMemcpyAsync(d.cfp, d.cfp_staging.data(), bytes, cudaMemcpyHostToDevice, g_sim.stream), "upload constant parameters");
    if (float* staged_cfp = d.variants[1][0].cfp)
        cuda_check(cudaMemcpyAsync(staged_cfp, d.cfp_staging.data(), bytes, cudaMemcpyHostToDevice, g_sim.stream), "upload constant parameters");
}

static rfk_iter_params_host base_params(flame& f) {
    flame_device& d = *f.device();
    rfk_iter_params_host p{};
    p.particles = d.particles;
    p.rng = g_sim.rng;
    p.fp_inflated = d.fp_inflated;
    p.palette = d.palette;
    p.counters = d.counters;
    p.ppt = (int)(g_sim.total_particles / g_sim.temporal_samples);
    // src/hammersley.cpp:33-42
    std::uint32_t count = (std::uint32_t)p.ppt, max = count;
    if (count % 2 != 0) { max = count - 1; max |= max >> 1; max |= max >> 2; max |= max >> 4; max |= max >> 8; max |= max >> 16; max++; }
    p.hammersley_inv_max = 1.0f / max;
    p.hammersley_bits = 0;
    for (std::uint32_t v = max; v >>= 1;) p.hammersley_bits++;
    p.deal_seed = d.deal_counter;
    d.deal_counter = d.deal_counter * 1664525u + 1013904223u;
    return p;
}

// Kernel option pair_particles = 1: which of the two value-specialised builds — one particle per thread, or two — is the
// faster one for this genome is MEASURED once per build: both run rfk_warm (first-run pass + 24 iterations, no histogram) on
// scratch copies of the particle and RNG buffers, CUDA-event timed; nothing the render reads or writes is touched, so a run
// is the same with and without the measurement. (Shipped genome: the paired build is 3.7 % faster; the 12-xform stress
// genome: 10 % slower — its xform bodies, inlined twice, no longer fit the instruction cache.)
static void choose_pairs(flame& f) {
    flame_device& d = *f.device();
    const kernel_options& o = f.options();
    if (d.pairs_state != 0) return;
    if (!f.pairs_allowed()) { d.pairs_state = 2; return; }
    if (o.pair_particles == 2) { ensure_variant(f, false, true, true); d.pairs_state = 1; return; }
    flame_device::variant* candidates[2] = {&ensure_variant(f, false, true, false), &ensure_variant(f, false, true, true)};
    d.pairs_state = 0;  // (ensure_variant resets it when it rebuilds)
    const std::size_t P = g_sim.total_particles;
    float4* particles = nullptr;
    uint4* rng = nullptr;
    cudaEvent_t ev[2][2] = {};
    float ms[2] = {0.0f, 0.0f};
    try {
        cuda_check(cudaMalloc(&particles, P * sizeof(float4)), "cudaMalloc(tuning scratch)");
        cuda_check(cudaMalloc(&rng, P * sizeof(uint4)), "cudaMalloc(tuning scratch)");
        for (auto& pair : ev) for (auto& e : pair) cuda_check(cudaEventCreate(&e), "cudaEventCreate");
        const unsigned int deal_counter = d.deal_counter;
        for (int round = 0; round < 2; round++) {  // round 0 loads the code, round 1 is timed
            for (int k = 0; k < 2; k++) {
                cuda_check(cudaMemcpyAsync(rng, g_sim.rng, P * sizeof(uint4), cudaMemcpyDeviceToDevice, g_sim.stream), "copy rng states");
                rfk_iter_params_host p = base_params(f);
                p.particles = particles;
                p.rng = rng;
                p.first_run = 1;
                p.num_iter = 24;
                void* args[] = {&p};
                cuda_check(cudaEventRecord(ev[k][0], g_sim.stream), "event");
                launch(candidates[k]->warm, (unsigned)(P / o.block_width), k ? o.block_width / 2 : o.block_width, args);
                cuda_check(cudaEventRecord(ev[k][1], g_sim.stream), "event");
            }
        }
        d.deal_counter = deal_counter;
        cuda_check(cudaStreamSynchronize(g_sim.stream), "pair_particles measurement");
        for (int k = 0; k < 2; k++) cuda_check(cudaEventElapsedTime(&ms[k], ev[k][0], ev[k][1]), "event time");
    } catch (...) {
        cudaFree(particles); cudaFree(rng);
        for (auto& pair : ev) for (auto& e : pair) if (e) cudaEventDestroy(e);
        throw;
    }
    cudaFree(particles); cudaFree(rng);
    for (auto& pair : ev) for (auto& e : pair) cudaEventDestroy(e);
    d.pairs_state = ms[1] < ms[0] ? 1 : 2;
    d.pairs_ms[0] = ms[0]; d.pairs_ms[1] = ms[1];
}

void flame::warmup(std::size_t num_passes, float tss_width) {
    ensure_buffers(*this);
    flame_device& d = *device_;
    needs_update_ = false;

    // The parameter block, the constant table and the palette go up through ONE pinned staging buffer owned by the flame:
    // truly asynchronous copies (a copy from pageable memory makes the driver synchronise the stream first), and warmup
    // returns without waiting — a frame of an animation is then one host synchronisation, the one that hands over the image.
    // The event keeps the next warmup from overwriting the staging buffer while these copies are still in flight.
    auto buf = copy_flame_data_to_buffer();
    d.cfp_staging = constant_table(buf.data());
    const std::size_t table_floats = std::min(d.cfp_floats, d.cfp_staging.size());
    if (!d.pinned) {
        cuda_check(cudaMallocHost(&d.pinned, (PARAM_BUFFER + 4 * PARAM_BUFFER + 256 * 4) * sizeof(float)), "cudaMallocHost(parameter staging)");
        cuda_check(cudaEventCreateWithFlags(&d.pinned_done, cudaEventDisableTiming), "cudaEventCreate");
    } else {
        cuda_check(cudaEventSynchronize(d.pinned_done), "wait for the previous parameter upload");
    }
    float* stage_fp = d.pinned, *stage_table = d.pinned + PARAM_BUFFER, *stage_palette = d.pinned + 5 * PARAM_BUFFER;
    std::memcpy(stage_fp, buf.data(), PARAM_BUFFER * sizeof(float));
    std::memcpy(stage_table, d.cfp_staging.data(), table_floats * sizeof(float));
    std::memcpy(stage_palette, palette.data(), 256 * sizeof(float4));
    cuda_check(cudaMemcpyAsync(d.fp, stage_fp, PARAM_BUFFER * sizeof(float), cudaMemcpyHostToDevice, g_sim.stream), "upload fp");
    if (d.cfp && table_floats) {
        cuda_check(cudaMemcpyAsync(d.cfp, stage_table, table_floats * sizeof(float), cudaMemcpyHostToDevice, g_sim.stream), "upload constant parameters");
        if (float* staged_cfp = d.variants[1][0].cfp)
            cuda_check(cudaMemcpyAsync(staged_cfp, stage_table, table_floats * sizeof(float), cudaMemcpyHostToDevice, g_sim.stream), "upload constant parameters");
    }
    cuda_check(cudaMemcpyAsync(d.palette, stage_palette, 256 * sizeof(float4), cudaMemcpyHostToDevice, g_sim.stream), "upload palette");
    cuda_check(cudaEventRecord(d.pinned_done, g_sim.stream), "event");
    cuda_check(cudaMemsetAsync(d.counters, 0, 64 * sizeof(unsigned long long), g_sim.stream), "clear counters");
    kernels::animate(d.fp, d.fp_inflated, buffer_map_.size, (int)g_sim.temporal_samples, tss_width, d.anim, d.anim_count, g_sim.stream);
    count_launch(1);

    // the re-deal keys restart with every warmup, so a run is reproducible from (seed, parameters) alone
    d.deal_counter = 0x5EED0001u ^ (unsigned int)(g_sim.seed * 0x9E3779B9ull);
    // kernel option `specialize`: 1 = the value-specialised kernels always (rebuilt when a value changed since they were
    // built); 2 = once these values have been seen by two warmups in a row (a still rendered repeatedly, an animation whose
    // frames only rotate affines, a benchmark) or while the ones built earlier still match
    const bool seen_before = d.previous_constants == d.cfp_staging;
    const bool have_match = d.variants[0][1].module && d.variants[0][1].values == d.cfp_staging;
    d.use_baked = !d.cfp_staging.empty() && (options_.specialize == 1 || (options_.specialize == 2 && (seen_before || have_match)));
    d.previous_constants = d.cfp_staging;
    CUfunction warm_fn = d.warm;
    bool pairs = false;
    if (d.use_baked) {
        ensure_variant(*this, false, true);  // (re)built here when the values changed: the measurement below starts over
        choose_pairs(*this);
        pairs = d.pairs_state == 1;
        warm_fn = ensure_variant(*this, false, true, pairs).warm;
    }

    rfk_iter_params_host p = base_params(*this);
    p.first_run = 1;
    p.num_iter = (int)num_passes;
    void* args[] = {&p};
    const unsigned warm_threads = pairs ? options_.block_width / 2 : options_.block_width;
    launch(warm_fn, (unsigned)(g_sim.total_particles / options_.block_width), warm_threads, args);
    d.binned_reported = 0;  // stream-ordered: the draw calls, read-backs and rfk_synchronize that follow wait for it
    d.warmed = true;
}

// ---------------------------------------------------------------------------------
// reference pass mode (same-hardware baseline + pass-level parity hook)
// ---------------------------------------------------------------------------------
void flame::set_shuffle_buffers(const std::uint32_t* host_tables, std::size_t count, std::uint64_t seed) {
    if (g_sim.total_particles == 0) throw std::runtime_error("set_sim_parameters has not been called");
    if (count == 0) throw std::invalid_argument("set_shuffle_buffers: no tables");
    const std::size_t ppt = g_sim.total_particles / g_sim.temporal_samples;
    cudaFree(g_sim.shuffle); g_sim.shuffle = nullptr;
    cuda_check(cudaMalloc(&g_sim.shuffle, count * ppt * sizeof(std::uint32_t)), "cudaMalloc(shuffle buffers)");
    if (host_tables) {
        cuda_check(cudaMemcpyAsync(g_sim.shuffle, host_tables, count * ppt * sizeof(std::uint32_t), cudaMemcpyHostToDevice, g_sim.stream), "upload shuffle buffers");
        cuda_check(cudaStreamSynchronize(g_sim.stream), "upload shuffle buffers");
    } else {
        kernels::make_shuffle_buffers(g_sim.shuffle, (std::uint32_t)ppt, (std::uint32_t)count, seed, g_sim.stream);
        count_launch(1);
    }
    g_sim.shuffle_tables = count;
    g_sim.shuffle_generation = g_sim.generation;
    g_sim.pass_ids.seed(0x5EED0001u ^ (unsigned)seed);
}

static void ensure_reference_buffers(flame& f) {
    ensure_buffers(f);
    flame_device& d = *f.device();
    const std::size_t ppt = g_sim.total_particles / g_sim.temporal_samples;
    if (ppt % 256 != 0) throw std::invalid_argument("reference pass mode: particles per temporal sample must be a multiple of 256");
    if (!g_sim.shuffle || g_sim.shuffle_generation != g_sim.generation) flame::set_shuffle_buffers(nullptr, g_sim.shuffle_count ? g_sim.shuffle_count : 64, g_sim.seed);
    if (!g_sim.samples || g_sim.samples_generation != g_sim.generation) {
        cudaFree(g_sim.samples); g_sim.samples = nullptr;
        cuda_check(cudaMalloc(&g_sim.samples, ppt * sizeof(float4)), "cudaMalloc(sample points)");
        kernels::make_sample_points(g_sim.samples, (std::uint32_t)ppt, g_sim.stream);
        count_launch(1);
        g_sim.samples_generation = g_sim.generation;
    }
    if (!d.swap) cuda_check(cudaMalloc(&d.swap, g_sim.total_particles * sizeof(float4)), "cudaMalloc(swap buffer)");
}

static void launch_pass(flame& f, rfk_pass_params_host& p) {
    flame_device& d = *f.device();
    void* args[] = {&p};
    const unsigned gx = (unsigned)(p.ppt / 256), gy = (unsigned)g_sim.temporal_samples;
    cu_check(driver().LaunchKernel(d.reference_pass, gx, gy, 1, 256, 1, 1, 0, (CUstream)g_sim.stream, args, nullptr), "cuLaunchKernel(rfk_reference_pass)");
    count_launch(1);
}

static std::uint32_t next_pass_id(const std::uint32_t* ids, std::size_t& cursor) {
    if (ids) return ids[cursor++];
    std::uniform_int_distribution<int> dist(0, int(g_sim.shuffle_tables - 1));  // flame.cpp:233
    return (std::uint32_t)dist(g_sim.pass_ids);
}

void flame::reference_warmup(std::size_t num_passes, float tss_width, const std::uint32_t* shuffle_ids) {
    ensure_reference_buffers(*this);
    flame_device& d = *device_;
    needs_update_ = false;
    auto buf = copy_flame_data_to_buffer();
    cuda_check(cudaMemcpyAsync(d.fp, buf.data(), PARAM_BUFFER * sizeof(float), cudaMemcpyHostToDevice, g_sim.stream), "upload fp");
    upload_constant_params(*this, buf.data());
    cuda_check(cudaMemcpyAsync(d.palette, palette.data(), 256 * sizeof(float4), cudaMemcpyHostToDevice, g_sim.stream), "upload palette");
    cuda_check(cudaMemsetAsync(d.counters, 0, 64 * sizeof(unsigned long long), g_sim.stream), "clear counters");
    kernels::animate(d.fp, d.fp_inflated, buffer_map_.size, (int)g_sim.temporal_samples, tss_width, d.anim, d.anim_count, g_sim.stream);
    count_launch(1);

    rfk_pass_params_host p{};
    p.rng = g_sim.rng; p.shuf_buf = g_sim.shuffle; p.fp_inflated = d.fp_inflated; p.palette = d.palette; p.counters = d.counters;
    p.ppt = (int)(g_sim.total_particles / g_sim.temporal_samples);
    p.random_read = 1; p.random_write = 1; p.first_run = 1; p.do_draw = 0;
    std::size_t cursor = 0;
    p.shuf_buf_idx_in = next_pass_id(shuffle_ids, cursor);
    p.shuf_buf_idx_out = next_pass_id(shuffle_ids, cursor);
    p.pos_in = g_sim.samples; p.pos_out = d.particles;
    launch_pass(*this, p);  // flame.cpp:252-264
    p.first_run = 0;
    float4* names[2] = {d.particles, d.swap};
    for (std::size_t i = 0; i < num_passes; i++) {  // flame.cpp:269-280
        p.pos_in = names[i % 2]; p.pos_out = names[(i + 1) % 2];
        p.shuf_buf_idx_in = next_pass_id(shuffle_ids, cursor);
        p.shuf_buf_idx_out = next_pass_id(shuffle_ids, cursor);
        launch_pass(*this, p);
    }
    if (num_passes % 2) std::swap(d.particles, d.swap);
    cuda_check(cudaStreamSynchronize(g_sim.stream), "reference_warmup");
    d.binned_reported = 0;
    d.warmed = true;
}

std::size_t flame::reference_draw_to_bins(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter, const std::uint32_t* shuffle_ids) {
    if (needs_warmup()) throw std::runtime_error("reference_draw_to_bins: warmup has not been run for the current parameters");
    ensure_reference_buffers(*this);
    if (!bins || bins_width == 0 || bins_len < bins_width) throw std::invalid_argument("reference_draw_to_bins: bad bins buffer");
    flame_device& d = *device_;
    const std::size_t W = bins_width, H = bins_len / bins_width;
    if (W >= (1u << 23) || H >= (1u << 23) || W * H > 0x7fffffffull) throw std::invalid_argument("reference_draw_to_bins: histogram too large");
    rfk_pass_params_host p{};
    p.rng = g_sim.rng; p.shuf_buf = g_sim.shuffle; p.fp_inflated = d.fp_inflated; p.palette = d.palette; p.counters = d.counters;
    p.bins = reinterpret_cast<float4*>(bins);
    p.ppt = (int)(g_sim.total_particles / g_sim.temporal_samples);
    auto ss = screen_space_affine(W, H);
    for (int i = 0; i < 6; i++) p.ss_affine[i] = ss[i];
    p.bin_w = (int)W; p.bin_h = (int)H; p.bin_wf = (float)W; p.bin_hf = (float)H;
    p.random_read = 1; p.random_write = 0; p.first_run = 0; p.do_draw = 1;  // flame.cpp:301-304
    float4* names[2] = {d.particles, d.swap};
    std::size_t cursor = 0;
    for (int i = 0; i < num_iter; i++) {  // flame.cpp:317-325
        p.pos_in = names[i % 2]; p.pos_out = names[(i + 1) % 2];
        p.shuf_buf_idx_in = next_pass_id(shuffle_ids, cursor);
        p.shuf_buf_idx_out = next_pass_id(shuffle_ids, cursor);
        launch_pass(*this, p);
    }
    if (num_iter % 2) std::swap(d.particles, d.swap);
    std::uint64_t total = binned_total();
    std::uint64_t delta = total - d.binned_reported;
    d.binned_reported = total;
    return (std::size_t)delta;
}

int flame_pairs_state(const flame& f, float ms_out[2]) {
    flame_device* d = const_cast<flame&>(f).device();
    if (!d) return 0;
    if (ms_out) { ms_out[0] = d->pairs_ms[0]; ms_out[1] = d->pairs_ms[1]; }
    return d->use_baked ? d->pairs_state : 0;
}

bool flame_uses_baked(const flame& f) { return const_cast<flame&>(f).device() && const_cast<flame&>(f).device()->use_baked; }

const unsigned long long* flame_binned_counter_dev(flame& f) { return f.device() ? f.device()->counters : nullptr; }

void flame_copy_particles(flame& f, float* out) {
    if (!f.device() || !f.device()->particles) throw std::runtime_error("no particle buffer (warmup first)");
    cuda_check(cudaMemcpyAsync(out, f.device()->particles, g_sim.total_particles * sizeof(float4), cudaMemcpyDeviceToHost, g_sim.stream), "copy particles");
    cuda_check(cudaStreamSynchronize(g_sim.stream), "copy particles");
}

void flame::draw_to_bins_async(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter) {
    if (needs_warmup()) throw std::runtime_error("draw_to_bins: warmup() has not been run for the current parameters");
    if (!bins || bins_width == 0 || bins_len < bins_width) throw std::invalid_argument("draw_to_bins: bad bins buffer");
    flame_device& d = *device_;
    const std::size_t W = bins_width, H = bins_len / bins_width;  // flame.cpp:290
    if (W >= (1u << 23) || H >= (1u << 23) || W * H > 0x7fffffffull) throw std::invalid_argument("draw_to_bins: histogram too large for 32-bit bin indices");

    rfk_iter_params_host p = base_params(*this);
    auto ss = screen_space_affine(W, H);
    for (int i = 0; i < 6; i++) p.ss_affine[i] = ss[i];
    p.bin_w = (int)W;
    p.bin_h = (int)H;
    p.bin_wf = (float)W;
    p.bin_hf = (float)H;
    p.bins = reinterpret_cast<float4*>(bins);
    p.num_iter = num_iter;
    p.first_run = 0;
    if (options_.l2_hints && d.hot_W == (int)W && d.hot_H == (int)H) {
        p.hot_map = d.hot_bitmap;
        p.hot_tiles_x = d.hot_tiles_x;
    }

    if (options_.deterministic) {
        if (d.fixed_len != W * H) {
            cudaFree(d.fixed_bins); d.fixed_bins = nullptr;
            cuda_check(cudaMalloc(&d.fixed_bins, W * H * 4 * sizeof(unsigned long long)), "cudaMalloc(fixed-point bins)");
            d.fixed_len = W * H;
        }
        cuda_check(cudaMemsetAsync(d.fixed_bins, 0, W * H * 4 * sizeof(unsigned long long), g_sim.stream), "clear fixed-point bins");
        p.fixed_bins = d.fixed_bins;
    }
    int stage_regions = 0, stage_shift = options_.staged_bins > 0 ? options_.staged_bins : 0;
    const bool baked = d.use_baked && !d.cfp_staging.empty();
    bool pairs = baked && d.pairs_state == 1;
    CUfunction draw_fn = baked ? ensure_variant(*this, false, true, pairs).draw : d.draw;
    if (options_.staged_bins < 0 && !d.stage_unavailable && !options_.deterministic && !options_.warp_aggregate && !options_.l2_hints &&
        W * H * sizeof(float4) >= (std::size_t(1) << 29)) {
        // automatic: a histogram of 512 MiB or more (four times the L2) is drawn through the queues, in at most 64 regions of
        // at least 2^22 bins (64 MB, half the L2; measured best on the 2.12 GB histogram of config 3). The crossover was measured
        // (profiles/r01_staged_threshold_probe.json): 299 MB direct 1.89 ms per call / queued 3.16; 531 MB 3.64 / 3.21; 944 MB 5.37 / 3.23
        int shift = 22;
        if (const char* e = std::getenv("RFK_STAGE_SHIFT")) shift = std::max(8, std::min(24, std::atoi(e)));  // tuning runs
        while (((W * H + (std::size_t(1) << shift) - 1) >> shift) > 64) shift++;
        if (shift <= 24) {  // a record holds 24 bits of bin index
            draw_fn = ensure_variant(*this, true, baked).draw;
            pairs = false;
            stage_shift = shift;
        }
    }
    if (stage_shift > 0) {
        // regions of 2^staged_bins bins, one queue of 4 KB chunks (512 records) per region. A queue holds sixteen times the
        // even share of the call's samples plus one open chunk per CTA (beyond that: direct reductions), 16 GiB at most in total
        // (environment variable RFK_STAGE_MAX_BYTES).
        constexpr std::size_t chunk = 512, max_regions = 64;
        const int shift = stage_shift;
        const std::size_t regions = (W * H + (std::size_t(1) << shift) - 1) >> shift;
        if (regions > max_regions) throw std::runtime_error("staged_bins: " + std::to_string(regions) + " regions of 2^" + std::to_string(shift) +
                                                            " bins; at most 64 (raise staged_bins)");
        const std::size_t samples = g_sim.total_particles * (std::size_t)num_iter, ctas = g_sim.total_particles / options_.block_width;
        std::size_t capacity = std::min(samples / chunk + ctas, 16 * samples / (chunk * regions) + ctas) + 1;
        std::size_t max_bytes = std::size_t(16) << 30;
        if (const char* e = std::getenv("RFK_STAGE_MAX_BYTES")) max_bytes = std::max<std::size_t>(1, std::strtoull(e, nullptr, 10));  // tests: exhausted queues
        max_bytes = std::min(max_bytes, std::size_t(31) << 30);  // record indices are 32-bit in the kernel
        capacity = std::max<std::size_t>(1, std::min(capacity, max_bytes / (chunk * sizeof(uint2)) / regions));
        if (d.stage_regions != regions || d.stage_requested != capacity) {
            auto release = [&] {
                cudaFree(d.stage_records); cudaFree(d.stage_cursors); cudaFree(d.stage_fill);
                d.stage_records = nullptr; d.stage_cursors = d.stage_fill = nullptr; d.stage_regions = d.stage_capacity = d.stage_requested = 0;
            };
            const char* fail_above = std::getenv("RFK_STAGE_FAIL_ABOVE_BYTES");  // tests: pretend larger allocations fail
            auto allocate = [&](std::size_t chunks) {
                if (fail_above && regions * chunks * chunk * sizeof(uint2) > std::strtoull(fail_above, nullptr, 10)) return false;
                return cudaMalloc(&d.stage_records, regions * chunks * chunk * sizeof(uint2)) == cudaSuccess &&
                       cudaMalloc(&d.stage_cursors, regions * sizeof(unsigned int)) == cudaSuccess &&
                       cudaMalloc(&d.stage_fill, regions * chunks * sizeof(unsigned int)) == cudaSuccess;
            };
            release();
            // queues that do not fit the free memory are halved (what overflows is reduced directly); in the automatic mode a
            // device without even 1 GiB to spare draws without staging
            std::size_t chunks = capacity;
            while (!allocate(chunks)) {
                release();
                cudaGetLastError();  // clear the allocation failure
                chunks /= 2;
                if (regions * chunks * chunk * sizeof(uint2) < (std::size_t(1) << 30) && regions * capacity * chunk * sizeof(uint2) >= (std::size_t(1) << 30)) { chunks = 0; break; }
                if (chunks == 0) break;
            }
            if (chunks == 0) {
                if (options_.staged_bins > 0) throw std::runtime_error("staged_bins: no device memory for the region queues");
                d.stage_unavailable = true;
            } else {
                cuda_check(cudaMemsetAsync(d.stage_cursors, 0, regions * sizeof(unsigned int), g_sim.stream), "clear staging cursors");
                d.stage_regions = regions; d.stage_capacity = chunks; d.stage_requested = capacity;
            }
        }
        if (d.stage_unavailable) {
            draw_fn = baked ? ensure_variant(*this, false, true, pairs = baked && d.pairs_state == 1).draw : d.draw;
            stage_shift = 0;
        }
    }
    if (stage_shift > 0) {
        const std::size_t regions = d.stage_regions;
        const int shift = stage_shift;
        stage_regions = (int)regions;
        p.stage_records = d.stage_records;
        p.stage_cursors = d.stage_cursors;
        p.stage_fill = d.stage_fill;
        p.stage_capacity = (unsigned int)d.stage_capacity;
        p.stage_region_shift = shift;
        p.stage_regions = stage_regions;
    }
    void* args[] = {&p};
    launch(draw_fn, (unsigned)(g_sim.total_particles / options_.block_width), pairs ? options_.block_width / 2 : options_.block_width, args);
    if (stage_regions) {
        kernels::stage_accumulate(d.stage_records, d.stage_cursors, d.stage_fill, p.stage_capacity, p.stage_region_shift, stage_regions, d.palette, p.bins,
                                  W * H, g_sim.stream);
        cuda_check(cudaMemsetAsync(d.stage_cursors, 0, d.stage_regions * sizeof(unsigned int), g_sim.stream), "clear staging cursors");
        count_launch(1);
    }
    if (options_.deterministic) {
        kernels::fixed_to_float(d.fixed_bins, p.bins, W * H, g_sim.stream);
        count_launch(1);
    }
}

flame::hot_map_info flame::build_hot_map(const float* bins, std::size_t bins_len, std::size_t bins_width, std::uint64_t budget_bytes) {
    if (!bins || bins_width == 0 || bins_len < bins_width) throw std::invalid_argument("build_hot_map: bad bins buffer");
    ensure_buffers(*this);
    flame_device& d = *device_;
    const std::size_t W = bins_width, H = bins_len / bins_width;
    if (W >= (1u << 23) || H >= (1u << 23) || W * H > 0x7fffffffull) throw std::invalid_argument("build_hot_map: histogram too large");
    if (budget_bytes == 0) {
        int dev = 0, l2 = 0;
        cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
        cuda_check(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev), "query L2 size");
        budget_bytes = (std::uint64_t)l2 / 2;
    }
    const int T = kernels::HOT_MAP_TILE;
    hot_map_info info;
    info.tiles_x = (int)((W + T - 1) / T);
    info.tiles_y = (int)((H + T - 1) / T);
    info.budget_bytes = budget_bytes;
    const std::size_t n_tiles = (std::size_t)info.tiles_x * info.tiles_y;
    if (n_tiles > d.hot_capacity_tiles) {
        cudaFree(d.hot_sums); cudaFree(d.hot_bitmap); d.hot_sums = nullptr; d.hot_bitmap = nullptr; d.hot_capacity_tiles = 0;
        cuda_check(cudaMalloc(&d.hot_sums, n_tiles * sizeof(float)), "cudaMalloc(hot map tile sums)");
        cuda_check(cudaMalloc(&d.hot_bitmap, ((n_tiles + 31) / 32) * sizeof(unsigned int)), "cudaMalloc(hot map)");
        d.hot_capacity_tiles = n_tiles;
    }
    if (!d.hot_scratch) cuda_check(cudaMalloc(&d.hot_scratch, kernels::HOT_MAP_SCRATCH_WORDS * sizeof(unsigned int)), "cudaMalloc(hot map scratch)");
    const std::uint64_t budget_tiles = budget_bytes / ((std::uint64_t)T * T * sizeof(float4));
    kernels::build_hot_map(reinterpret_cast<const float4*>(bins), (int)W, (int)H, (unsigned int)std::min<std::uint64_t>(budget_tiles, 0xffffffffull),
                           d.hot_sums, d.hot_scratch, d.hot_bitmap, g_sim.stream);
    count_launch(3);
    unsigned int tail[2] = {0, 0};
    cuda_check(cudaMemcpyAsync(tail, d.hot_scratch + kernels::HOT_MAP_BUCKETS, sizeof(tail), cudaMemcpyDeviceToHost, g_sim.stream), "read hot map summary");
    cuda_check(cudaStreamSynchronize(g_sim.stream), "build_hot_map");
    info.threshold_bucket = tail[0];
    info.hot_tiles = tail[1];
    d.hot_W = (int)W; d.hot_H = (int)H; d.hot_tiles_x = info.tiles_x;
    return info;
}

void flame::clear_hot_map() {
    if (device_) { device_->hot_W = device_->hot_H = device_->hot_tiles_x = 0; }
}

std::vector<std::uint32_t> flame::copy_hot_map() {
    std::vector<std::uint32_t> out;
    if (!device_ || device_->hot_W == 0) return out;
    flame_device& d = *device_;
    const int T = kernels::HOT_MAP_TILE;
    const std::size_t n_tiles = (std::size_t)d.hot_tiles_x * ((d.hot_H + T - 1) / T);
    out.resize((n_tiles + 31) / 32);
    cuda_check(cudaMemcpyAsync(out.data(), d.hot_bitmap, out.size() * sizeof(std::uint32_t), cudaMemcpyDeviceToHost, g_sim.stream), "copy hot map");
    cuda_check(cudaStreamSynchronize(g_sim.stream), "copy hot map");
    return out;
}

std::uint64_t flame::binned_total() {
    if (!device_ || !device_->counters) return 0;
    unsigned long long total = 0;
    cuda_check(cudaMemcpyAsync(&total, device_->counters, sizeof(total), cudaMemcpyDeviceToHost, g_sim.stream), "read binned counter");
    cuda_check(cudaStreamSynchronize(g_sim.stream), "read binned counter");
    return total;
}

std::size_t flame::draw_to_bins(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter) {
    draw_to_bins_async(bins, bins_len, bins_width, num_iter);
    std::uint64_t total = binned_total();  // blocks, like counters_.get_one(0) at flame.cpp:329
    std::uint64_t delta = total - device_->binned_reported;
    device_->binned_reported = total;
    return (std::size_t)delta;
}

// ---------------------------------------------------------------------------------
// test hooks (device work on host-supplied vectors)
// ---------------------------------------------------------------------------------
namespace {
template <typename T>
struct dev_buf {
    T* p = nullptr;
    std::size_t n = 0;
    explicit dev_buf(std::size_t count) : n(count) { cuda_check(cudaMalloc(&p, (count ? count : 1) * sizeof(T)), "cudaMalloc(test buffer)"); }
    ~dev_buf() { cudaFree(p); }
    void upload(const T* h) { cuda_check(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice), "upload"); }
    void download(T* h) { cuda_check(cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost), "download"); }
};
}  // namespace

void flame_single_step(flame& f, int n, const float* xyz, const int* xid, std::uint32_t* rng, const float* fp, int first_run, float* out) {
    ensure_module(f);
    dev_buf<float> d_xyz(3 * (std::size_t)n), d_fp(flame::PARAM_BUFFER), d_out(4 * (std::size_t)n);
    dev_buf<int> d_xid(n);
    dev_buf<std::uint32_t> d_rng(4 * (std::size_t)n);
    d_xyz.upload(xyz); d_xid.upload(xid); d_rng.upload(rng);
    std::array<float, flame::PARAM_BUFFER> own;
    if (!fp) { own = f.copy_flame_data_to_buffer(); fp = own.data(); }
    d_fp.upload(fp);
    upload_constant_params(f, fp);
    float4* outp = reinterpret_cast<float4*>(d_out.p);
    uint4* rngp = reinterpret_cast<uint4*>(d_rng.p);
    void* args[] = {&n, &d_xyz.p, &d_xid.p, &rngp, &d_fp.p, &first_run, &outp};
    // kernel option specialize = 1 and the flame's own values: through the value-specialised build (built here if need be),
    // so the parity tests reach the code rfk_draw runs
    CUfunction fn = f.device()->single_step;
    if (f.options().specialize == 1 && fp == own.data()) fn = ensure_variant(f, false, true).single_step;
    launch(fn, (unsigned)((n + 127) / 128), 128, args);
    cuda_check(cudaStreamSynchronize(g_sim.stream), "rfk_single_step");
    d_out.download(out); d_rng.download(rng);
    if (fp != own.data()) { own = f.copy_flame_data_to_buffer(); upload_constant_params(f, own.data()); }  // the flame's own values again
}

void flame_select_xform(flame& f, int n, const float* ratio, const float* fp, int* out) {
    ensure_module(f);
    dev_buf<float> d_ratio(n), d_fp(flame::PARAM_BUFFER);
    dev_buf<int> d_out(n);
    d_ratio.upload(ratio);
    std::array<float, flame::PARAM_BUFFER> own;
    if (!fp) { own = f.copy_flame_data_to_buffer(); fp = own.data(); }
    d_fp.upload(fp);
    upload_constant_params(f, fp);
    void* args[] = {&n, &d_ratio.p, &d_fp.p, &d_out.p};
    launch(f.device()->select_xform, (unsigned)((n + 127) / 128), 128, args);
    cuda_check(cudaStreamSynchronize(g_sim.stream), "rfk_select_xform");
    d_out.download(out);
    if (fp != own.data()) { own = f.copy_flame_data_to_buffer(); upload_constant_params(f, own.data()); }
}

void flame_bucket_index(flame& f, int n, const float* xyzw, const float ss_affine[6], int W, int H, int* idx_out, int* pal_out) {
    ensure_module(f);
    dev_buf<float> d_in(4 * (std::size_t)n);
    dev_buf<int> d_idx(n), d_pal(n);
    d_in.upload(xyzw);
    struct { float ss[6]; int w, h; } bp;
    for (int i = 0; i < 6; i++) bp.ss[i] = ss_affine[i];
    bp.w = W; bp.h = H;
    void* args[] = {&n, &d_in.p, &bp, &d_idx.p, &d_pal.p};
    launch(f.device()->bucket_index, (unsigned)((n + 127) / 128), 128, args);
    cuda_check(cudaStreamSynchronize(g_sim.stream), "rfk_bucket_index");
    d_idx.download(idx_out); d_pal.download(pal_out);
}

void flame_animate_host(flame& f, float tss_width, int temporal_samples, float* out) {
    ensure_module(f);
    const int total = f.buffer_map().size;
    dev_buf<float> d_fp(flame::PARAM_BUFFER), d_out((std::size_t)temporal_samples * total);
    auto buf = f.copy_flame_data_to_buffer();
    d_fp.upload(buf.data());
    std::vector<kernels::animate_xform> ax;
    auto add = [&](const xform_slots& m) {
        kernels::animate_xform a;
        for (int i = 0; i < 6; i++) a.affine[i] = m.affine[i];
        a.rotation_frequency = m.rotation_frequency;
        ax.push_back(a);
    };
    for (auto& m : f.buffer_map().xforms) add(m);
    if (f.buffer_map().final_xform) add(*f.buffer_map().final_xform);
    dev_buf<kernels::animate_xform> d_ax(ax.size());
    d_ax.upload(ax.data());
    kernels::animate(d_fp.p, d_out.p, total, temporal_samples, tss_width, d_ax.p, (int)ax.size(), g_sim.stream);
    count_launch(1);
    cuda_check(cudaStreamSynchronize(g_sim.stream), "animate");
    d_out.download(out);
}

void flame_kernel_info(flame& f, const char* kernel, int* regs, int* smem_bytes, int* blocks_per_sm) {
    ensure_module(f);
    CUfunction fn = nullptr;
    cu_check(driver().ModuleGetFunction(&fn, f.device()->module, kernel), kernel);
    cu_check(driver().FuncGetAttribute(regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn), "regs");
    cu_check(driver().FuncGetAttribute(smem_bytes, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, fn), "smem");
    cu_check(driver().OccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, f.options().block_width, 0), "occupancy");
}

void flame_read_counters(flame& f, unsigned long long* out, int n) {
    if (!f.device() || !f.device()->counters) { std::memset(out, 0, n * sizeof(*out)); return; }
    cuda_check(cudaMemcpyAsync(out, f.device()->counters, (std::size_t)(n < 64 ? n : 64) * sizeof(*out), cudaMemcpyDeviceToHost, g_sim.stream), "read counters");
    cuda_check(cudaStreamSynchronize(g_sim.stream), "read counters");
}

}  // namespace rfk
