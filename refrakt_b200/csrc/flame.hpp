// Host-side flame object: genome model, flam3 parser, parameter-buffer layout and
// the warmup / draw_to_bins entry points. Same names, argument meaning and error
// behaviour as the reference's `struct flame` (src/flame.hpp:21-174, src/flame.cpp),
// with the OpenGL compute path replaced by sm_100a CUDA kernels.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace rfk {

class flame_compiler;
struct flame_device;  // CUDA-side state (flame_device.cpp)

// src/flame.hpp:15-19. The reference declares it and keeps a map of it per xform (:36), but its parser for the <motion>
// children of an xform is commented out (src/flame.cpp:199-210) and nothing ever reads the map. Here the children are parsed
// as that block intends — one entry per animated attribute, frequency and function shared by the element — with the
// attribute's value as the amplitude (the block stores it in `freq`, leaving `amplitude` unset: its bug), and
// flame::apply_motion evaluates them the way flam3 does: value(t) = loaded value + amplitude * f(frequency * t),
// f = sin: sin(2 pi x); triangle: 1 - 4 |frac(x + 1/4) - 1/2| ... (see apply_motion); hill: (1 - cos(2 pi x)) / 2.
struct motion_info {
    float freq = 0;
    std::string function;
    float amplitude = 0;
};

// src/flame.hpp:21-37
struct flame_xform {
    using affine_t = std::array<float, 6>;

    affine_t affine{};
    std::optional<affine_t> post;
    std::map<std::string, float> variations;
    std::map<std::string, float> var_param;

    float weight = 0;
    float color = 0;
    float color_speed = 0;

    float rotation_frequency = 0;
    float opacity = 0;

    std::map<std::string, motion_info> motion;  // src/flame.hpp:36; key = the animated attribute (a variation, a parameter, color, ...)
};

// Float-slot layout of one xform inside fp[] (src/flame.cpp:39-60).
struct xform_slots {
    int start = 0, end = 0, size = 0;
    int weight = 0;
    std::array<int, 6> affine{};
    bool has_post = false;
    std::array<int, 6> post{};
    std::map<std::string, int> variations;
    std::map<std::string, int> param;
    int color = 0, color_speed = 0, opacity = 0, rotation_frequency = 0;
};

// The reference keeps this as nlohmann::json (src/flame.cpp:33-71); same content.
struct buffer_map_t {
    std::vector<xform_slots> xforms;
    std::optional<xform_slots> final_xform;
    int size = 0;
    std::string dump_json() const;
};

// Options of the generated chaos-game kernel (no reference counterpart: the GLSL
// path has a single mode).
struct kernel_options {
    int math_mode = 1;            // 0 libdevice + IEEE div/sqrt, 1 balanced (default, inside the 1e-5 contract), 2 --use_fast_math
    bool fmad = true;             // allow FMA contraction of a*b+c
    bool per_lane_xform = false;  // every particle picks its own xform (divergent); default one pick per warp
    bool warp_aggregate = false;  // match_any de-duplication of same-bin updates inside a warp
    bool deterministic = false;   // fixed-point (integer) accumulation: bit-identical histograms
    bool count_xforms = false;    // per-xform selection counters
    int min_blocks = 0;           // __launch_bounds__ second argument; 0 = automatic (2048 / block_width unless that spills, else 1536 / block_width), -1 = compiler's choice
    int block_width = 256;        // threads per CTA = particles per re-deal pool (128, 256 or 512; the reference's workgroup is 256)
    int deal_period = 1;          // re-deal particles across warps every n-th iteration
    int l2_hints = 0;             // histograms much larger than L2: evict-first reductions outside the hot map
    int staged_bins = -1;         // histograms much larger than L2: samples go to per-region queues (regions of 2^staged_bins bins,
                                  // 22 = 64 MB) and are accumulated region by region after the draw kernel; 0 = off;
                                  // -1 = automatic: on for histograms of 512 MiB or more, in regions of 2^22 bins or larger
    int specialize = 2;           // value-specialised kernels: every parameter slot that is the same for all temporal samples is compiled in
                                  // as a literal (operand-form immediates instead of constant-bank loads, branches on parameters folded).
                                  // 0 = off (one module per genome structure, as the reference's one shader per genome);
                                  // 1 = rebuilt by warmup() whenever a value changed; 2 = automatic (default): built the second time
                                  // warmup() runs with unchanged values (stills, animations, benchmarks), generic kernels meanwhile
    int pair_particles = 1;       // value-specialised build only: two particles per thread (CTAs of block_width / 2 threads over the same
                                  // pool of block_width particles): one pick, one walk to the xform's case, one re-deal key and one
                                  // barrier per two iterations. 1 (default) = measured: when the specialised kernels are built, both
                                  // forms run a short warm-up on scratch copies of the particle and RNG buffers and the faster one is
                                  // kept (genomes with many heavy xforms lose with pairs); 2 = always; 0 = never. Not with
                                  // per_lane_xform, warp_aggregate, deterministic, count_xforms, l2_hints, staged_bins, deal_period > 1
    bool operator==(const kernel_options&) const = default;
};

struct flame {
    static constexpr int BLOCK_WIDTH = 256;     // src/flame.cpp:15
    static constexpr int PARAM_BUFFER = 1024;   // src/flame.hpp:166, shaders/flame.glsl:27

    std::vector<flame_xform> xforms;
    std::array<std::array<float, 4>, 256> palette{};

    std::array<unsigned int, 2> size{};
    std::array<float, 2> center{};
    float scale = 0;
    float rotate = 0;

    std::optional<flame_xform> final_xform;

    int estimator_min = 0;
    int estimator_radius = 0;
    float estimator_curve = 0;

    float gamma = 0;
    float vibrancy = 0;
    float brightness = 0;

    template <typename Func>
    void for_each_xform(Func&& func) {
        int idx = 0;
        for (auto& x : xforms) func(idx++, x);
        if (final_xform) func(-1, final_xform.value());
    }

    // src/flame.hpp:77, src/flame.cpp:105-158. `seed` offsets the particle RNG seeds
    // (reference: seed = particle index, i.e. seed 0); shuffle_count is kept for
    // signature parity (the re-deal permutations are computed on chip).
    static void set_sim_parameters(std::size_t total_particles, std::size_t temporal_samples, std::size_t shuffle_count,
                                   std::uint64_t seed = 0);

    // src/flame.hpp:84, src/flame.cpp:160-226. nullptr on an unknown xform attribute,
    // unreadable file or a kernel that fails to compile; message in last_error().
    static std::unique_ptr<flame> load_flame(const std::string& path, const flame_compiler& vt);
    static std::unique_ptr<flame> load_flame_string(const std::string& xml_text, const std::string& origin, const flame_compiler& vt);
    static const std::string& last_error();

    bool needs_warmup() const;                                 // src/flame.hpp:86
    void warmup(std::size_t num_passes, float tss_width);      // src/flame.cpp:228-281
    // bins: DEVICE pointer to bins_len float4 (RGB + density), accumulated in place.
    // Returns the number of samples binned by this call (src/flame.cpp:283-330).
    std::size_t draw_to_bins(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter);
    // Same without the blocking counter read-back; collect with binned_total().
    void draw_to_bins_async(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter);
    std::uint64_t binned_total();  // all draw calls since the last warmup
    // Kernel option l2_hints: classify the 16 x 16-bin tiles of `bins` by the density accumulated so far and mark the densest
    // ones (at most budget_bytes of histogram; 0 = half the device's L2) as worth keeping in L2. Later draw calls into a
    // histogram of the same dimensions send the reductions of all other tiles with an evict-first hint.
    struct hot_map_info { int tiles_x = 0, tiles_y = 0; std::uint32_t hot_tiles = 0, threshold_bucket = 0; std::uint64_t budget_bytes = 0; };
    hot_map_info build_hot_map(const float* bins, std::size_t bins_len, std::size_t bins_width, std::uint64_t budget_bytes);
    void clear_hot_map();
    std::vector<std::uint32_t> copy_hot_map();  // one bit per tile, row-major

    // The reference's own dispatch structure on the GPU (one iteration per launch, particle / RNG state through global
    // memory, one xform per 256-thread workgroup, shuffle-buffer gather / scatter: flame.cpp:252-280, :317-325 driving
    // flame.glsl:41-90). A same-hardware baseline and a pass-level parity hook, not the product path. `shuffle_ids` holds
    // one (in, out) pair per pass — 1 + num_passes pairs for warmup, num_iter pairs for draw — or nullptr to draw them
    // from a seeded std::mt19937 (the reference seeds it from std::random_device).
    void reference_warmup(std::size_t num_passes, float tss_width, const std::uint32_t* shuffle_ids);
    std::size_t reference_draw_to_bins(float* bins, std::size_t bins_len, std::size_t bins_width, int num_iter, const std::uint32_t* shuffle_ids);
    // uploads `count` permutations of [0, particles per temporal sample) as the shuffle buffers of the reference mode
    // (binding 4); nullptr generates them on the device from `seed`
    static void set_shuffle_buffers(const std::uint32_t* host_tables, std::size_t count, std::uint64_t seed);
    void reset_animation();                                    // src/flame.cpp:332-336
    // Evaluates every <motion> entry at `time` (seconds) and writes base + amplitude * f(freq * time) into the animated fields;
    // the base values are the ones load_flame read (kept aside on the first call). Marks the flame as needing warmup.
    // Returns the number of fields written. No reference counterpart beyond the unused motion map (see motion_info).
    int apply_motion(float time);
    static float motion_function(const std::string& name, float x);  // sin / triangle / hill of flam3; unknown names: 0

    ~flame();

    // src/flame.hpp:97-128
    static flame_xform::affine_t rotate_affine(const flame_xform::affine_t& a, float deg);
    static flame_xform::affine_t scale_affine(const flame_xform::affine_t& a, float scale);
    static flame_xform::affine_t translate_affine(const flame_xform::affine_t& a, const std::array<float, 2>& t);
    // src/flame.cpp:289-296
    flame_xform::affine_t screen_space_affine(std::size_t bins_width, std::size_t bins_height) const;

    // src/flame.cpp:33-71 and :73-103
    void make_shader_buffer_map();
    std::array<float, PARAM_BUFFER> copy_flame_data_to_buffer() const;
    const buffer_map_t& buffer_map() const { return buffer_map_; }

    // generated text: reference-identical GLSL body and the CUDA translation unit
    const std::string& glsl_source() const { return glsl_source_; }
    const std::string& cuda_source() const { return cuda_source_; }
    const kernel_options& options() const { return options_; }
    // the translation unit of a kernel variant: `staged` compiles rfk_draw's region queues in (rfk_draw alone); `baked`, when
    // not null, holds the 4 * size constants of rfk_cfp[] and replaces every rfk_cfp[k] of the text with its value
    std::string variant_source(bool staged, const std::vector<float>* baked, bool pairs = false) const;
    bool pairs_allowed() const;  // the options admit the two-particles-per-thread build of the value-specialised kernels
    std::vector<char> variant_cubin(bool staged, const std::vector<float>* baked, bool pairs = false) const;
    std::vector<float> constant_table(const float* fp) const;  // the contents of rfk_cfp[] for the parameter buffer `fp`
    // Rebuilds the CUDA module with new options (the structure of the genome is fixed
    // after load, only values change without a rebuild: src/flame.hpp, main.cpp:335-369).
    bool set_options(const kernel_options& opt);
    // sm_100a cubin of the generated kernels (compiled on first use; needs no GPU)
    const std::vector<char>& cubin();

    flame_device* device() { return device_.get(); }
    std::unique_ptr<flame_device>& device_slot() { return device_; }
    void mark_dirty() { needs_update_ = true; }

private:
    flame();
    bool do_common_init(const flame_compiler& fc);  // src/flame.cpp:17-31
    friend class flame_compiler;
    friend struct flame_device;

    buffer_map_t buffer_map_;
    std::string glsl_source_;
    std::string cuda_body_;    // generated dispatch()/get_xform_id() in the CUDA dialect
    std::string cuda_source_;  // prelude + options + body + kernels
    kernel_options options_;
    std::vector<char> cubin_;
    std::unique_ptr<flame_device> device_;
    bool needs_update_ = true;
    std::vector<flame_xform> motion_base_;            // xforms (and the final xform last) as loaded, once apply_motion ran
    void rebuild_cuda_source();
};

void set_last_error(const std::string& msg);

}  // namespace rfk
