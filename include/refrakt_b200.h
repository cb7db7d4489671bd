/*
 * refrakt_b200 — C ABI of the B200-native fractal-flame render core.
 *
 * Drop-in boundary for the render path of untrioctium/refrakt. The reference has no
 * FFI layer: its path sits behind the C++ methods of `struct flame`
 * (src/flame.hpp:41-174), `struct flame_compiler` (src/variation_table.hpp:20-60) and
 * the density / tonemap block inlined in the main loop (src/main.cpp:490-535). Every
 * entry point below names the reference interface it replaces. INTEGRATION.md shows
 * the binding a refrakt maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only. Pointers named *_dev are CUDA device pointers on
 *    the current device; all others are host pointers.
 *  - Functions returning int return 0 on success and a negative RFK_E_* code on
 *    failure; functions returning a handle return NULL on failure. The message is
 *    available from rfk_last_error() (thread-local), mirroring the reference's
 *    "nullptr + message on stdout" convention (src/flame.cpp:196, :222-225).
 *  - Like the reference (one thread owning the GL context, class-static buffers),
 *    the library is not re-entrant: one host thread per process drives one device.
 *  - There is no CPU fallback. Calls that need the GPU fail with RFK_E_CUDA when no
 *    device is available; parsing and code generation work without one.
 */
#ifndef REFRAKT_B200_H
#define REFRAKT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFK_ABI_VERSION 4

enum {
    RFK_OK = 0,
    RFK_E_INVALID = -1, /* bad argument */
    RFK_E_CUDA = -2,    /* CUDA / NVRTC failure, or no device */
    RFK_E_STATE = -3,   /* call order: e.g. draw before warmup */
    RFK_E_NOTFOUND = -4 /* unknown name or index */
};

typedef struct rfk_compiler rfk_compiler; /* flame_compiler, src/variation_table.hpp:20 */
typedef struct rfk_flame rfk_flame;       /* flame, src/flame.hpp:41 */

int rfk_abi_version(void);
const char* rfk_last_error(void);

/* ---- device plumbing (replaces the GL context + storage_buffer<T> RAII of src/buffer_objects.hpp) ---- */
int rfk_set_device(int ordinal);
int rfk_set_stream(void* cuda_stream); /* cudaStream_t; NULL = default stream */
int rfk_synchronize(void);
void* rfk_device_alloc(size_t bytes);  /* storage_buffer<T>{n}, buffer_objects.hpp:11-22 */
int rfk_device_free(void* ptr_dev);
int rfk_device_zero(void* ptr_dev, size_t bytes);                         /* storage_buffer::zero_out, :39-41 */
int rfk_memcpy_to_device(void* dst_dev, const void* src, size_t bytes);   /* storage_buffer::update_all, :43-45 */
int rfk_memcpy_to_host(void* dst, const void* src_dev, size_t bytes);     /* storage_buffer::get_*, :24-37 */
uint64_t rfk_kernel_launch_count(void); /* kernels launched by this library since load */
/* Frees the device memory the library itself owns: the RNG states and shuffle tables of rfk_set_sim_parameters (class
 * statics in the reference, src/flame.hpp:150-156) and the frame buffers rfk_render_frame keeps between calls (4.5 GB after
 * one 15360 x 8640 frame). Flames stay valid; call rfk_set_sim_parameters and rfk_flame_warmup before drawing again. */
int rfk_release_buffers(void);

/* ---- global simulation parameters: flame::set_sim_parameters, src/flame.hpp:77, src/flame.cpp:105-158 ----
 * total_particles / temporal_samples must be a multiple of 256 (the reference dispatches
 * (P/TS)/256 workgroups, flame.cpp:263). Seeds one JSF32 state per particle slot with
 * warmup_ctx(state, seed + slot) (src/util.hpp:90-95; the reference uses seed 0), on the
 * device. shuffle_count is accepted for signature parity; the per-pass permutations are
 * computed on chip. Invalidates the particle buffers of every live flame. */
int rfk_set_sim_parameters(size_t total_particles, size_t temporal_samples, size_t shuffle_count, uint64_t seed);

/* ---- flame_compiler: src/variation_table.hpp:20-60, src/variation_table.cpp:182-265 ---- */
rfk_compiler* rfk_compiler_create(const char* variations_yaml_path); /* ctor, variation_table.cpp:182-215 */
rfk_compiler* rfk_compiler_create_from_text(const char* yaml_text);
int rfk_compiler_load_overlay(rfk_compiler* c, const char* yaml_path); /* extra / corrected definitions, same format */
void rfk_compiler_destroy(rfk_compiler* c);
int rfk_compiler_is_param(const rfk_compiler* c, const char* name);     /* variation_table.hpp:39 */
int rfk_compiler_is_variation(const rfk_compiler* c, const char* name); /* variation_table.hpp:40 */
int rfk_compiler_is_common(const rfk_compiler* c, const char* name);    /* variation_table.hpp:41 */
int rfk_compiler_variation_count(const rfk_compiler* c);                /* variations().size(), :47 */
const char* rfk_compiler_variation_name(const rfk_compiler* c, int index); /* alphabetical, as std::map iterates */
/* get_parameters_for_variation, :44 — names joined by '\n' into buf; returns the count or a negative code */
int rfk_compiler_get_parameters_for_variation(const rfk_compiler* c, const char* name, char* buf, size_t buf_len);
const char* rfk_compiler_param_owner(const rfk_compiler* c, const char* param); /* :43 */

/* ---- flame: load / destroy. flame::load_flame, src/flame.hpp:84, src/flame.cpp:160-226 ----
 * NULL on an unreadable file, an unknown xform attribute, or generated kernels that do
 * not compile (the reference: shader compile failure, flame.cpp:30). */
rfk_flame* rfk_flame_load(const char* path, const rfk_compiler* c);
rfk_flame* rfk_flame_load_string(const char* xml_text, const rfk_compiler* c);
void rfk_flame_destroy(rfk_flame* f); /* ~flame, src/flame.hpp:91-93 */

/* ---- flame: public genome fields, src/flame.hpp:44-60. Writing a field marks the flame as
 * needing warmup (the UI does this by hand, src/main.cpp:335-369, :397-409). ---- */
typedef struct rfk_flame_info {
    uint32_t size[2];
    float center[2];
    float scale, rotate;
    int32_t estimator_min, estimator_radius;
    float estimator_curve;
    float gamma, vibrancy, brightness;
    int32_t num_xforms;      /* read-only */
    int32_t has_final_xform; /* read-only */
    int32_t param_count;     /* read-only: buffer_map["size"], src/flame.cpp:68 */
} rfk_flame_info;

/* flame_xform, src/flame.hpp:21-37. index -1 is the final xform (for_each_xform, :62-68). */
typedef struct rfk_xform_info {
    float affine[6];
    int32_t has_post;
    float post[6];
    float weight, color, color_speed;
    float rotation_frequency, opacity;
    int32_t num_variations; /* read-only */
    int32_t num_params;     /* read-only */
} rfk_xform_info;

int rfk_flame_get_info(const rfk_flame* f, rfk_flame_info* out);
int rfk_flame_set_info(rfk_flame* f, const rfk_flame_info* in);
int rfk_flame_get_xform(const rfk_flame* f, int index, rfk_xform_info* out);
int rfk_flame_set_xform(rfk_flame* f, int index, const rfk_xform_info* in); /* has_post must keep its loaded value */
/* variations / var_param maps of an xform (std::map order = alphabetical) */
const char* rfk_flame_variation_name(const rfk_flame* f, int xform, int k);
const char* rfk_flame_param_name(const rfk_flame* f, int xform, int k);
int rfk_flame_get_variation(const rfk_flame* f, int xform, const char* name, float* out);
int rfk_flame_set_variation(rfk_flame* f, int xform, const char* name, float value); /* existing names only: structure is fixed after load */
int rfk_flame_get_param(const rfk_flame* f, int xform, const char* name, float* out);
int rfk_flame_set_param(rfk_flame* f, int xform, const char* name, float value);
int rfk_flame_get_palette(const rfk_flame* f, float* rgba_256x4);
int rfk_flame_set_palette(rfk_flame* f, const float* rgba_256x4);

/* ---- flame: parameter buffer and generated code ---- */
const char* rfk_flame_buffer_map_json(const rfk_flame* f); /* buffer_map_.dump(), src/flame.hpp:130-132 */
int rfk_flame_copy_params(const rfk_flame* f, float* out_1024); /* copy_flame_data_to_buffer, src/flame.cpp:73-103 */
const char* rfk_flame_glsl_source(const rfk_flame* f); /* flame_compiler::compile_flame_xforms, variation_table.cpp:217-265 */
const char* rfk_flame_cuda_source(const rfk_flame* f); /* the translation unit handed to NVRTC */
/* sm_100a cubin of the generated kernels; *size receives the byte count, buf may be NULL to query. Needs no GPU. */
int rfk_flame_get_cubin(rfk_flame* f, void* buf, size_t buf_len, size_t* size);
/* The further builds of the same kernels the library makes on first use: `staged` = rfk_draw with the region queues of
 * kernel option staged_bins compiled in; `specialised` = 1: the current parameter values compiled in (kernel option
 * specialize), 2: the same with two particles per thread (kernel option pair_particles). Translation unit and sm_100a
 * cubin; need no GPU. */
const char* rfk_flame_variant_source(rfk_flame* f, int staged, int specialised);
int rfk_flame_get_variant_cubin(rfk_flame* f, int staged, int specialised, void* buf, size_t buf_len, size_t* size);
int rfk_flame_uses_specialised(const rfk_flame* f); /* 1 when the last warmup chose the value-specialised kernels */
/* kernel option pair_particles: 0 = generic kernels in use, 1 = the specialised kernels hold two particles per thread, 2 = one;
 * probe_ms_out (optional) receives the measurement that decided it: ms of the probe launch with one / two particles per thread */
int rfk_flame_pair_particles_state(const rfk_flame* f, float probe_ms_out[2]);

/* Options of the generated kernels (no reference counterpart). Changing them rebuilds the module. */
typedef struct rfk_kernel_options {
    int32_t math_mode;      /* 0: libdevice functions, IEEE division and sqrt. 1 (default): 2-ulp division/sqrt, range-reduced SFU
                               sine/cosine, lg2/ex2 pow for small exponents - inside the 1e-5 single-step contract.
                               2: --use_fast_math (outside the contract) */
    int32_t fmad;           /* FMA contraction; default 1 */
    int32_t per_lane_xform; /* 1: every particle picks its own xform (divergent). default 0: one pick per warp + on-chip re-deal */
    int32_t warp_aggregate; /* 1: match_any de-duplication of same-bin updates inside a warp */
    int32_t deterministic;  /* 1: fixed-point integer accumulation, bit-identical histograms run to run */
    int32_t count_xforms;   /* 1: count xform selections (rfk_flame_xform_counts) */
    int32_t min_blocks;     /* __launch_bounds__ minBlocksPerSM; 0 = automatic (2048 / block_width, or 1536 / block_width when that spills), -1 = leave it to the compiler */
    int32_t block_width;    /* threads per CTA = particles per re-deal pool: 128, 256 (default, the reference's workgroup) or 512 */
    int32_t deal_period;    /* re-deal the CTA's particles across warps every n-th iteration; default 1 */
    int32_t l2_hints;       /* histograms several times larger than L2: reductions outside the hot map (rfk_flame_build_hot_map)
                               carry an L2 evict-first hint; default 0 */
    int32_t staged_bins;    /* histograms several times larger than L2: rfk_draw appends every sample as an 8-byte record to the
                               queue of its region (2^n consecutive bins) and a second kernel accumulates the queues region by
                               region, so the reductions meet in L2 instead of being DRAM read-modify-writes at random addresses.
                               Same samples, same histogram. -1 = automatic (default): on for histograms of 512 MiB or more, in
                               at most 64 regions of 2^22 bins (64 MB) or larger; 0 = off; n = 8..24 = always on with regions of
                               2^n bins (at most 64 regions; excludes deterministic, warp_aggregate and l2_hints, which also
                               switch the automatic mode off). Queue memory: 16 GiB at most (RFK_STAGE_MAX_BYTES) */
    int32_t specialize;     /* value-specialised kernels: every parameter slot that is the same for all temporal samples (weights,
                               variation amounts, parameters, colours; not the rotated affine coefficients) is compiled into
                               rfk_warm / rfk_draw as a literal. The reference compiles one shader per genome STRUCTURE and edits
                               values live (src/main.cpp:335-369); that generic build stays the one rfk_flame_load makes. 0 = generic
                               kernels only; 1 = warmup rebuilds the specialised kernels (NVRTC, 1-2 s) whenever a value changed;
                               2 = automatic (default): built the second time warmup runs with unchanged values. Same results to
                               rounding: constants fold at compile time with IEEE arithmetic */
    int32_t pair_particles; /* the value-specialised kernels hold two particles per thread (CTAs of block_width / 2 threads over the
                               same pool of block_width particles): one xform pick, one walk to its code, one re-deal key and one
                               barrier per two iterations. 1 (default) = measured: when the specialised kernels are built both forms
                               run a short warm-up on scratch copies of the particle and RNG buffers and the faster one is kept
                               (the shipped genome gains 4 %, a genome of twelve heavy xforms would lose 10 %); 2 = always,
                               0 = never. Ignored (one particle per thread) together with per_lane_xform, warp_aggregate,
                               deterministic, count_xforms, l2_hints, staged_bins, deal_period > 1 or an explicit min_blocks */
} rfk_kernel_options;
int rfk_flame_get_options(const rfk_flame* f, rfk_kernel_options* out);
int rfk_flame_set_options(rfk_flame* f, const rfk_kernel_options* in);
/* registers / static shared memory / resident CTAs per SM of "rfk_warm" or "rfk_draw" */
int rfk_flame_kernel_info(rfk_flame* f, const char* kernel, int* regs, int* smem_bytes, int* blocks_per_sm);

/* ---- flame: run. src/flame.hpp:86-89 ---- */
int rfk_flame_needs_warmup(const rfk_flame* f); /* flame::needs_warmup, :86 */
/* flame::warmup, src/flame.cpp:228-281: uploads the current field values, builds the per-temporal-sample
 * parameter blocks (animate.tpl.glsl), re-seeds the particles from the Hammersley set and runs one first-run
 * pass plus num_passes undrawn iterations. Does not clear any histogram. */
int rfk_flame_warmup(rfk_flame* f, size_t num_passes, float tss_width);
/* flame::draw_to_bins, src/flame.cpp:283-330: num_iter iterations of every particle, accumulated in place into
 * the caller's histogram bins_dev (bins_len float4 = RGB + density; height = bins_len / bins_width, :290).
 * Returns the number of samples binned by this call, or a negative code. Blocks on the counter read-back. */
int64_t rfk_flame_draw_to_bins(rfk_flame* f, float* bins_dev, size_t bins_len, size_t bins_width, int num_iter);
/* Same without the read-back; rfk_flame_binned_total() returns the count since the last warmup. */
int rfk_flame_draw_to_bins_async(rfk_flame* f, float* bins_dev, size_t bins_len, size_t bins_width, int num_iter);
int64_t rfk_flame_binned_total(rfk_flame* f);
/* Hot map for histograms much larger than the L2 (no reference counterpart; kernel option l2_hints). Cuts the histogram
 * into 16 x 16-bin tiles, sums the density accumulated so far per tile and marks the densest tiles, at most budget_bytes
 * of histogram (0 = half the device's L2), as worth keeping in L2. Later draw calls into a histogram of the same
 * dimensions then send the reductions of every other tile with an evict-first hint, so that the one-off misses stop
 * evicting the tiles most samples land in. Results are unchanged (it is a cache hint); call it after the first draw call
 * of a frame. rfk_render_frame does so by itself when the option is set and the histogram exceeds twice the L2. */
typedef struct rfk_hot_map_info {
    int32_t tiles_x, tiles_y;
    uint32_t hot_tiles;          /* tiles marked hot */
    uint32_t threshold_bucket;   /* density bucket (upper 13 bits of the binary32 tile sum) a tile must exceed */
    uint64_t budget_bytes;       /* the budget actually used */
} rfk_hot_map_info;
int rfk_flame_build_hot_map(rfk_flame* f, const float* bins_dev, size_t bins_len, size_t bins_width, uint64_t budget_bytes, rfk_hot_map_info* info_out);
int rfk_flame_clear_hot_map(rfk_flame* f);
/* copies the map (one bit per tile, row-major, tiles_x tiles per row) into out[0..n_words); returns the words it holds */
int64_t rfk_flame_copy_hot_map(rfk_flame* f, uint32_t* out, size_t n_words);
int rfk_flame_reset_animation(rfk_flame* f); /* flame::reset_animation, src/flame.hpp:95 */
/* per-xform selection counts since the last warmup (count_xforms option); the reference's read-out is
 * src/main.cpp:595-611 */
int rfk_flame_xform_counts(rfk_flame* f, uint64_t* out, int n);
/* screen-space affine of draw_to_bins, src/flame.cpp:289-296 */
int rfk_flame_screen_affine(const rfk_flame* f, size_t bins_width, size_t bins_height, float out[6]);
/* per-frame animation step of the main loop, src/main.cpp:383-395: rotates every xform with a non-zero
 * rotation_frequency by degrees * rotation_frequency */
int rfk_flame_rotate_xforms(rfk_flame* f, float degrees);

/* <motion> children of an xform: motion_info, src/flame.hpp:15-19, :36. The reference declares the map and never fills it
 * (its parser is commented out, src/flame.cpp:199-210). Parsed here as that block intends: one entry per animated attribute
 * of the element (a variation, a parameter, weight, color, color_speed, opacity), sharing its motion_frequency and
 * motion_function; the attribute's value is the amplitude. rfk_flame_apply_motion evaluates them as flam3 does,
 * field = loaded value + amplitude * f(frequency * time) with f = sin | triangle | hill, and marks the flame for warmup;
 * it returns the number of fields written. rfk_render --motion applies it per frame (time = frame / fps). */
typedef struct rfk_motion_info {
    float freq, amplitude;
    char function[16];
    char target[64];
} rfk_motion_info;
int rfk_flame_motion_count(const rfk_flame* f, int xform);
int rfk_flame_get_motion(const rfk_flame* f, int xform, int k, rfk_motion_info* out); /* std::map order = alphabetical by target */
int rfk_flame_apply_motion(rfk_flame* f, float time_seconds);
float rfk_motion_function(const char* name, float x);

/* ---- affine helpers: flame::rotate_affine / scale_affine / translate_affine, src/flame.hpp:97-128 ---- */
void rfk_rotate_affine(const float a[6], float deg, float out[6]);
void rfk_scale_affine(const float a[6], float scale, float out[6]);
void rfk_translate_affine(const float a[6], const float t[2], float out[6]);

/* ---- density estimation + tonemap: the block inlined at src/main.cpp:490-535 ----
 * estimator_radius is clamped to 100 (main.cpp:502); scale_constant is 10^-exponent with exponent 4 in the
 * reference (main.cpp:228, :528). Images are row-major, row 0 = screen row 0 (= PNG row 0). */
typedef struct rfk_post_params {
    int32_t estimator_radius, estimator_min;
    float estimator_curve;
    float gamma, brightness, vibrancy;
    float scale_constant;
} rfk_post_params;
int rfk_flame_post_params(const rfk_flame* f, rfk_post_params* out); /* the uniforms main.cpp:502-509, :526-531 would set */
/* density_vert.glsl + density_frag.glsl, main.cpp:490-515: histogram -> float4 image */
int rfk_density_estimate(const float* bins_dev, float* image_dev, size_t width, size_t height, const rfk_post_params* p);
/* tonemap.glsl, main.cpp:522-535: float4 image -> float4 image (alpha = 1); rgba8_dev optional */
int rfk_tonemap(const float* image_in_dev, float* image_out_dev, uint8_t* rgba8_dev, size_t width, size_t height, const rfk_post_params* p);
/* both in one kernel; either output may be NULL */
int rfk_density_tonemap(const float* bins_dev, float* image_out_dev, uint8_t* rgba8_dev, size_t width, size_t height, const rfk_post_params* p);
/* The same for the output rows [y0, y1) only, from a histogram SLAB: bins_rows_dev holds the histogram rows of the source
 * rows [src_y0, src_y1) (image row cy is histogram row height - 1 - cy, so the slab starts at histogram row
 * height - src_y1), which must cover [y0 - R, y1 + R) inside the image, R = max(estimator_radius, estimator_min); output
 * row cy is written to row cy - out_y0 of the output buffers. What every rank of rfk_render_frame_sharded runs on its rows. */
int rfk_density_tonemap_rows(const float* bins_rows_dev, float* image_out_dev, uint8_t* rgba8_dev, size_t width, size_t height, const rfk_post_params* p,
                             uint32_t y0, uint32_t y1, uint32_t src_y0, uint32_t src_y1, uint32_t out_y0);
/* 2x2 box average of a float4 image: out is out_width x out_height, in is twice that in each dimension */
int rfk_downsample2x(const float* image_in_dev, float* image_out_dev, size_t out_width, size_t out_height);
/* flam3-style spatial filter + supersample reduction (absent in the reference, main.cpp:195-196: SURVEY §8f item 2):
 * separable Gaussian exp(-2x^2) of support 1.5 and radius `filter_radius` output pixels (the genome's filter="..."),
 * in is (out_width*ss) x (out_height*ss). rfk_spatial_filter_taps returns the per-axis tap count and weights (<= 64). */
int rfk_spatial_downsample(const float* image_in_dev, float* image_out_dev, size_t out_width, size_t out_height, int supersample, float filter_radius);
int rfk_spatial_filter_taps(int supersample, float filter_radius, float* taps_out_64);

/* ---- seeding kernels (src/flame.cpp:105-158 moved to the device) ---- */
int rfk_seed_rng_states(uint32_t* states_dev, size_t count, uint32_t seed_base); /* jsf32::warmup_ctx, src/util.hpp:90-95 */
int rfk_make_sample_points(float* points_dev, uint32_t count);                   /* make_sample_points, src/hammersley.cpp:29-48 */
int rfk_make_shuffle_buffers(uint32_t* out_dev, uint32_t size, uint32_t count, uint64_t seed); /* src/shuffle_buffers.cpp:9-52 */
int rfk_copy_rng_states(uint32_t* out, size_t first, size_t count); /* the global states, to host */

/* ---- end to end with host buffers: what one iteration of the main loop does for a still
 * (src/main.cpp:403-535): warmup, draw_to_bins until `target_binned` samples are in the histogram
 * (the `accumulated < quality * W * H` rule of :411), density estimation, tonemap, read-back
 * (texture::get_pixels, src/buffer_objects.hpp:113-119). The histogram lives in library-owned
 * device memory. ---- */
typedef struct rfk_frame_request {
    uint32_t width, height;    /* histogram = image size (supersampling 1, main.cpp:195) */
    uint32_t warmup_passes;    /* 16, main.cpp:263 */
    uint32_t drawing_passes;   /* iterations per draw_to_bins call: 128, main.cpp:264 */
    float tss_width;           /* 1.2/60, main.cpp:237 */
    uint64_t target_binned;    /* stop once this many samples are binned (quality * W * H, main.cpp:411); 0 = use max_draw_calls only */
    uint32_t max_draw_calls;   /* upper bound on draw_to_bins calls; 0 = unlimited */
    float scale_constant_exp;  /* 4, main.cpp:228 */
    uint32_t supersample;      /* histogram = image size x supersample in each dimension (main.cpp:195-196); 0 or 1 = none.
                                  With supersample > 1 the tonemapped image is reduced by rfk_spatial_downsample */
    float filter_radius;       /* spatial filter radius in output pixels (the genome's filter="..."); used when supersample > 1 */
} rfk_frame_request;
typedef struct rfk_frame_stats {
    uint64_t iterations;   /* chaos-game iterations run (warmup excluded) */
    uint64_t binned;       /* samples that landed in the histogram */
    uint32_t draw_calls;
    float ms_warmup, ms_draw, ms_post, ms_readback; /* CUDA-event times of the stages */
} rfk_frame_stats;
int rfk_render_frame(rfk_flame* f, const rfk_frame_request* req, uint8_t* rgba8_out, float* image_out /* optional float4 */,
                     rfk_frame_stats* stats);

/* ---- several GPUs of one box, one process per GPU (the reference is single-GPU: no counterpart) ----
 * Independent particle streams per GPU (rfk_set_sim_parameters with disjoint seeds) into a private histogram; ONE exchange
 * step per frame: the histograms are summed. NCCL (loaded at run time; RFK_NCCL_LIBRARY overrides the soname) carries the
 * bootstrap, the barriers and the fallback data path; where the GPUs can map each other's memory (NVLink / NVSwitch, CUDA
 * IPC) the sum is a reduce-scatter over row slabs PULLED by the rank that owns the slab, and the finished rows are written
 * straight into rank 0's image (RFK_COMM_P2P=0 forces the NCCL path).
 *   rank 0: rfk_comm_unique_id(id); hand `id` to the other processes (file, pipe, MPI, torch.distributed ...);
 *   every rank: rfk_set_device(local_rank); rfk_comm_init(id, rank, world);  ...  rfk_comm_destroy(). */
#define RFK_COMM_ID_BYTES 128
int rfk_comm_unique_id(uint8_t id_out[RFK_COMM_ID_BYTES]);
int rfk_comm_init(const uint8_t id[RFK_COMM_ID_BYTES], int rank, int world); /* collective; at most 16 ranks */
int rfk_comm_destroy(void);
int rfk_comm_rank(void);
int rfk_comm_world(void);  /* 1 when no communicator is up */
int rfk_comm_p2p(void);    /* 1 once a sharded frame has mapped the peers' buffers, 0 = NCCL data path */
int rfk_comm_barrier(void); /* all ranks; blocks the host */
/* in-place sum of the per-rank histograms (bins_len float4 each) onto rank `root`, or onto every rank when root < 0 */
int rfk_comm_reduce_histogram(float* bins_dev, size_t bins_len, int root);
/* Row slabs (pure geometry, no GPU needed): rank r of `world` produces the output rows [y0, y1) of `height` — the first
 * height % world ranks take one row more — and a density estimation of radius `halo` reads the source rows [src_y0, src_y1) */
typedef struct rfk_row_slab { uint32_t y0, y1, src_y0, src_y1; } rfk_row_slab;
int rfk_comm_row_slab(uint32_t height, uint32_t halo, int rank, int world, rfk_row_slab* out);
/* One frame over all ranks (collective: every rank calls it with the same request). Strong scaling of rfk_render_frame:
 * every rank warms up its own particles, the ranks together draw until `target_binned` samples are in the histograms
 * (each rank the same number of passes, the last call shorter than drawing_passes so that no rank overshoots by a whole
 * call), the histograms are reduce-scattered over row slabs with an estimator-radius halo, every rank runs density
 * estimation + tonemap (+ the spatial filter when supersample > 1) on its rows, and the rows are gathered on rank 0.
 * want_rgba8 / want_image say which outputs rank 0 wants; the output pointers are used on rank 0 only. */
typedef struct rfk_sharded_request {
    rfk_frame_request frame;
    uint32_t want_rgba8, want_image;
} rfk_sharded_request;
typedef struct rfk_sharded_stats {
    uint64_t iterations_global; /* chaos-game iterations of all ranks (warmup excluded) */
    uint64_t binned_global;     /* samples in the summed histogram */
    uint64_t passes;            /* drawn passes of this rank */
    uint32_t draw_calls;        /* of this rank */
    uint32_t p2p;               /* 1 = peer-memory data path, 0 = NCCL */
    uint32_t y0, y1;            /* output rows this rank produced */
    float ms_warmup, ms_draw, ms_reduce, ms_post, ms_readback; /* CUDA-event times of this rank's stages (ms_readback: barrier + device-to-host copy) */
} rfk_sharded_stats;
int rfk_render_frame_sharded(rfk_flame* f, const rfk_sharded_request* req, uint8_t* rgba8_out, float* image_out, rfk_sharded_stats* stats);

/* ---- on-disk buffer cache in the reference's format (src/buffer_cache.hpp:12-60, src/buffer_cache.cpp:7-21):
 * <root>/cache/<type>/<group>/<name>.bin = size_t byte count + payload. refrakt caches its 1024 shuffle permutations
 * under shuffle/<particles per temporal sample>/ and its JSF32 states under rand_state/<total particles>/
 * (src/flame.cpp:112-148). The render path here never needs the cache (seeding runs on the device); these calls let it
 * produce buffers a refrakt build consumes and consume the ones refrakt cached, for seeded A/B runs. ---- */
int rfk_cache_write_buffer(const char* root, const char* type, const char* group, const void* data, size_t bytes,
                           const char* name_or_null, char* name_out, size_t name_out_len); /* buffer_group::write_buffer */
int64_t rfk_cache_read_buffer(const char* root, const char* type, const char* group, const char* name, void* out, size_t out_len); /* bytes, or the size needed when out is NULL */
int rfk_cache_list(const char* root, const char* type, const char* group, char* names_out, size_t names_out_len); /* cached_buffers: count; names joined by '\n' */
/* writes the current global JSF32 states (rand_state/<P>/) and `shuffle_count` device-generated permutations of
 * [0, P/TS) (shuffle/<P/TS>/) of the current simulation parameters */
int rfk_export_sim_cache(const char* root, uint64_t shuffle_seed);
/* replaces the global JSF32 states with cached ones (rand_state/<P>/<name>; name NULL = first in the directory) */
int rfk_import_rng_states(const char* root, const char* name_or_null);
int rfk_set_rng_states(const uint32_t* states, size_t first, size_t count); /* host (a, b, c, d) quadruples -> global states */

/* ---- output: the reference's screenshot (src/main.cpp:590-593, stbi_write_png of get_pixels()) ---- */
int rfk_write_png(const char* path, const uint8_t* rgba8, size_t width, size_t height); /* host pixels, rows top to bottom */
/* the float frame (rfk_render_frame's image_out; the reference keeps RGBA32F textures, main.cpp:218-219, and saves 8 bits):
 * OpenEXR, scanline, uncompressed, four 32-bit float channels */
int rfk_write_exr(const char* path, const float* rgba32f, size_t width, size_t height);

/* ---- reference pass mode: the reference's own dispatch structure on the GPU (flame.cpp:252-280, :317-325 driving
 * flame.glsl:41-90: ONE iteration per launch on a (PPT/256, TS) grid, particle and RNG state through global memory, one
 * xform per 256-thread workgroup, shuffle-buffer gather / scatter). Not the product path: a same-hardware baseline for
 * the register-resident kernels and a pass-level parity hook (fed the oracle's shuffle tables and pass ids it reproduces
 * the oracle's RNG states bit for bit). shuffle_ids: one (in, out) pair per pass — 1 + num_passes pairs for warmup,
 * num_iter pairs for draw — or NULL to draw them from a seeded std::mt19937. ---- */
int rfk_set_shuffle_buffers(const uint32_t* tables, size_t count, uint64_t seed); /* count x (P/TS) permutations; NULL = device-generated */
int rfk_flame_reference_warmup(rfk_flame* f, size_t num_passes, float tss_width, const uint32_t* shuffle_ids);
int64_t rfk_flame_reference_draw_to_bins(rfk_flame* f, float* bins_dev, size_t bins_len, size_t bins_width, int num_iter, const uint32_t* shuffle_ids);
int rfk_flame_copy_particles(rfk_flame* f, float* out_px4); /* the particle buffer (x, y, colour, 0), to host */

/* ---- test hooks: the generated device functions on host-supplied vectors ---- */
/* one dispatch(v, xid) per element (variation_table.cpp:222-263). xyz: n x 3 in, xid: n, rng: n x 4 in/out,
 * fp: 1024 floats or NULL for the flame's current values, out: n x 4 (x, y, colour, opacity) */
int rfk_flame_single_step(rfk_flame* f, int n, const float* xyz, const int* xid, uint32_t* rng, const float* fp, int first_run, float* out);
/* get_xform_id(ratio) per element (xform_select.tpl.glsl) */
int rfk_flame_select_xform(rfk_flame* f, int n, const float* ratio, const float* fp, int* out);
/* flame.glsl:78-84 per element: xyzw n x 4 (x, y, colour, opacity) -> bin index (-1 = rejected) and palette index */
int rfk_flame_bucket_index(rfk_flame* f, int n, const float* xyzw, const float ss_affine[6], int width, int height, int* idx_out, int* palette_out);
/* the macro grammar of the flame compiler: replace_macro / find_macros, src/util.cpp:6-23. Results are written to out
 * (NUL-terminated; find_macros joins the names with '\n' in std::set order); returns the length or a negative code */
int rfk_text_replace_macro(const char* str, const char* name, const char* value, char* out, size_t out_len);
int rfk_text_find_macros(const char* str, char* out, size_t out_len);
/* animate.tpl.glsl: the temporal_samples x param_count blocks, to host */
int rfk_flame_animate(rfk_flame* f, float tss_width, int temporal_samples, float* out);

#ifdef __cplusplus
}
#endif
#endif /* REFRAKT_B200_H */
