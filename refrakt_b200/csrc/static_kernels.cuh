// Launchers of the genome-independent kernels (static_kernels.cu, built by nvcc for sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace rfk::kernels {

// src/flame.cpp:132-149 + src/util.hpp:90-95: states[i] = warmup_ctx(seed_base + i)
void seed_rng_states(uint4* states, std::size_t count, std::uint32_t seed_base, cudaStream_t s);

// src/hammersley.cpp:29-48 as a kernel (test/ABI surface; the chaos kernel computes its point inline)
void make_sample_points(float4* out, std::uint32_t count, cudaStream_t s);

// src/shuffle_buffers.cpp:9-52 as a kernel: `count` permutations of [0, size), reproducible from `seed`
void make_shuffle_buffers(std::uint32_t* out, std::uint32_t size, std::uint32_t count, std::uint64_t seed, cudaStream_t s);

struct animate_xform { int affine[6]; int rotation_frequency; };
// shaders/templates/animate.tpl.glsl:18-96: per-temporal-sample parameter blocks
void animate(const float* fp, float* fp_inflated, int total_params, int temporal_samples, float temporal_sample_width,
             const animate_xform* xforms_dev, int num_xforms, cudaStream_t s);

struct density_params {
    int W, H;
    int estimator_radius, estimator_min;  // density_vert.glsl:3-4 (radius already clamped to <= 100, main.cpp:502)
    float estimator_curve;
    float gamma, brightness, vibrancy, scale_constant;  // tonemap.glsl:13-16 (scale_constant = 10^-4, main.cpp:528)
    // radius >= k  <=>  density <= thresholds[k] (exact steps of the reference's formula); filled by density_tonemap()
    float thresholds[102];
    int use_pow;           // estimator_curve <= 0: no thresholds, the kernel evaluates the formula per bin; filled by density_tonemap()
    // Row slab of a multi-GPU frame (all zero = the whole image): the launch produces output rows [y0, y1); `bins` holds
    // the histogram rows of the source rows [src_y0, src_y1) only (source row cy = histogram row H - 1 - cy, so bins[0] is
    // the bin (0, src_y1 - 1)); output row cy is written to row cy - out_y0 of the output buffers.
    int y0, y1, src_y0, src_y1, out_y0;
};
// Density estimation (density_vert.glsl:26-63 + density_frag.glsl:9-19 + main.cpp:490-515) and/or tonemap
// (tonemap.glsl:18-39) in one pass. out_f4 / out_rgba8 may each be null.
void density_tonemap(const float4* bins, float4* out_f4, uchar4* out_rgba8, density_params p, bool do_density, bool do_tonemap, cudaStream_t s);

// Hot map of a histogram much larger than L2 (kernel option l2_hints; DESIGN.md "hot map"): the histogram is cut into
// 16 x 16-bin tiles (4 KB each), every tile's accumulated density is summed, and the densest tiles — as many as fit in
// `budget_tiles` — get their bit set in `bitmap` (one bit per tile, row-major, tiles_x = ceil(W/16) tiles per row).
// `tile_sums` holds ceil(W/16) * ceil(H/16) floats, `scratch` HOT_MAP_SCRATCH_WORDS words, `bitmap` one word per 32 tiles.
// After the call scratch[HOT_MAP_BUCKETS] is the density-bucket threshold and scratch[HOT_MAP_BUCKETS + 1] the number of
// hot tiles.
constexpr int HOT_MAP_TILE = 16;
constexpr int HOT_MAP_BUCKETS = 4096;
constexpr int HOT_MAP_SCRATCH_WORDS = HOT_MAP_BUCKETS + 2;
void build_hot_map(const float4* bins, int W, int H, unsigned int budget_tiles, float* tile_sums, unsigned int* scratch, unsigned int* bitmap, cudaStream_t s);

// deterministic mode: bins += fixed * 2^-24, one thread per bin
void fixed_to_float(const unsigned long long* fixed, float4* bins, std::size_t count, cudaStream_t s);
// kernel option staged_bins: adds the queued samples of one draw call to the histogram, region by region
void stage_accumulate(const uint2* records, const unsigned int* cursors, const unsigned int* fill, unsigned int capacity, int region_shift, int regions,
                      const float4* palette, float4* bins, std::size_t nbins, cudaStream_t s);

// 2x2 box average of a float4 image (config 3: 2x supersampled histogram -> image); W, H are the OUTPUT dims
void downsample2x(const float4* in, float4* out, int W, int H, cudaStream_t s);

// flam3-style spatial filter + supersample reduction (the reference has none: main.cpp:195-196).
// Separable Gaussian g(x) = exp(-2 x^2) sqrt(2/pi) of support 1.5, width fw = 2 * 1.5 * ss * filter_radius supersampled
// pixels rounded up to fwidth taps of the parity of ss, weights normalised to 1. The tap table (<= 64 per axis) is
// returned in `taps_out` (host, may be null). in is (W*ss) x (H*ss), out is W x H; taps outside the image are dropped
// and the weights of the remaining ones renormalised.
int spatial_filter_taps(int ss, float filter_radius, float* taps_out);
// float4 in [0, 1] -> RGBA8, round to nearest (the GL UNORM conversion of buffer_objects.hpp:113-119)
void pack_rgba8(const float4* in, uchar4* out, std::size_t count, cudaStream_t s);
void spatial_downsample(const float4* in, float4* out, int W, int H, int ss, float filter_radius, cudaStream_t s);
// the same for the output rows [oy0, oy1) only, written to rows 0 .. oy1 - oy0 of `out`; `in` starts at row in_y0 of the
// supersampled image (row slabs of a multi-GPU frame). spatial_filter_rows: the supersampled rows [in_y0, in_y1) those
// output rows read.
void spatial_downsample_rows(const float4* in, float4* out, int W, int H, int ss, float filter_radius, int oy0, int oy1, int in_y0, cudaStream_t s);
void spatial_filter_rows(int ss, float filter_radius, int oy0, int oy1, int full_height, int* in_y0, int* in_y1);
// out[i] = sources[0][offset + i] + sources[1][offset + i] + ... for i < count (n_sources <= 16; the sources may be peer
// memory): the pulled reduce-scatter of a multi-GPU frame
void slab_reduce(const float4* const* sources, int n_sources, std::size_t offset, std::size_t count, float4* out, cudaStream_t s);

}  // namespace rfk::kernels
