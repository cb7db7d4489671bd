// Genome-independent device kernels: RNG seeding, sample points, shuffle buffers,
// temporal-sample parameter inflation, density estimation + tonemap.
#include "static_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace rfk::kernels {

namespace {

__device__ __forceinline__ unsigned int rot32(unsigned int x, int k) { return (x << k) | (x >> (32 - k)); }

__global__ void seed_rng_kernel(uint4* states, size_t count, unsigned int seed_base) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned int seed = seed_base + (unsigned int)i;
    unsigned int a = 0xf1ea5eedu, b = seed, c = seed, d = seed;
    for (int k = 0; k < 20; k++) {
        unsigned int e = a - rot32(b, 27);
        a = b ^ rot32(c, 17);
        b = c + d;
        c = d + e;
        d = e + a;
    }
    states[i] = make_uint4(a, b, c, d);
}

__global__ void sample_points_kernel(float4* out, unsigned int count, int bits, float inv_max) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned int flipped = bits ? (__brev(i) >> (32 - bits)) : 0u;
    float fx = (float)i * inv_max;
    float fy = (float)flipped * inv_max;
    out[i] = make_float4((float)((double)fx * 2.0 - 1.0), (float)((double)fy * 2.0 - 1.0), 0.0f, 0.0f);
}

__device__ __forceinline__ unsigned int hash32(unsigned int h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// One permutation per blockIdx.y: a 4-round Feistel network on `bits` bits (a bijection
// of [0, 2^bits)) cycle-walked into [0, size). Replaces std::shuffle + mt19937_64.
__global__ void shuffle_kernel(unsigned int* out, unsigned int size, int bits, unsigned long long seed) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const unsigned int perm = blockIdx.y;
    const int half_lo = bits / 2, half_hi = bits - half_lo;
    const unsigned int mask_lo = (1u << half_lo) - 1u, mask_hi = (1u << half_hi) - 1u;
    const unsigned int k0 = hash32((unsigned int)seed ^ (perm * 0x9E3779B9u)), k1 = hash32((unsigned int)(seed >> 32) + perm);
    unsigned int v = i;
    do {
        unsigned int lo = v & mask_lo, hi = v >> half_lo;
        for (int r = 0; r < 4; r++) {
            // (lo: half_lo bits, hi: half_hi bits) -> (hi', lo') with widths swapped back every two rounds
            unsigned int f = hash32(hi ^ k0 ^ (r * 0x632BE5ABu)) + k1;
            unsigned int nlo = (lo ^ f) & mask_lo;
            unsigned int g = hash32(nlo + k1 * (r + 1u)) ^ k0;
            unsigned int nhi = (hi ^ g) & mask_hi;
            lo = nlo; hi = nhi;
        }
        v = (hi << half_lo) | lo;
    } while (v >= size);
    out[(size_t)perm * size + i] = v;
}

__global__ void animate_kernel(const float* __restrict__ fp, float* __restrict__ fp_inflated, int total_params, int temporal_samples,
                               float temporal_sample_width, const animate_xform* __restrict__ xf, int num_xforms) {
    const float rads_per_second = 0.31415926535f;  // animate.tpl.glsl:16
    int ts = blockIdx.x * blockDim.x + threadIdx.x;
    if (ts >= temporal_samples) return;
    int half_width = temporal_samples / 2;
    int sample_pos = ts - half_width;
    // one temporal sample: no motion blur (the reference always runs multiples of 32; 0 / 0 here would poison every affine)
    float dt = half_width ? sample_pos / float(half_width) * temporal_sample_width : 0.0f;
    float* dst = fp_inflated + (size_t)ts * total_params;
    for (int i = 0; i < total_params; i++) dst[i] = fp[i];
    for (int k = 0; k < num_xforms; k++) {
        const animate_xform x = xf[k];
        float ang = rads_per_second * dt * fp[x.rotation_frequency];
        float sino = sinf(ang), coso = cosf(ang);
        float a = fp[x.affine[0]], b = fp[x.affine[1]], c = fp[x.affine[2]], d = fp[x.affine[3]];
        dst[x.affine[0]] = a * coso + c * sino;
        dst[x.affine[1]] = b * coso + d * sino;
        dst[x.affine[2]] = c * coso - a * sino;
        dst[x.affine[3]] = d * coso - b * sino;
    }
}

__global__ void fixed_to_float_kernel(const unsigned long long* __restrict__ fixed, float4* bins, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double s = 1.0 / 16777216.0;
    float4 b = bins[i];
    b.x += (float)((double)fixed[4 * i + 0] * s);
    b.y += (float)((double)fixed[4 * i + 1] * s);
    b.z += (float)((double)fixed[4 * i + 2] * s);
    b.w += (float)((double)fixed[4 * i + 3] * s);
    bins[i] = b;
}

// ---- hot map (see static_kernels.cuh) ---------------------------------------------------------------------------------
// Density buckets: the upper 13 bits of the binary32 pattern of a non-negative float, i.e. 16 buckets per octave.
__device__ __forceinline__ unsigned int hot_bucket(float v) { return v > 0.0f ? (__float_as_uint(v) >> 19) : 0u; }

// one warp per tile: sum of the density channel over its 16 x 16 bins, and the population of its bucket
__global__ void hot_tile_sums_kernel(const float4* __restrict__ bins, int W, int H, int tiles_x, int n_tiles, float* __restrict__ sums, unsigned int* counts) {
    const int tile = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    const int row = (tile / tiles_x) * HOT_MAP_TILE + (lane >> 1), col0 = (tile % tiles_x) * HOT_MAP_TILE + (lane & 1) * 8;
    float sum = 0.0f;
    if (row < H)
        for (int k = 0; k < 8; k++)
            if (col0 + k < W) sum += __ldg(&bins[(size_t)row * W + col0 + k].w);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
        sums[tile] = sum;
        atomicAdd(&counts[hot_bucket(sum)], 1u);
    }
}

// densest buckets first until the budget is spent: tiles whose bucket is ABOVE the threshold are hot
__global__ void hot_threshold_kernel(unsigned int* counts, unsigned int budget_tiles) {
    unsigned int acc = 0;
    int b = HOT_MAP_BUCKETS - 1;
    for (; b >= 1; b--) {
        if (acc + counts[b] > budget_tiles) break;
        acc += counts[b];
    }
    counts[HOT_MAP_BUCKETS] = (unsigned int)b;  // 0 when every non-empty tile fits
    counts[HOT_MAP_BUCKETS + 1] = acc;
}

__global__ void hot_bitmap_kernel(const float* __restrict__ sums, int n_tiles, const unsigned int* __restrict__ counts, unsigned int* __restrict__ bitmap) {
    const int tile = (int)(blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    const unsigned int threshold = counts[HOT_MAP_BUCKETS];
    const bool hot = tile < n_tiles && hot_bucket(sums[tile]) > threshold;
    const unsigned int word = __ballot_sync(0xffffffffu, hot);
    if ((threadIdx.x & 31) == 0 && tile < ((n_tiles + 31) & ~31)) bitmap[tile >> 5] = word;
}

__global__ void downsample2x_kernel(const float4* __restrict__ in, float4* __restrict__ out, int W, int H) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float4* r0 = in + (size_t)(2 * y) * (2 * W) + 2 * x;
    const float4* r1 = r0 + 2 * W;
    float4 a = r0[0], b = r0[1], c = r1[0], d = r1[1];
    out[(size_t)y * W + x] = make_float4((a.x + b.x + c.x + d.x) * 0.25f, (a.y + b.y + c.y + d.y) * 0.25f,
                                         (a.z + b.z + c.z + d.z) * 0.25f, (a.w + b.w + c.w + d.w) * 0.25f);
}

struct filter_taps { float w[64]; int n; int offset; };

// Output rows [oy0, oy1) of the W x H image; `in` holds the rows [in_y0, ...) of the supersampled image (row slabs of a
// multi-GPU frame; in_y0 = 0, oy0 = 0, oy1 = H for the whole image); output row y goes to row y - oy0 of `out`.
__global__ void spatial_downsample_kernel(const float4* __restrict__ in, float4* __restrict__ out, int W, int H, int ss, const __grid_constant__ filter_taps t,
                                          int oy0, int oy1, int in_y0) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = oy0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= oy1) return;
    const int IW = W * ss, IH = H * ss;
    const int x0 = x * ss - t.offset, y0 = y * ss - t.offset;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.0f;
    for (int j = 0; j < t.n; j++) {
        int sy = y0 + j;
        if (sy < 0 || sy >= IH) continue;
        for (int i = 0; i < t.n; i++) {
            int sx = x0 + i;
            if (sx < 0 || sx >= IW) continue;
            float w = t.w[i] * t.w[j];
            float4 v = __ldg(in + (size_t)(sy - in_y0) * IW + sx);
            acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
            wsum += w;
        }
    }
    float inv = wsum > 0.0f ? 1.0f / wsum : 0.0f;
    out[(size_t)(y - oy0) * W + x] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// Sum of the same `count` float4 bins of up to 16 histograms — this GPU's and its peers', mapped over NVLink — in the fixed
// order src[0] + src[1] + ...: the reduce-scatter of a multi-GPU frame, pulled by the rank that owns the slab
// (csrc/comm.cpp). Every thread keeps four 128-bit loads per source in flight.
struct slab_sources { const float4* src[16]; int n; };
__global__ void __launch_bounds__(256) slab_reduce_kernel(const __grid_constant__ slab_sources srcs, size_t offset, size_t count, float4* __restrict__ out) {
    // a CTA walks 16 KB pieces (4 x 256 float4, contiguous), grid-strided: every thread has four 128-bit loads per source in
    // flight and a warp's requests stay inside one 2 KB run of a peer's memory (pieces 4.8 MB apart per thread ran the 2.12 GB
    // histogram at 250 GB/s per peer)
    constexpr size_t kPiece = 4 * 256;
    for (size_t base = (size_t)blockIdx.x * kPiece; base < count; base += (size_t)gridDim.x * kPiece) {
        float4 a[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const size_t i = base + u * 256 + threadIdx.x;
            ok[u] = i < count;
            a[u] = ok[u] ? srcs.src[0][offset + i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int k = 1; k < srcs.n; k++) {
            float4 b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t i = base + u * 256 + threadIdx.x;
                b[u] = ok[u] ? srcs.src[k][offset + i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { a[u].x += b[u].x; a[u].y += b[u].y; a[u].z += b[u].z; a[u].w += b[u].w; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) out[base + u * 256 + threadIdx.x] = a[u];
    }
}

__global__ void pack_rgba8_kernel(const float4* __restrict__ in, uchar4* __restrict__ out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float4 v = in[i];
    out[i] = make_uchar4((unsigned char)rintf(fminf(fmaxf(v.x, 0.f), 1.f) * 255.0f), (unsigned char)rintf(fminf(fmaxf(v.y, 0.f), 1.f) * 255.0f),
                         (unsigned char)rintf(fminf(fmaxf(v.z, 0.f), 1.f) * 255.0f), (unsigned char)rintf(fminf(fmaxf(v.w, 0.f), 1.f) * 255.0f));
}

// density_vert.glsl:35-44, as the reference writes it (estimator_curve <= 0 only: the thresholds below need a radius that
// does not grow with the density)
__device__ __forceinline__ int estimator_radius_pow(float density, const density_params& p) {
    int r = (int)((float)p.estimator_radius / powf(density, p.estimator_curve));
    r = min(p.estimator_radius, r);
    return max(p.estimator_min, r);
}

// x^g for x > 0 through lg2 / ex2 (relative error ~1e-6 for the exponents 1/gamma used here; the outputs are
// colours in [0, 1] compared at 1e-4 / 1 LSB)
__device__ __forceinline__ float pow_pos(float x, float g) { return exp2f(g * __log2f(x)); }

// tonemap.glsl:25-36; an empty pixel (alpha 0) is 0 * (0/0) in the reference — defined as black here
__device__ __forceinline__ float4 tonemap_pixel(float4 color, const density_params& p) {
    if (!(color.w > 0.0f)) return make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    float s = .5f * p.brightness * logf(1.0f + color.w * p.scale_constant) * 0.434294481903251827651128918916f / color.w;  // (the SFU log loses the small arguments)
    color.x *= s; color.y *= s; color.z *= s; color.w *= s;
    float inv_gamma = 1.0f / p.gamma;
    float z = pow_pos(color.w, inv_gamma);
    float gamma_factor = z / color.w;
    float v = p.vibrancy;
    auto chan = [&](float ch) {
        // mix(pow(ch, 1/gamma), gamma_factor * ch, vibrancy); at vibrancy 1 the first term is multiplied by zero
        float m = v == 1.0f ? gamma_factor * ch : (ch > 0.0f ? pow_pos(ch, inv_gamma) : 0.0f) * (1.0f - v) + (gamma_factor * ch) * v;
        return fminf(fmaxf(m, 0.0f), 1.0f);
    };
    return make_float4(chan(color.x), chan(color.y), chan(color.z), 1.0f);
}

constexpr int DE_TILE_W = 32, DE_ROWS_PER_WARP = 4, DE_WARPS = 8, DE_TILE_H = DE_ROWS_PER_WARP * DE_WARPS;
constexpr int DE_THREADS = DE_WARPS * 32;
constexpr int DE_CAP = 512;        // candidate-list entries per batch (32 B each: 16 KB)
constexpr int DE_SMALL_R = 3;      // radius classes: 1..3 ("small": the warp's band of source rows is +-3) and above

// Packed FP32 (sm_100 FFMA2: two lanes per issue slot; the kernel is issue bound): acc.xy += c.xy * w, acc.zw += c.zw * w
__device__ __forceinline__ void fma4(float4& acc, const float4& c, float w) {
    const float2 w2 = make_float2(w, w);
    const float2 lo = __ffma2_rn(make_float2(c.x, c.y), w2, make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(make_float2(c.z, c.w), w2, make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Gather form of the reference's point-sprite splat. Source bin (bx, cy) with radius r lands on
// out[cy + m][bx - 1 + i], i, m in [-r, r], weight (1 - n(i)^2 - n(m)^2) * (2/pi) / r^2 when that is >= 0,
// n(k) = 2k/(2r+1) + 1/(2r+1)^2; r == 0 copies the bin to out[cy][bx-1]. |n(k)| > 1 for every k
// outside [-r, r], so max(weight, 0) alone restricts a candidate to its (2r+1)^2 footprint.
//
// A CTA produces a 32 x 32 output tile; a warp owns 4 rows x 32 columns and keeps their float4 sums in
// registers (no atomics, fixed summation order => deterministic).
//  A0. The CTA reads the densities of its (32 + 2R)^2 source window once and stores one radius byte per bin in shared
//      memory, 0 = not a candidate: the radius is the number of thresholds T[1] > T[2] > ... the density stays under
//      (host table, density_thresholds(): exactly the densities where the reference's int(R / pow(d, curve)) steps), so
//      the dense bulk of an image costs one 4-byte load and one compare per scanned bin; a candidate has radius >= 1
//      and a footprint that reaches the tile.
//  A1-A2. Candidates are counted per (class, row, 32-column chunk) — class "small" = radius <= 3, "large" above — and
//      prefix-summed: the list is class-major, row-major inside a class.
//  A3. The list is written to shared memory in batches of DE_CAP entries: colour premultiplied by (2/pi)/r^2, and
//      (bx - 1, cy, 2/S, 1/S^2), S = 2r + 1.
//  B.  Every warp walks, per class, the contiguous list range whose source rows can reach its 4 rows — +-3 rows for the
//      small class, +-R for the large one — and accumulates; lanes are output columns.
// Multi-GPU row slabs (rfk_comm_*): `bins` then holds the source rows [src_y0, src_y1) only and the launch produces the
// output rows [y0, y1); rows outside [src_y0, src_y1) count as empty (they are outside the image, or not needed).
// MAX_CHUNKS: compile-time bound of the 32-column chunks per window row, (32 + 2R + 31) / 32 — 2 up to radius 16 (the usual
// 9..11), 4 up to 48, 8 up to 100: the window scan's loops over the chunks unroll to exactly that many loads
template <bool DENSITY, bool TONEMAP, int MAX_CHUNKS = 8, int MIN_BLOCKS = 5>
__global__ void __launch_bounds__(DE_THREADS, MIN_BLOCKS) density_tonemap_kernel(const float4* __restrict__ bins, float4* __restrict__ out_f4,
                                                                     uchar4* __restrict__ out_rgba8, const __grid_constant__ density_params p) {
    __shared__ int s_part[DE_THREADS];
    __shared__ float4 s_list[2 * DE_CAP];  // entry e: [2e] = (bx - 1, cy, 2/S, 1/S^2), [2e + 1] = colour * (2/pi)/r^2
    extern __shared__ __align__(16) unsigned char s_dyn[];  // [nrows][pitch] radius bytes, then 2 * ncnt + 1 list offsets (unsigned short)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * DE_TILE_W, ty0 = p.y0 + blockIdx.y * DE_TILE_H;
    const int ox = tx0 + lane;
    const int oy0 = ty0 + warp * DE_ROWS_PER_WARP;
    const int W = p.W;
    // bin (bx, cy) of the histogram: row flip of flame.glsl:84, relative to the rows `bins` holds
    auto bin_at = [&](int bx, int cy) -> const float4* { return bins + (size_t)(p.src_y1 - 1 - cy) * W + bx; };

    float4 acc[DE_ROWS_PER_WARP];
#pragma unroll
    for (int k = 0; k < DE_ROWS_PER_WARP; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (DENSITY) {
        const int R = max(p.estimator_radius, p.estimator_min);
        // radius-0 sources: out[cy][ox] takes bin (ox + 1, cy). Loaded after the window scan (whose registers it would
        // otherwise occupy; the scan has just pulled these sectors into L1 / L2).
        auto load_radius0 = [&](const unsigned char* rad, int pitch_) {
#pragma unroll
            for (int k = 0; k < DE_ROWS_PER_WARP; k++) {
                const int cy = oy0 + k, bx = ox + 1;
                // a bin of the tile's own rows and columns with a radius >= 1 is not a radius-0 source (its footprint always reaches the tile)
                const bool splat = rad && rad[(R + warp * DE_ROWS_PER_WARP + k) * pitch_ + R + lane] != 0;
                if (cy >= p.src_y0 && cy < p.src_y1 && cy < p.y1 && bx < W && !splat) acc[k] = __ldg(bin_at(bx, cy));
            }
        };
        if (R < 1) load_radius0(nullptr, 0);
        if (R >= 1) {
            const int win_x0 = tx0 + 1 - R, win_w = DE_TILE_W + 2 * R, nch = (win_w + 31) >> 5, pitch = nch << 5;
            const int win_y0 = ty0 - R, nrows = DE_TILE_H + 2 * R;
            const int ncnt = nrows * nch;
            unsigned char* const s_rad = s_dyn;
            unsigned short* const s_off = reinterpret_cast<unsigned short*>(s_dyn + ((nrows * pitch + 15) & ~15));

            // A0 + A1: radius byte of every window bin and the candidate counts of its (row, chunk) cell, per class. Lean on
            // purpose — this loop runs for every tile of the image, candidates or not: column validity is one bit per chunk
            // computed up front, a row is one 64-bit address, the shared-memory stores go through window addresses.
            const float t_any = p.thresholds[1];
            const unsigned int rad_addr = (unsigned int)__cvta_generic_to_shared(s_rad), off_addr = (unsigned int)__cvta_generic_to_shared(s_off);
            unsigned int colmask = 0;
            for (int c = 0; c < nch; c++) {
                const int col = (c << 5) + lane, bx = win_x0 + col;
                if (col < win_w && bx >= 0 && bx < W) colmask |= 1u << c;
            }
            // radius bytes and counters start at zero; only cells with a candidate are written below
            {
                const unsigned int words16 = (unsigned int)((((nrows * pitch + 15) & ~15) + (2 * ncnt + 2) * 2 + 15) >> 4);
                for (unsigned int i = tid; i < words16; i += DE_THREADS)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(rad_addr + (i << 4)), "r"(0u) : "memory");
            }
            __syncthreads();
            bool any_candidate = false;
            // two window rows per trip, all their loads issued before the first is looked at: the scan is a chain of
            // load -> vote -> branch, and with one load in flight per warp its latency was a tenth of the kernel
            constexpr int kMaxChunks = MAX_CHUNKS;
            for (int row0 = warp; row0 < nrows; row0 += 2 * DE_WARPS) {
                float dv[2][kMaxChunks];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int row = row0 + rr * DE_WARPS, cy = win_y0 + row;
                    const bool row_ok = row < nrows && cy >= p.src_y0 && cy < p.src_y1;  // else outside the image or the slab: empty
                    const float* row_w = &bin_at(win_x0 + lane, row_ok ? cy : p.src_y1 - 1)->w;  // density of this lane's bin in chunk 0
#pragma unroll
                    for (int c = 0; c < kMaxChunks; c++) {
                        dv[rr][c] = 0.0f;
                        if (c < nch && row_ok && ((colmask >> c) & 1u)) dv[rr][c] = __ldg(row_w + c * 128);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int row = row0 + rr * DE_WARPS, cy = win_y0 + row;
#pragma unroll
                    for (int c = 0; c < kMaxChunks; c++) {
                        if (c >= nch) break;
                        const float d = dv[rr][c];
                        const bool cand = d != 0.0f && (p.use_pow || d <= t_any);
                        if (!__any_sync(0xffffffffu, cand)) continue;  // the common case: dense or empty bins only
                        int r = 0;
                        if (p.use_pow) {
                            if (cand) r = estimator_radius_pow(d, p);
                        } else {
                            r = cand ? 1 : 0;
                            for (int k = 2; k <= R; k++) {  // warp-uniform trip count: until no lane's density is under T[k]
                                const bool under = cand && d <= p.thresholds[k];
                                if (!__any_sync(0xffffffffu, under)) break;
                                r += under ? 1 : 0;
                            }
                        }
                        const int bx = win_x0 + (c << 5) + lane;
                        if (r >= 1 && (bx - 1 + r < tx0 || bx - 1 - r > tx0 + DE_TILE_W - 1 || cy + r < ty0 || cy - r > ty0 + DE_TILE_H - 1)) r = 0;
                        const unsigned int small = __ballot_sync(0xffffffffu, r != 0 && r <= DE_SMALL_R), large = __ballot_sync(0xffffffffu, r > DE_SMALL_R);
                        if ((small | large) == 0) continue;
                        any_candidate = true;
                        const unsigned int cell = (unsigned int)(row * nch + c);
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(rad_addr + (cell << 5) + lane), "r"(r) : "memory");
                        if (lane == 0) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(off_addr + 2u * cell), "h"((unsigned short)__popc(small)) : "memory");
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(off_addr + 2u * (cell + (unsigned int)ncnt)), "h"((unsigned short)__popc(large)) : "memory");
                        }
                    }
                }
            }
            // no candidate anywhere in the window (the dense interior of an image, and its empty surroundings): the sums are the
            // radius-0 copies already loaded
            if (__syncthreads_or(any_candidate ? 1 : 0) == 0) {
                load_radius0(nullptr, 0);
                goto epilogue;
            }
            load_radius0(s_rad, pitch);
            // A2: exclusive prefix sum over the 2 * ncnt counters (a window holds fewer than 65536 bins)
            const int n2 = 2 * ncnt;
            {
                const int per = (n2 + DE_THREADS - 1) / DE_THREADS;
                const int lo = min(tid * per, n2), hi = min(lo + per, n2);
                int sum = 0;
                for (int i = lo; i < hi; i++) sum += s_off[i];
                s_part[tid] = sum;
                __syncthreads();
                if (warp == 0) {
                    int carry = 0;
                    for (int base = 0; base < DE_THREADS; base += 32) {
                        int v = s_part[base + lane], incl = v;
                        for (int o = 1; o < 32; o <<= 1) {
                            int t = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += t;
                        }
                        s_part[base + lane] = carry + incl - v;
                        carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (lane == 0) s_off[n2] = (unsigned short)carry;
                }
                __syncthreads();
                int run = s_part[tid];
                for (int i = lo; i < hi; i++) {
                    int v = s_off[i];
                    s_off[i] = (unsigned short)run;
                    run += v;
                }
                __syncthreads();
            }
            const int total = s_off[n2];
            if (total > 0) {  // CTA-uniform
                // list ranges of this warp, per class: the source rows that can reach its output rows, as window rows
                int w_lo[2], w_hi[2];
#pragma unroll
                for (int cls = 0; cls < 2; cls++) {
                    const int reach = cls == 0 ? min(R, DE_SMALL_R) : R;
                    const int rlo = max(0, R + warp * DE_ROWS_PER_WARP - reach), rhi = min(nrows - 1, R + warp * DE_ROWS_PER_WARP + DE_ROWS_PER_WARP - 1 + reach);
                    w_lo[cls] = s_off[cls * ncnt + rlo * nch];
                    w_hi[cls] = s_off[cls * ncnt + (rhi + 1) * nch];
                }
                const float oxf = (float)ox, oy0f = (float)oy0;

                for (int base = 0; base < total; base += DE_CAP) {
                    // A3: write this batch of the list. A warp takes a contiguous share of the (class, row, chunk) cells; its lanes look
                    // at 32 cells at a time and only the cells with entries in this batch are walked.
                    {
                        const int share = (n2 + DE_WARPS - 1) / DE_WARPS, c_lo = warp * share, c_hi = min(n2, c_lo + share);
                        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                            const int mine_cell = c0 + lane;
                            int first = 0, last = 0;
                            if (mine_cell < c_hi) { first = s_off[mine_cell]; last = s_off[mine_cell + 1]; }
                            unsigned int todo = __ballot_sync(0xffffffffu, last > first && last > base && first < base + DE_CAP);
                            while (todo) {
                                const int src = __ffs(todo) - 1;
                                todo &= todo - 1;
                                const int i = c0 + src, cell_first = __shfl_sync(0xffffffffu, first, src);
                                const bool large = i >= ncnt;
                                const int cell = large ? i - ncnt : i;
                                const int r = s_rad[(cell << 5) + lane];
                                const bool mine = large ? r > DE_SMALL_R : (r != 0 && r <= DE_SMALL_R);
                                const unsigned int vote = __ballot_sync(0xffffffffu, mine);
                                if (mine) {
                                    const int e = cell_first + __popc(vote & ((1u << lane) - 1u)) - base;
                                    if (e >= 0 && e < DE_CAP) {
                                        const int row = cell / nch, c = cell - row * nch;
                                        const int cy = win_y0 + row, bx = win_x0 + (c << 5) + lane;
                                        float4 col = __ldg(bin_at(bx, cy));
                                        const float S = (float)(2 * r + 1), norm = 0.63661977236f / (float)(r * r);  // density_vert.glsl:62
                                        col.x *= norm; col.y *= norm; col.z *= norm; col.w *= norm;
                                        s_list[2 * e] = make_float4((float)(bx - 1), (float)cy, 2.0f / S, 1.0f / (S * S));
                                        s_list[2 * e + 1] = col;
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                    // B: accumulate. weight / norm = 1 - n(i)^2 - n(m)^2, kept when >= 0 (density_frag.glsl:17 discards distance > 1)
#pragma unroll
                    for (int cls = 0; cls < 2; cls++) {
                        const int e_lo = max(w_lo[cls], base) - base, e_hi = min(w_hi[cls], base + DE_CAP) - base;
                        // the entries by shared-window address: one register walks the list, both reads are [reg + immediate]
                        const unsigned int list0 = (unsigned int)__cvta_generic_to_shared(s_list);
                        for (unsigned int at = list0 + 32u * (unsigned int)max(e_lo, 0), end = list0 + 32u * (unsigned int)max(e_hi, 0); at < end; at += 32u) {
                            float4 g;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "r"(at));
                            // n(m) of every row from m = output row - source row itself (exact in binary32), not from the row's
                            // offset inside the tile: a pixel's sum must not depend on where the tiling starts (row slabs)
                            const float m0 = oy0f - g.y;
                            const float nm0 = fmaf(m0, g.z, g.w);                                        // n(m) of the warp's first row
                            const float nm3 = fmaf(m0 + (float)(DE_ROWS_PER_WARP - 1), g.z, g.w);        // ... and of its last one
                            if (nm0 > 1.0f || nm3 < -1.0f) continue;  // no row of this warp inside the footprint (warp-uniform)
                            float4 col;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(col.x), "=f"(col.y), "=f"(col.z), "=f"(col.w) : "r"(at));
                            const float ni = fmaf(oxf - g.x, g.z, g.w);
                            const float one_minus_ci = fmaf(-ni, ni, 1.0f);
#pragma unroll
                            for (int q = 0; q < DE_ROWS_PER_WARP; q++) {
                                const float nm = q == 0 ? nm0 : (q == DE_ROWS_PER_WARP - 1 ? nm3 : fmaf(m0 + (float)q, g.z, g.w));
                                fma4(acc[q], col, fmaxf(fmaf(-nm, nm, one_minus_ci), 0.0f));
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < DE_ROWS_PER_WARP; k++) {
            int cy = oy0 + k;
            if (cy < p.y1 && ox < W) acc[k] = __ldg(bins + (size_t)(cy - p.src_y0) * W + ox);  // tonemap only: input is already an image
        }
    }

epilogue:
    if (ox >= W) return;
#pragma unroll
    for (int k = 0; k < DE_ROWS_PER_WARP; k++) {
        int cy = oy0 + k;
        if (cy >= p.y1) continue;
        float4 v = TONEMAP ? tonemap_pixel(acc[k], p) : acc[k];
        size_t o = (size_t)(cy - p.out_y0) * W + ox;
        if (out_f4) out_f4[o] = v;
        if (out_rgba8) out_rgba8[o] = make_uchar4((unsigned char)rintf(fminf(fmaxf(v.x, 0.f), 1.f) * 255.0f), (unsigned char)rintf(fminf(fmaxf(v.y, 0.f), 1.f) * 255.0f),
                                                  (unsigned char)rintf(fminf(fmaxf(v.z, 0.f), 1.f) * 255.0f), (unsigned char)rintf(fminf(fmaxf(v.w, 0.f), 1.f) * 255.0f));
    }
}

}  // namespace

void seed_rng_states(uint4* states, std::size_t count, std::uint32_t seed_base, cudaStream_t s) {
    if (!count) return;
    seed_rng_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(states, count, seed_base);
}

void make_sample_points(float4* out, std::uint32_t count, cudaStream_t s) {
    if (!count) return;
    // src/hammersley.cpp:33-42
    std::uint32_t max = count;
    if (count % 2 != 0) { max = count - 1; max |= max >> 1; max |= max >> 2; max |= max >> 4; max |= max >> 8; max |= max >> 16; max++; }
    float inv_max = 1.0f / max;
    int bits = 0;
    for (std::uint32_t v = max; v >>= 1;) bits++;
    sample_points_kernel<<<(count + 255) / 256, 256, 0, s>>>(out, count, bits, inv_max);
}

void make_shuffle_buffers(std::uint32_t* out, std::uint32_t size, std::uint32_t count, std::uint64_t seed, cudaStream_t s) {
    if (!size || !count) return;
    int bits = 2;
    while ((1ull << bits) < size) bits++;
    dim3 grid((size + 255) / 256, count);
    shuffle_kernel<<<grid, 256, 0, s>>>(out, size, bits, seed);
}

void animate(const float* fp, float* fp_inflated, int total_params, int temporal_samples, float temporal_sample_width,
             const animate_xform* xforms_dev, int num_xforms, cudaStream_t s) {
    animate_kernel<<<(temporal_samples + 31) / 32, 32, 0, s>>>(fp, fp_inflated, total_params, temporal_samples, temporal_sample_width, xforms_dev, num_xforms);
}

// radius >= k  <=>  density <= thresholds[k], for k = 1 .. R = max(estimator_radius, estimator_min): the exact steps of the
// reference's  max(estimator_min, min(estimator_radius, int(estimator_radius / pow(d, curve))))  (density_vert.glsl:35-44)
// in binary32 with this host's powf, found by bisection on the bit pattern (positive floats order like their bits).
// Needs a radius that does not grow with the density: estimator_curve > 0; returns false otherwise (the kernel then
// evaluates the formula per bin).
static bool density_thresholds(float* host_out, int estimator_radius, int estimator_min, float estimator_curve) {
    const int R = estimator_radius > estimator_min ? estimator_radius : estimator_min;
    host_out[0] = INFINITY;
    for (int k = 1; k <= R && k < 102; k++) host_out[k] = INFINITY;
    if (!(estimator_curve > 0.0f)) return false;
    auto radius_of = [&](float d) {
        // int(x) of the shader saturates on a GPU; in C++ the conversion of a quotient beyond INT_MAX (tiny densities) is undefined
        const float q = (float)estimator_radius / powf(d, estimator_curve);
        int r = q >= 2147483648.0f ? estimator_radius : (q > -2147483648.0f ? (int)q : estimator_min);
        r = r < estimator_radius ? r : estimator_radius;
        return r > estimator_min ? r : estimator_min;
    };
    for (int k = 1; k <= R && k < 102; k++) {
        if (k <= estimator_min) { host_out[k] = INFINITY; continue; }
        // largest finite positive float d with radius_of(d) >= k; radius_of is non-increasing in d
        std::uint32_t lo = 1u, hi = 0x7f7fffffu;  // smallest subnormal .. FLT_MAX
        auto as_float = [](std::uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; };
        if (radius_of(as_float(lo)) < k) { host_out[k] = 0.0f; continue; }
        if (radius_of(as_float(hi)) >= k) { host_out[k] = INFINITY; continue; }
        while (hi - lo > 1) {
            const std::uint32_t mid = lo + (hi - lo) / 2;
            if (radius_of(as_float(mid)) >= k) lo = mid; else hi = mid;
        }
        host_out[k] = as_float(lo);
    }
    return true;
}

void density_tonemap(const float4* bins, float4* out_f4, uchar4* out_rgba8, density_params p, bool do_density, bool do_tonemap, cudaStream_t s) {
    if (p.estimator_radius > 100) p.estimator_radius = 100;  // main.cpp:502
    if (p.estimator_radius < 0) p.estimator_radius = 0;
    if (p.estimator_min < 0) p.estimator_min = 0;
    if (p.estimator_min > 100) throw std::invalid_argument("density estimation: estimator_min above 100 (the radius table and the source window hold 100)");
    if (p.y1 == 0 && p.y0 == 0) { p.y1 = p.H; p.src_y0 = 0; p.src_y1 = p.H; p.out_y0 = 0; }  // the whole image
    if (p.y0 < 0 || p.y1 > p.H || p.y0 >= p.y1 || p.src_y0 < 0 || p.src_y1 > p.H || p.src_y0 >= p.src_y1)
        throw std::invalid_argument("density estimation: bad row range");
    p.use_pow = density_thresholds(p.thresholds, p.estimator_radius, p.estimator_min, p.estimator_curve) ? 0 : 1;
    dim3 grid((p.W + DE_TILE_W - 1) / DE_TILE_W, (p.y1 - p.y0 + DE_TILE_H - 1) / DE_TILE_H);
    dim3 block(DE_THREADS);
    // dynamic shared memory: radius bytes of the (32 + 2R)^2 source window + 2 list offsets per (row, 32-column chunk):
    // 3.9 KB at R = 11, 67 KB at the maximum R = 100
    const int R = p.estimator_radius > p.estimator_min ? p.estimator_radius : p.estimator_min;
    const size_t nrows = DE_TILE_H + 2 * R, nch = (DE_TILE_W + 2 * R + 31) >> 5;
    const size_t dyn_bytes = do_density && R >= 1 ? (((nrows * (nch << 5) + 15) & ~size_t(15)) + (2 * nrows * nch + 2) * sizeof(unsigned short) + 15) & ~size_t(15) : 0;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev >= 0 && dev < 64 && !configured[dev]) {  // per device: the opt-in is a property of the function on that device
        cudaFuncSetAttribute(density_tonemap_kernel<true, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);  // radius above 48 only
        cudaFuncSetAttribute(density_tonemap_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        configured[dev] = true;
    }
    if (do_density && do_tonemap) {
        if (nch <= 2) density_tonemap_kernel<true, true, 2><<<grid, block, dyn_bytes, s>>>(bins, out_f4, out_rgba8, p);
        else if (nch <= 4) density_tonemap_kernel<true, true, 4><<<grid, block, dyn_bytes, s>>>(bins, out_f4, out_rgba8, p);
        else density_tonemap_kernel<true, true, 8><<<grid, block, dyn_bytes, s>>>(bins, out_f4, out_rgba8, p);
    }
    else if (do_density) density_tonemap_kernel<true, false><<<grid, block, dyn_bytes, s>>>(bins, out_f4, out_rgba8, p);
    else density_tonemap_kernel<false, true><<<grid, block, 0, s>>>(bins, out_f4, out_rgba8, p);
}

void build_hot_map(const float4* bins, int W, int H, unsigned int budget_tiles, float* tile_sums, unsigned int* scratch, unsigned int* bitmap, cudaStream_t s) {
    const int tiles_x = (W + HOT_MAP_TILE - 1) / HOT_MAP_TILE, tiles_y = (H + HOT_MAP_TILE - 1) / HOT_MAP_TILE, n_tiles = tiles_x * tiles_y;
    if (n_tiles <= 0) return;
    cudaMemsetAsync(scratch, 0, HOT_MAP_SCRATCH_WORDS * sizeof(unsigned int), s);
    hot_tile_sums_kernel<<<(unsigned)(((size_t)n_tiles * 32 + 255) / 256), 256, 0, s>>>(bins, W, H, tiles_x, n_tiles, tile_sums, scratch);
    hot_threshold_kernel<<<1, 1, 0, s>>>(scratch, budget_tiles);
    const int padded = (n_tiles + 31) & ~31;
    hot_bitmap_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, s>>>(tile_sums, n_tiles, scratch, bitmap);
}

// Kernel option staged_bins, second half: rfk_draw left every sample of the call as an 8-byte record in the queue of its
// region (2^region_shift consecutive bins). A queue is `cursors[region]` chunks of 512 records, chunk c holding fill[c]
// of them. Grid = (CTAs per region, regions): the block scheduler hands out blockIdx.x fastest, so with four 512-thread
// CTAs resident per SM the whole GPU works on at most two or three regions at any time and their bins (32 MB each by
// default) stay in L2 while their samples stream in — every reduction is an L2 hit, DRAM sees each touched sector once
// in and once out. The records are read with the streaming (evict-first) policy so that they do not push the bins out.
constexpr unsigned int kStageChunk = 512;  // RFK_STAGE_CHUNK of chaos_kernels.cuh
__global__ void __launch_bounds__(kStageChunk) stage_accumulate_kernel(const uint2* __restrict__ records, const unsigned int* __restrict__ cursors,
                                                                       const unsigned int* __restrict__ fill, unsigned int capacity, int region_shift,
                                                                       const float4* __restrict__ palette, float4* bins, size_t nbins) {
    __shared__ float4 pal[256];
    if (threadIdx.x < 256) pal[threadIdx.x] = palette[threadIdx.x];
    __syncthreads();
    const unsigned int region = blockIdx.y;
    const size_t region_base = (size_t)region << region_shift;
    const unsigned int chunks = min(cursors[region], capacity);
    const unsigned int* region_fill = fill + (size_t)region * capacity;
    const uint2* queue = records + (size_t)region * capacity * kStageChunk;
    auto add = [&](uint2 rec) {
        const size_t idx = region_base + (rec.x >> 8);
        if (idx >= nbins) return;  // cannot happen for records rfk_draw wrote
        const float4 col = pal[rec.x & 255u];
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(bins + idx), "f"(col.x), "f"(col.y), "f"(col.z), "f"(__uint_as_float(rec.y)) : "memory");
    };
    // four chunks in flight per thread: their fill counts first, then the records, then the reductions
    constexpr int kInFlight = 4;
    for (unsigned int c0 = blockIdx.x; c0 < chunks; c0 += kInFlight * gridDim.x) {
        unsigned int n[kInFlight];
        uint2 rec[kInFlight];
#pragma unroll
        for (int j = 0; j < kInFlight; j++) {
            const unsigned int c = c0 + j * gridDim.x;
            n[j] = c < chunks ? min(__ldg(region_fill + c), kStageChunk) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kInFlight; j++)
            if (threadIdx.x < n[j]) rec[j] = __ldcs(queue + (size_t)(c0 + j * gridDim.x) * kStageChunk + threadIdx.x);
#pragma unroll
        for (int j = 0; j < kInFlight; j++)
            if (threadIdx.x < n[j]) add(rec[j]);
    }
}

void stage_accumulate(const uint2* records, const unsigned int* cursors, const unsigned int* fill, unsigned int capacity, int region_shift, int regions,
                      const float4* palette, float4* bins, std::size_t nbins, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dim3 grid((unsigned)sms * 4, (unsigned)regions);  // four 512-thread CTAs per SM: one region at a time on the whole GPU
    stage_accumulate_kernel<<<grid, kStageChunk, 0, s>>>(records, cursors, fill, capacity, region_shift, palette, bins, nbins);
}

void fixed_to_float(const unsigned long long* fixed, float4* bins, std::size_t count, cudaStream_t s) {
    if (!count) return;
    fixed_to_float_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(fixed, bins, count);
}

int spatial_filter_taps(int ss, float filter_radius, float* taps_out) {
    const double support = 1.5;
    const double fw = 2.0 * support * ss * (double)filter_radius;
    int fwidth = (int)fw + 1;
    if ((fwidth ^ ss) & 1) fwidth++;  // same parity as the supersample factor: the taps centre on the output pixel
    if (fwidth > 64) fwidth = 64 - ((64 ^ ss) & 1);
    const double adjust = fw > 0.0 ? support * fwidth / fw : 1.0;
    double sum = 0.0, w[64];
    for (int i = 0; i < fwidth; i++) {
        double x = ((2.0 * i + 1.0) / fwidth - 1.0) * adjust;
        w[i] = std::exp(-2.0 * x * x) * std::sqrt(2.0 / 3.14159265358979323846);
        sum += w[i];
    }
    if (taps_out) for (int i = 0; i < fwidth; i++) taps_out[i] = (float)(w[i] / sum);
    return fwidth;
}

void spatial_downsample(const float4* in, float4* out, int W, int H, int ss, float filter_radius, cudaStream_t s) {
    spatial_downsample_rows(in, out, W, H, ss, filter_radius, 0, H, 0, s);
}

void spatial_downsample_rows(const float4* in, float4* out, int W, int H, int ss, float filter_radius, int oy0, int oy1, int in_y0, cudaStream_t s) {
    if (oy1 <= oy0) return;
    filter_taps t{};
    t.n = spatial_filter_taps(ss, filter_radius, t.w);
    t.offset = (t.n - ss) / 2;
    dim3 block(32, 8), grid((W + 31) / 32, (oy1 - oy0 + 7) / 8);
    spatial_downsample_kernel<<<grid, block, 0, s>>>(in, out, W, H, ss, t, oy0, oy1, in_y0);
}

void spatial_filter_rows(int ss, float filter_radius, int oy0, int oy1, int full_height, int* in_y0, int* in_y1) {
    filter_taps t{};
    t.n = spatial_filter_taps(ss, filter_radius, t.w);
    t.offset = (t.n - ss) / 2;
    const int lo = oy0 * ss - t.offset, hi = (oy1 - 1) * ss - t.offset + t.n;
    *in_y0 = lo < 0 ? 0 : lo;
    *in_y1 = hi > full_height ? full_height : hi;
}

void slab_reduce(const float4* const* sources, int n_sources, std::size_t offset, std::size_t count, float4* out, cudaStream_t s) {
    if (!count || n_sources <= 0) return;
    if (n_sources > 16) throw std::invalid_argument("slab_reduce: more than 16 sources");
    slab_sources srcs{};
    srcs.n = n_sources;
    for (int k = 0; k < n_sources; k++) srcs.src[k] = sources[k];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const std::size_t want = (count + 4 * 256 - 1) / (4 * 256);
    const unsigned grid = (unsigned)std::min<std::size_t>(want, (std::size_t)sms * 8);  // eight 256-thread CTAs per SM: a multiple of the SM count
    slab_reduce_kernel<<<grid, 256, 0, s>>>(srcs, offset, count, out);
}

void pack_rgba8(const float4* in, uchar4* out, std::size_t count, cudaStream_t s) {
    if (!count) return;
    pack_rgba8_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(in, out, count);
}

void downsample2x(const float4* in, float4* out, int W, int H, cudaStream_t s) {
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    downsample2x_kernel<<<grid, block, 0, s>>>(in, out, W, H);
}

}  // namespace rfk::kernels
