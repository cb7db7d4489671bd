// TEST INFRASTRUCTURE — the parity oracle. Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may build or call this; the product
// (refrakt_b200/) never does.
//
// CPU restatement of the reference's render path, one function per reference function:
//   shaders/include/random.glsl:21-41      -> load/save_random_state, ranval, randf
//   src/util.hpp:75-97                     -> jsf32_host_ranval, jsf32_warmup_ctx
//   src/hammersley.cpp:6-48                -> make_sample_points
//   src/shuffle_buffers.cpp:9-21           -> create_shuffle_buffer (seeded, see note)
//   shaders/templates/animate.tpl.glsl     -> animate
//   shaders/flame.glsl:41-90               -> iterate_thread / iterate_pass
//   src/flame.cpp:228-281, :283-330        -> warmup, draw_to_bins
//   shaders/density_vert.glsl:26-63 + density_frag.glsl:9-19 + src/main.cpp:490-515 -> density_estimate
//   shaders/tonemap.glsl:18-39             -> tonemap
// The genome-specific get_xform_id()/dispatch() are the reference compiler's own GLSL
// output (restated in refrakt_oracle.py) pasted ahead of this header by the generator,
// compiled against glsl_shim.hpp.
//
// PARITY PINNED against the reference's own code run here: the reference has no tests, golden vectors or fixtures
//  and needs a GL context, but its host sources compile unmodified against stand-in headers and a software GL
// that executes its GLSL text on the CPU (oracle/ref_host.cpp, oracle/softgl/, built by oracle/Makefile into oracle/_ref/).
// Replaying recorded runs of that build, this restatement reproduces RNG states, particle buffers, histogram and
// fp_inflated bit for bit and density estimation to 1e-4 (tests/test_reference_device_golden.py, fixtures under
// tests/golden/reference_device_*.npz). Caveat stated there: shaders run through g++ and glibc's libm, not a GLSL compiler.
//
// Where the reference leaves behaviour open, the oracle fixes it and says so:
//   * operand evaluation order of several randf() in one statement: textual order
//     (the generator hoists them);
//   * the clock-seeded std::mt19937_64 / random_device seeds: explicit seeds;
//   * the racy `bins[i] +=` (flame.glsl:82-84): every sample is added (no lost updates);
//   * ivec2(floor(NaN)) and uint(negative): non-finite positions never bin, negative
//     colour indexes clamp to 0;
//   * tonemap of an empty pixel (0 * 0/0): black.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "oracle_core_pre.hpp"

namespace oracle {

using glsl::uint;

// src/util.hpp:81-95 (host generator: returns d)
struct jsf32_ctx { uint32_t a, b, c, d; };
inline uint32_t jsf32_host_ranval(jsf32_ctx& x) {
    uint32_t e = x.a - glsl::rot32(x.b, 27);
    x.a = x.b ^ glsl::rot32(x.c, 17);
    x.b = x.c + x.d;
    x.c = x.d + e;
    x.d = e + x.a;
    return x.d;
}
inline void jsf32_warmup_ctx(jsf32_ctx& x, uint32_t seed) {
    x.a = 0xf1ea5eed;
    x.b = x.c = x.d = seed;
    for (int i = 0; i < 20; i++) (void)jsf32_host_ranval(x);
}

// src/hammersley.cpp:6-48
inline uint32_t reverse_bits32(uint32_t n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
    n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
    n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
    n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
    return n;
}
inline void make_sample_points(uint32_t count, float* out /* count x 4 */) {
    uint32_t max = count;
    if (count % 2 != 0) { uint32_t v = count; v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++; max = v; }
    float inv_max = 1.0f / max;
    unsigned int v = max;
    unsigned r = 0;
    while (v >>= 1) r++;
    for (uint32_t i = 0; i < count; i++) {
        uint32_t flipped = r ? (reverse_bits32(i) >> (32 - r)) : 0;
        out[4 * i + 0] = (float)((i * inv_max) * 2.0 - 1.0);
        out[4 * i + 1] = (float)((flipped * inv_max) * 2.0 - 1.0);
        out[4 * i + 2] = 0.0f;
        out[4 * i + 3] = 0.0f;
    }
}

// animate.tpl.glsl:18-96 for one temporal sample
struct xform_slots { int affine[6]; int rotation_frequency; };
inline void animate(const float* fp, float* fp_inflated, int size, int temporal_samples, float temporal_sample_width,
                    const xform_slots* xf, int nxf) {
    const float rads_per_second = 0.31415926535f;
    for (int g = 0; g < temporal_samples; g++) {
        int half_width = temporal_samples / 2;
        int sample_pos = g - half_width;
        float dt = sample_pos / float(half_width) * temporal_sample_width;
        float* dst = fp_inflated + (size_t)g * size;
        for (int i = 0; i < size; i++) dst[i] = fp[i];  // the template copies every slot it names; the rest is never read
        for (int k = 0; k < nxf; k++) {
            float sino = glsl::sin(rads_per_second * dt * fp[xf[k].rotation_frequency]);
            float coso = glsl::cos(rads_per_second * dt * fp[xf[k].rotation_frequency]);
            const int* a = xf[k].affine;
            dst[a[0]] = fp[a[0]] * coso + fp[a[2]] * sino;
            dst[a[1]] = fp[a[1]] * coso + fp[a[3]] * sino;
            dst[a[2]] = fp[a[2]] * coso - fp[a[0]] * sino;
            dst[a[3]] = fp[a[3]] * coso - fp[a[1]] * sino;
        }
    }
}

// flame.glsl:78-84: bin index or -1, and the palette index
inline int bin_index(float x, float y, float w, const float* ss_affine, int W, int H) {
    float px = glsl::fma(ss_affine[0], x, glsl::fma(ss_affine[2], y, ss_affine[4]));
    float py = glsl::fma(ss_affine[1], x, glsl::fma(ss_affine[3], y, ss_affine[5]));
    float fx = glsl::floor(px), fy = glsl::floor(py);
    if (!(fx >= -2147483648.0f && fx < 2147483648.0f && fy >= -2147483648.0f && fy < 2147483648.0f)) return -1;  // NaN / out of int range
    int cx = (int)fx, cy = (int)fy;
    if (cx >= 0 && cy >= 0 && cx < W && cy < H && w > 0) return (H - cy - 1) * W + cx;
    return -1;
}
inline int palette_index(float z) {
    float c = glsl::ceil(z * 255.0f);
    uint u = c > 0.0f ? (c < 4294967040.0f ? (uint)c : 0xffffffffu) : 0u;
    return (int)std::min(255u, u);
}

struct sim {
    size_t P = 0, TS = 0, PPT = 0, nshuf = 0;
    int block = 256;  // BLOCK_WIDTH, src/flame.cpp:15
    std::vector<uint32_t> shuffle;        // nshuf x PPT   (binding 4)
    std::vector<glsl::uvec4> rng;         // P             (binding 11)
    std::vector<float> samples;           // PPT x 4       (sample_buffer_)
    std::vector<float> local_buf, swap_buf;  // P x 4 each (local_buffer_, swap_buffer_)
    std::vector<float> fp_inflated;       // TS x size     (binding 9)
    std::vector<float> palette;           // 256 x 4       (binding 6)
    int total_params = 0;
    bool has_final = false;
    std::mt19937 pass_rng{0x5EED0001u};   // the per-pass shuffle ids (flame.cpp:231-233: random_device-seeded)
    unsigned long long counter = 0;       // flame_atomic_counters[0]
    unsigned long long xform_picks[64] = {0};
    std::vector<uint32_t> id_log;         // every (shuf_buf_idx_in, shuf_buf_idx_out) pair drawn, in order
    std::vector<uint32_t> forced_ids;     // when non-empty, the host loops take their shuffle ids from here (replaying a recorded run)
    size_t forced_pos = 0;
};

// src/flame.cpp:105-158
inline void set_sim_parameters(sim& S, size_t total_particles, size_t temporal_samples, size_t shuffle_count, uint64_t shuffle_seed, uint32_t rng_seed) {
    S.P = total_particles; S.TS = temporal_samples; S.nshuf = shuffle_count; S.PPT = total_particles / temporal_samples;
    S.shuffle.resize(S.nshuf * S.PPT);
    for (size_t k = 0; k < S.nshuf; k++) {  // create_shuffle_buffer, shuffle_buffers.cpp:9-21 (clock seed -> shuffle_seed + k)
        std::mt19937_64 generator{shuffle_seed + k};
        uint32_t* vec = S.shuffle.data() + k * S.PPT;
        for (uint32_t i = 0; i < S.PPT; i++) vec[i] = i;
        std::shuffle(vec, vec + S.PPT, generator);
    }
    S.rng.resize(S.P);
    for (uint32_t i = 0; i < S.P; i++) {
        jsf32_ctx c;
        jsf32_warmup_ctx(c, rng_seed + i);
        S.rng[i] = glsl::uvec4{c.a, c.b, c.c, c.d};
    }
    S.samples.resize(S.PPT * 4);
    make_sample_points((uint32_t)S.PPT, S.samples.data());
    S.local_buf.assign(S.P * 4, 0.0f);
    S.swap_buf.assign(S.P * 4, 0.0f);
}

struct pass_uniforms {
    bool random_read, random_write, first_run, do_draw;
    uint shuf_buf_idx_in, shuf_buf_idx_out;
    int bin_w, bin_h;
    float ss_affine[6];
};

// One dispatch of flame.glsl (grid PPT/block x TS): every workgroup in turn, its threads in turn.
// The shader's barriers only order the shared xid / fp writes before their reads, which a
// sequential walk over the group's threads preserves.
inline void iterate_pass(sim& S, const pass_uniforms& u, const float* pos_in, float* pos_out, float* bins /* may be per-thread */,
                         unsigned long long& binned, unsigned long long* picks) {
    using glsl::vec3; using glsl::vec4; using glsl::local_random_state; using glsl::randf; using glsl::PI;
    using glsl::get_xform_id; using glsl::dispatch;
    const size_t groups_x = S.PPT / S.block;
#pragma omp for schedule(static)
    for (long long ts = 0; ts < (long long)S.TS; ts++) {
        glsl::fp = S.fp_inflated.data() + (size_t)ts * S.total_params;  // flame.glsl:44-49
        glsl::first_run = u.first_run;
        for (size_t wg = 0; wg < groups_x; wg++) {
            const size_t base = (size_t)ts * S.PPT;  // gl_WorkGroupID.y * gl_WorkGroupSize.x * gl_NumWorkGroups.x
            int xid = 0;
            for (int lt = 0; lt < S.block; lt++) {
                const size_t gid = wg * S.block + lt;          // gl_GlobalInvocationID.x
                local_random_state = S.rng[base + gid];         // load_random_state()
                if (lt == 0) {                                  // flame.glsl:51-53
                    xid = get_xform_id(randf());
                    if (picks) picks[xid] += S.block;
                }
                uint i_idx = u.random_read ? S.shuffle[gid + S.PPT * u.shuf_buf_idx_in] : (uint)gid;
                uint o_idx = u.random_write ? S.shuffle[gid + S.PPT * u.shuf_buf_idx_out] : (uint)gid;
                size_t offset = u.first_run ? 0 : base;
                const float* ps = pos_in + 4 * (offset + i_idx);
                vec3 part_state(ps[0], ps[1], ps[2]);
                if (u.first_run) {
                    float r0 = randf();
                    float r1 = randf();
                    part_state.xy += r0 * .1f * PI * 2.0f * glsl::sincos(glsl::sqrt(r1));
                }
                vec4 result = dispatch(part_state, xid);
                float* po = pos_out + 4 * (base + o_idx);
                po[0] = result.x; po[1] = result.y; po[2] = result.z; po[3] = 0.0f;
                if (u.do_draw) {
                    if (S.has_final) result = dispatch(result.xyz, -1) * vec4(1.0f, 1.0f, 1.0f, result.w);
                    int idx = bin_index(result.x, result.y, result.w, u.ss_affine, u.bin_w, u.bin_h);
                    if (idx >= 0) {
                        const float* pal = S.palette.data() + 4 * palette_index(result.z);
                        float* b = bins + 4 * (size_t)idx;
                        b[0] += pal[0]; b[1] += pal[1]; b[2] += pal[2]; b[3] += result.w;
                        binned++;
                    }
                }
                S.rng[base + gid] = local_random_state;  // save_random_state()
            }
        }
    }
}

// src/main.cpp:490-515 + density_vert.glsl + density_frag.glsl, in the closed form of SURVEY Appendix D:
// bin (bx, by), cy = H-1-by, radius r: r == 0 -> out[cy][bx-1] += color; else for i, m in [-r, r],
// n(k) = 2k/(2r+1) + 1/(2r+1)^2, dist = n(i)^2 + n(m)^2 <= 1: out[cy+m][bx-1+i] += color*(1-dist)*0.63661977236/r^2
inline int estimator_radius_of(float density, int estimator_radius, int estimator_min, float estimator_curve) {
    float q = (float)estimator_radius / glsl::pow(density, estimator_curve);
    int r = (q >= (float)estimator_radius || q != q) ? estimator_radius : (q <= -1.0f ? (int)std::max(q, -2147483648.0f) : (int)q);
    return std::max(estimator_min, std::min(estimator_radius, r));
}
inline void density_estimate(const float* bins, float* out, int W, int H, int estimator_radius, int estimator_min, float estimator_curve) {
    std::memset(out, 0, sizeof(float) * 4 * (size_t)W * H);
    if (estimator_radius > 100) estimator_radius = 100;  // main.cpp:502
    for (int by = 0; by < H; by++)
        for (int bx = 0; bx < W; bx++) {  // gl_VertexID order
            const float* color = bins + 4 * ((size_t)by * W + bx);
            if (color[3] == 0.0f) continue;  // density_vert.glsl:30-33
            int radius = estimator_radius_of(color[3], estimator_radius, estimator_min, estimator_curve);
            int cy = H - 1 - by;
            if (radius == 0) {
                int x = bx - 1;
                if (x >= 0) { float* o = out + 4 * ((size_t)cy * W + x); for (int k = 0; k < 4; k++) o[k] += color[k]; }
                continue;
            }
            float S = (float)(2 * radius + 1);
            float bias = 1.0f / (S * S);
            float norm_factor = 0.63661977236f / float(radius * radius);
            for (int m = -radius; m <= radius; m++) {
                int y = cy + m;
                if (y < 0 || y >= H) continue;
                float nm = 2.0f * (float)m / S + bias;
                for (int i = -radius; i <= radius; i++) {
                    int x = bx - 1 + i;
                    if (x < 0 || x >= W) continue;
                    float ni = 2.0f * (float)i / S + bias;
                    float distance = ni * ni + nm * nm;
                    if (distance > 1) continue;
                    float wgt = (1 - distance) * norm_factor;
                    float* o = out + 4 * ((size_t)y * W + x);
                    for (int k = 0; k < 4; k++) o[k] += color[k] * wgt;
                }
            }
        }
}

// tonemap.glsl:18-39
inline void tonemap(const float* in, float* out, size_t count, float gamma, float scale_constant, float brightness, float vibrancy) {
    using glsl::log; using glsl::pow; using glsl::clamp; using glsl::mix;
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < (long long)count; p++) {
        float c[4] = {in[4 * p], in[4 * p + 1], in[4 * p + 2], in[4 * p + 3]};
        float* o = out + 4 * p;
        if (!(c[3] > 0.0f)) { o[0] = o[1] = o[2] = 0.0f; o[3] = 1.0f; continue; }
        float s = .5f * brightness * log(1.0f + c[3] * scale_constant) * 0.434294481903251827651128918916f / c[3];
        for (int k = 0; k < 4; k++) c[k] *= s;
        float inv_gamma = 1.0f / gamma;
        float z = pow(c[3], inv_gamma);
        float gamma_factor = z / c[3];
        for (int k = 0; k < 3; k++) o[k] = clamp(mix(pow(c[k], inv_gamma), gamma_factor * c[k], vibrancy), 0.0f, 1.0f);
        o[3] = 1.0f;
    }
}

}  // namespace oracle
