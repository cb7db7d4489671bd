// Chaos-game kernels for one genome (sm_100a, compiled by NVRTC together with the
// generated dispatch()/get_xform_id()). Replaces shaders/flame.glsl:41-90 and the
// per-pass dispatch loops of src/flame.cpp:252-280 and :317-325.
//
// The reference runs ONE iteration per dispatch and moves the particle, its RNG
// state, two shuffle indices and a histogram read-modify-write through global memory
// every time (~104 B per iteration). Here a thread keeps its particle and its RNG in
// registers for all `num_iter` iterations of a call; HBM sees one float4 + one uint4
// load and store per particle per call, plus the histogram reductions.
//
// Divergence: the reference picks one xform per 256-thread workgroup per pass and
// relies on its shuffle-buffer permutations to mix particles between workgroups.
// Same idea at warp granularity: lane 0 of every warp draws the xform for the warp
// (no divergence inside dispatch()), and after every iteration the CTA re-deals its
// particles across warps through shared memory with a fresh bijection of [0, BLOCK),
// computed on chip (the shuffle buffers of src/shuffle_buffers.cpp, moved on chip).
//
// Expects these macros from the host (flame_device.cpp): RFK_BLOCK, RFK_LOG2_BLOCK,
// RFK_TOTAL_PARAMS, RFK_NUM_XFORMS, RFK_HAS_FINAL, RFK_LAUNCH_BOUNDS and the
// option macros RFK_PER_LANE_XFORM, RFK_WARP_AGGREGATE, RFK_DETERMINISTIC,
// RFK_COUNT_XFORMS, RFK_L2_HINTS, RFK_STAGED_BINS (each 0 or 1).

using namespace rfk_glsl;

// This CTA's copy of the parameter block: fp[] (flame.glsl:44-48) and the rotated affine coefficients as one float4 per
// xform (device_prelude.cuh). `src` is a row of fp_inflated, or fp itself. The caller synchronises.
__device__ __forceinline__ void rfk_stage_params(const float* __restrict__ src) {
    for (int i = threadIdx.x; i < RFK_TOTAL_PARAMS; i += blockDim.x) rfk_glsl::fp[i] = src[i];
    for (int i = threadIdx.x; i <= RFK_NUM_XFORMS; i += blockDim.x) {
        const int s = rfk_affine_slot[i];
        if (s >= 0) rfk_aff[i] = make_float4(src[s], src[s + 1], src[s + 2], src[s + 3]);
    }
}

// dispatch(v, xform) of the generated text (variation_table.cpp:222-263) with the picked xform's rotated coefficients
template <bool first_run>
__device__ __forceinline__ vec4 dispatch(vec3 v, int xform, rfk_rng& rs) {
    return dispatch_a<first_run>(v, xform, rs, rfk_aff[xform + 1]);
}

#ifndef RFK_EXPERIMENT
#define RFK_EXPERIMENT 0  // timing builds of tools/gpu_probe_l1tex.sh (paired rfk_draw only)
#endif
struct rfk_iter_params {
    float4* particles;                  // [P] (x, y, colour, 0): pos_in/pos_out of buffers.glsl:1-9
    uint4* rng;                         // [P] JSF32 state per thread slot (random.glsl:1-4)
    const float* fp_inflated;           // [TS * RFK_TOTAL_PARAMS] (buffers.glsl:26-29)
    const float4* palette;              // [256] (buffers.glsl:16-19)
    float4* bins;                       // [W * H] RGB + density (buffers.glsl:21-24)
    unsigned long long* fixed_bins;     // [W * H * 4] fixed-point accumulators (deterministic mode)
    unsigned long long* counters;       // [0] binned samples (buffers.glsl:31-34), [1 + i] picks of xform i
    float ss_affine[6];                 // flame.glsl:22
    int bin_w, bin_h;                   // flame.glsl:17
    float bin_wf, bin_hf;               // the same as binary32 (exact: both are below 2^24)
    int num_iter;
    int ppt;                            // particles per temporal sample
    int first_run;                      // flame.glsl:13
    unsigned int deal_seed;             // key of this call's re-deal permutations
    int hammersley_bits;                // log2 of the sample-point count (src/hammersley.cpp:37-42)
    float hammersley_inv_max;
    const unsigned int* hot_map;        // RFK_L2_HINTS: one bit per 16 x 16-bin tile, set = keep in L2; null = no hints
    int hot_tiles_x;                    // tiles per histogram row
    // RFK_STAGED_BINS: samples are appended to per-region queues in HBM and accumulated region by region afterwards
    uint2* stage_records;               // [regions][stage_capacity chunks][RFK_STAGE_CHUNK] (local bin << 8 | palette row, opacity bits)
    unsigned int* stage_cursors;        // [regions] chunks handed out so far (may run past the capacity)
    unsigned int* stage_fill;           // [regions][stage_capacity] records written into each chunk
    unsigned int stage_capacity;        // chunks per region
    int stage_region_shift;             // a region is 2^shift consecutive bins
    int stage_regions;                  // <= RFK_STAGE_MAX_REGIONS
};

#define RFK_FIXED_SCALE 16777216.0f  // 2^24
#define RFK_STAGE_CHUNK 512u         // records per chunk (4 KB); at least RFK_BLOCK, so one iteration of a CTA never spans three chunks
#define RFK_STAGE_MAX_REGIONS 64u
#define RFK_STAGE_DEAD 0xffffffffu    // no chunk left in the region's queue: the samples of this chunk number are reduced directly

__device__ __forceinline__ unsigned int rfk_hash32(unsigned int h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// A bijection of [0, RFK_BLOCK): j = (tid * a + b) mod RFK_BLOCK with a odd, a and b fresh every iteration.
// Two particles of one warp land in the same warp again with probability ~1/8 (8 warps), as under a uniform
// permutation; the multiplier changes every iteration, so no pair stays together.
// Only the low log2(RFK_BLOCK) bits of the multiplier matter: the low byte of the LCG key walks all 256 residues
// (full period), `| 1` makes it odd; the offset b comes from the well-mixed high half (bits 16 and up).
// Returned as the byte offset of the 16-byte exchange slot, j * 16, from tid * 16: the shift rides on the multiply.
__device__ __forceinline__ unsigned int rfk_deal_offset(unsigned int tid16, unsigned int key) {
    return (tid16 * (key | 1u) + (key >> 12)) & ((RFK_BLOCK - 1) << 4);
}

// src/hammersley.cpp:29-48 for point `i` (z = w = 0)
__device__ __forceinline__ float2 rfk_sample_point(unsigned int i, int bits, float inv_max) {
    unsigned int flipped = bits ? (__brev(i) >> (32 - bits)) : 0u;
    float fx = (float)i * inv_max;
    float fy = (float)flipped * inv_max;
    return make_float2((float)((double)fx * 2.0 - 1.0), (float)((double)fy * 2.0 - 1.0));
}

// flame.glsl:78-84: screen affine, floor, bounds and opacity test, row flip. Returns the bin index or -1.
// The truncation coords = ivec2(floor(pos)) is an addition of 2^23 rounded toward zero: for 0 <= p < 2^23 the sum's
// mantissa field is floor(p) exactly, so its bit pattern is B + floor(p) with B = 0x4B000000 — an FADD.RZ on the FMA
// pipe instead of an F2I on the quarter-rate conversion unit. The same integer also carries the bounds test:
// t = bits - B, as unsigned, is below W exactly when 0 <= floor(p) < W (W < 2^23, checked by the host): a negative p
// gives a sum below 2^23 (bits < B, t wraps to a huge value; -0.0 gives t = 0, and floor(-0.0) is in bounds), p >= 2^23
// moves the exponent (t >= 2^23), a sum that is negative, infinite or NaN has bits >= 0x7F800000 or the sign bit set
// (t >= 0x34800000). Non-finite positions therefore never bin, where the reference leaves ivec2(floor(NaN)) undefined.
__device__ __forceinline__ unsigned int rfk_trunc_biased(float p) { return __float_as_uint(__fadd_rz(p, 8388608.0f)) - 0x4B000000u; }
__device__ __forceinline__ bool rfk_bin_test(float x, float y, float w, const float* ss, int W, int H, unsigned int& cx, unsigned int& cy) {
    const vec2 pos = rfk_affine(ss[0], ss[1], ss[2], ss[3], ss[4], ss[5], x, y);  // nested fma, as flame.glsl:79-80
    cx = rfk_trunc_biased(pos.x);
    cy = rfk_trunc_biased(pos.y);
    return cx < (unsigned int)W && cy < (unsigned int)H && w > 0.0f;
}
__device__ __forceinline__ int rfk_bin_of(unsigned int cx, unsigned int cy, int W, int H) { return (int)(((unsigned int)(H - 1) - cy) * (unsigned int)W + cx); }
__device__ __forceinline__ int rfk_bin_index(float x, float y, float w, const float* ss, int W, int H) {
    unsigned int cx, cy;
    if (!rfk_bin_test(x, y, w, ss, W, H, cx, cy)) return -1;
    return rfk_bin_of(cx, cy, W, H);
}

#if RFK_L2_HINTS
// The same test and index as rfk_bin_index, also returning the bit of the bin's 16 x 16 tile in the hot map.
__device__ __forceinline__ int rfk_bin_index_hot(float x, float y, float w, const float* ss, int W, int H,
                                                 const unsigned int* hot_map, int tiles_x, bool& hot) {
    unsigned int cx, cy;
    hot = true;
    if (!rfk_bin_test(x, y, w, ss, W, H, cx, cy)) return -1;
    const int row = H - 1 - (int)cy, col = (int)cx;
    if (hot_map) {
        const unsigned int tile = (unsigned int)(row >> 4) * (unsigned int)tiles_x + (unsigned int)(col >> 4);
        hot = (__ldg(hot_map + (tile >> 5)) >> (tile & 31u)) & 1u;
    }
    return row * W + col;
}

// Reduction into a bin of a tile that is NOT worth keeping in L2: the sector is marked evict-first, so the flood of
// one-off misses of a histogram many times larger than L2 stops evicting the densely hit tiles ("hot map").
__device__ __forceinline__ void rfk_red_add_v4_evict_first(float4* addr, float r, float g, float b, float a, unsigned long long policy) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(r), "f"(g), "f"(b), "f"(a), "l"(policy) : "memory");
}
#endif

__device__ __forceinline__ unsigned int rfk_palette_index(float z) {
    // min(255, uint(ceil(z * 255))): one saturating convert (negative and NaN -> 0; uint(negative) is undefined in GLSL)
    unsigned int u = __float2uint_ru(z * 255.0f);
    return u < 255u ? u : 255u;
}

__device__ __forceinline__ void rfk_red_add_v4(float4* addr, float r, float g, float b, float a) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(r), "f"(g), "f"(b), "f"(a) : "memory");
}

#if RFK_WARP_AGGREGATE
// Sums `v` over the lanes of `peers` (lanes that hit the same bin); the lowest lane
// of each group ends up with the group total. All lanes of `mask` must call.
__device__ __forceinline__ float4 rfk_reduce_peers(unsigned int mask, unsigned int peers, float4 v, int lane) {
    int rel_pos = __popc(peers & ((1u << lane) - 1u));
    peers &= (0xfffffffeu << lane);
    while (__any_sync(mask, peers)) {
        int next = __ffs(peers);
        int src = next ? next - 1 : lane;
        float tx = __shfl_sync(mask, v.x, src), ty = __shfl_sync(mask, v.y, src);
        float tz = __shfl_sync(mask, v.z, src), tw = __shfl_sync(mask, v.w, src);
        if (next) { v.x += tx; v.y += ty; v.z += tz; v.w += tw; }
        unsigned int done = rel_pos & 1;
        peers &= ~__ballot_sync(mask, done);
        rel_pos >>= 1;
    }
    return v;
}
#endif

template <bool DRAW>
__device__ __forceinline__ void rfk_iterate_body(const rfk_iter_params& p) {
    __shared__ float4 pal[256];
    // re-deal exchange, double buffered: (x, y, colour, -) in one 128-bit slot — one STS.128 and one LDS.128 per iteration;
    // conflict-free for any odd multiplier (the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups)
    __shared__ float4 ex[2][RFK_BLOCK];
#if RFK_COUNT_XFORMS
    __shared__ unsigned int xcount[RFK_NUM_XFORMS + 1];
#endif
#if RFK_STAGED_BINS
    // per region: samples of this CTA so far, and the places in the region's queue of the CTA's four most recent chunks
    __shared__ unsigned int st_fill[DRAW ? RFK_STAGE_MAX_REGIONS : 1], st_chunk[DRAW ? RFK_STAGE_MAX_REGIONS : 1][4];
#endif

    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31u;
    const unsigned int blocks_per_ts = (unsigned int)p.ppt / RFK_BLOCK;
    const unsigned int ts = blockIdx.x / blocks_per_ts;                  // gl_WorkGroupID.y
    const size_t slot = (size_t)blockIdx.x * RFK_BLOCK + tid;            // ts * ppt + gl_GlobalInvocationID.x

    rfk_stage_params(p.fp_inflated + (size_t)ts * RFK_TOTAL_PARAMS);
    if (DRAW) for (int i = tid; i < 256; i += RFK_BLOCK) pal[i] = p.palette[i];
#if RFK_COUNT_XFORMS
    if (tid <= RFK_NUM_XFORMS) xcount[tid] = 0;
#endif
#if RFK_STAGED_BINS
    if (DRAW) for (int i = tid; i < p.stage_regions; i += RFK_BLOCK) st_fill[i] = 0;
#endif

    rfk_rng rs = p.rng[slot];
    float x, y, c;
    __syncthreads();

    unsigned int binned = 0;
#if RFK_L2_HINTS
    unsigned long long evict_first_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(evict_first_policy));
#endif

    // flame.glsl:51-53: the first thread of a workgroup burns one randf() per pass to pick the group's xform. Here
    // every lane burns one every 32 iterations and runs it through get_xform_id() — the generated if-chain of
    // xform_select.tpl.glsl, same binary32 additions in the same order — and iteration i uses lane (i mod 32)'s pick:
    // the same number of extra draws per warp, all lanes active when they are made, one shuffle per iteration.
    int pick_pool = 0;
    int pick_count = 0;
    auto pick_xform = [&]() -> int {
#if RFK_PER_LANE_XFORM
        return get_xform_id(rfk_randf(rs));
#else
        if ((pick_count & 31) == 0) pick_pool = get_xform_id(rfk_randf(rs));
        const int xid = __shfl_sync(0xffffffffu, pick_pool, pick_count & 31);
        pick_count++;
        return xid;
#endif
    };
    const unsigned int tid16 = tid << 4;
#if RFK_DEAL_PERIOD != 1
    int since_deal = 0;
#endif
    unsigned int deal_key = rfk_hash32(p.deal_seed ^ (blockIdx.x * 0x9E3779B9u));
    // the two exchange buffers by shared-window address; `ex_cur = ex_both - ex_cur` flips between them (one uniform op)
    unsigned int ex_cur = (unsigned int)__cvta_generic_to_shared(&ex[0][0]);
    const unsigned int ex_both = ex_cur + (unsigned int)__cvta_generic_to_shared(&ex[1][0]);
    auto deal_store = [&]() {
        deal_key = deal_key * 1664525u + 1013904223u;  // CTA-uniform
        // one 128-bit store. (Tried: 64 + 32 bits, which saves the three moves that line x, y and the colour up in a quad — two
        // shared-memory stores per iteration cost more in the MIO queue than the moves cost in issue slots: 1.29 -> 1.38 ms.)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %3};" ::"r"(ex_cur + rfk_deal_offset(tid16, deal_key)), "f"(x), "f"(y), "f"(c) : "memory");
    };
    auto deal_load = [&]() {  // after the barrier that follows deal_store
        float unused;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(c), "=f"(unused) : "r"(ex_cur + tid16) : "memory");
        ex_cur = ex_both - ex_cur;
    };
    auto deal = [&]() { deal_store(); __syncthreads(); deal_load(); };

    if (!DRAW && p.first_run) {
        // flame.glsl:58-65: every temporal sample starts from the same Hammersley set, jittered
        const unsigned int gid = (blockIdx.x % blocks_per_ts) * RFK_BLOCK + tid;
        int xid = pick_xform();
        float2 s = rfk_sample_point(gid, p.hammersley_bits, p.hammersley_inv_max);
        float r0 = rfk_randf(rs);
        float r1 = rfk_randf(rs);
        vec2 sc = sincos(sqrtf(r1));
        float m = r0 * .1f * PI * 2.0f;
        vec4 r = dispatch<true>(vec3(s.x + m * sc.x, s.y + m * sc.y, 0.0f), xid, rs);
        x = r.x; y = r.y; c = r.z;
        deal();
    } else {
        float4 st = p.particles[slot];
        x = st.x; y = st.y; c = st.z;
    }

    // Iterations run in blocks that end where the pool of picks does: the refill test and the pool index leave the inner loop.
    for (int it = 0; it < p.num_iter;) {
#if RFK_PER_LANE_XFORM
        const int block_end = it + 1;
#else
        if ((pick_count & 31) == 0) pick_pool = get_xform_id(rfk_randf(rs));
        int pick_lane = pick_count & 31;
        const int block_len = ::min(32 - pick_lane, p.num_iter - it);
        const int block_end = it + block_len;
        pick_count += block_len;
        // the inner loop counts on the pool index alone
        const int pick_lane_end = pick_lane + block_len;
#endif
#if RFK_PER_LANE_XFORM
      for (; it < block_end; ++it) {
        const int xid = get_xform_id(rfk_randf(rs));
#else
      // (the pick and the coefficients of the next iteration fetched one iteration ahead, as rfk_iterate_pairs does: -1 % on
      // the shipped genome, +0.6 % / +4 % on the stress genome's specialised / generic build at 32 registers — not here)
      for (; pick_lane != pick_lane_end; ++pick_lane) {
        const int xid = __shfl_sync(0xffffffffu, pick_pool, pick_lane);
#endif
#if RFK_COUNT_XFORMS
        if (DRAW) {  // picks of drawn iterations only (the read-out of main.cpp:595-611)
  #if RFK_PER_LANE_XFORM
            atomicAdd(&xcount[xid], 1u);
  #else
            if (lane == 0) atomicAdd(&xcount[xid], 32u);
  #endif
        }
#endif
        __builtin_assume(xid >= 0 && xid < (RFK_NUM_XFORMS > 0 ? RFK_NUM_XFORMS : 1));  // get_xform_id() returns nothing else: no clamp before the jump table
        vec4 r = dispatch<false>(vec3(x, y, c), xid, rs);
        x = r.x; y = r.y; c = r.z;  // flame.glsl:72

        if (DRAW) {
            float fx = r.x, fy = r.y, fc = r.z, fw = r.w;
#if RFK_HAS_FINAL
            {   // src/flame.cpp:23: result = dispatch(result.xyz, -1) * vec4(1, 1, 1, result.w)
                vec4 q = dispatch<false>(vec3(r.x, r.y, r.z), -1, rs);
                fx = q.x; fy = q.y; fc = q.z; fw = q.w * r.w;
            }
#endif
#if RFK_L2_HINTS
            bool hot;
            const int idx = rfk_bin_index_hot(fx, fy, fw, p.ss_affine, p.bin_w, p.bin_h, p.hot_map, p.hot_tiles_x, hot);
            const bool in_bounds = idx >= 0;
#else
            unsigned int cx, cy;
            const bool in_bounds = rfk_bin_test(fx, fy, fw, p.ss_affine, p.bin_w, p.bin_h, cx, cy);
            const int idx = rfk_bin_of(cx, cy, p.bin_w, p.bin_h);  // meaningful only where in_bounds
#endif
#if RFK_STAGED_BINS
            // A histogram many times larger than L2 turns every reduction into a DRAM sector read-modify-write at a random
            // address (76 B of traffic per sample on the 2.12 GB histogram of config 3). Here the sample is appended instead,
            // as an 8-byte record, to the queue of its region (2^shift consecutive bins, sized to sit in L2), and
            // stage_accumulate_kernel then walks the queues region by region, so the reductions of one region meet in L2
            // (static_kernels.cu). A queue is a list of 4 KB chunks. A CTA numbers its samples per region with a
            // shared-memory counter: sample n goes to slot n mod 512 of the CTA's (n / 512)-th chunk of that region, and the
            // thread that draws slot 0 takes the chunk from the region's global cursor — one global atomic per 512 samples
            // (one per sample ran at the same-address atomic rate of L2, 22 G/s). The records are written after the
            // barrier of the re-deal, when the chunk numbers of this iteration are all published. A region whose queue is
            // exhausted falls back to the direct reduction.
            // (Tried and dropped: chunks of 32 records owned by a warp, numbered without atomics or barrier — the sectors of
            // 75 000 open chunks fill too slowly, L2 writes them back half empty: 3.4 GB written and 1.5 GB read per call
            // instead of 2.0 and 0.3.)
            const unsigned int st_region = (unsigned int)idx >> p.stage_region_shift;  // junk where out of bounds, and unused there
            // one shared-memory atomic per in-bounds lane. (Grouping the lanes of a region first, so that one lane adds for all:
            // with MATCH.ANY the MIO queue was the limit, issue slots 43 % busy; with six votes on the 6-bit region number
            // 83 % busy but 47 more instructions per iteration; the plain ATOMS — 2 cycles per lane on the LSU, which has the
            // room — is 199 instead of 246 instructions per iteration and 1.73 instead of 2.19 ms per call.)
            unsigned int st_pos = 0;
            if (in_bounds) st_pos = atomicAdd(&st_fill[st_region], 1u);
            unsigned int* const st_base = &st_chunk[st_region][(st_pos / RFK_STAGE_CHUNK) & 3u];
            unsigned int st_rec = 0;
            if (in_bounds) {
                st_rec = (((unsigned int)idx & ((1u << p.stage_region_shift) - 1u)) << 8) | rfk_palette_index(fc);
                if ((st_pos & (RFK_STAGE_CHUNK - 1u)) == 0u) {
                    const unsigned int k = atomicAdd(p.stage_cursors + st_region, 1u);
                    const bool got = k < p.stage_capacity;
                    // published as the index of the chunk's first record (below 2^32: the host caps the queues at 32 GiB);
                    // four slots: a reader of chunk n and the opener of chunk n + 4 are never in the same or adjacent iterations
                    *st_base = got ? (st_region * p.stage_capacity + k) * RFK_STAGE_CHUNK : RFK_STAGE_DEAD;
                    if (got) p.stage_fill[(size_t)st_region * p.stage_capacity + k] = RFK_STAGE_CHUNK;  // full, unless it is the CTA's last one
                }
                binned++;
            }
  #if !RFK_PER_LANE_XFORM && RFK_DEAL_PERIOD == 1
            deal_store();  // the re-deal's barrier doubles as the one before the records are written
  #endif
            __syncthreads();
            if (in_bounds) {
                const unsigned int first = *st_base;
                if (first != RFK_STAGE_DEAD) {
                    __stcs(p.stage_records + (first | (st_pos & (RFK_STAGE_CHUNK - 1u))), make_uint2(st_rec, __float_as_uint(fw)));
                } else {
                    const float4 col = pal[st_rec & 255u];
                    rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, fw);
                }
            }
#elif RFK_WARP_AGGREGATE && !RFK_DETERMINISTIC
            const unsigned int hit = __ballot_sync(0xffffffffu, in_bounds);
            if (in_bounds) {
                float4 col = pal[rfk_palette_index(fc)];
                float4 v = make_float4(col.x, col.y, col.z, fw);
                unsigned int peers = __match_any_sync(hit, idx);
                v = rfk_reduce_peers(hit, peers, v, lane);
                if ((int)lane == __ffs(peers) - 1) rfk_red_add_v4(p.bins + idx, v.x, v.y, v.z, v.w);
                binned++;
            }
#else
            if (in_bounds) {
  #if !RFK_DETERMINISTIC && !RFK_L2_HINTS
                // The palette row is read as 64 + 32 bits into the first three registers of the reduction's operand quad;
                // the density lane is already there. (One LDS.128 would overwrite that lane and cost three MOVs to regroup;
                // `volatile` keeps ptxas from fusing the two loads back into one.)
                const unsigned int prow = (unsigned int)__cvta_generic_to_shared(&pal[rfk_palette_index(fc)]);
                float4 col;
                asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(col.x), "=f"(col.y) : "r"(prow));
                asm volatile("ld.volatile.shared.f32 %0, [%1+8];" : "=f"(col.z) : "r"(prow));
  #else
                float4 col = pal[rfk_palette_index(fc)];
  #endif
  #if RFK_DETERMINISTIC
                unsigned long long* b = p.fixed_bins + (size_t)idx * 4;
                atomicAdd(b + 0, (unsigned long long)__float2ll_rn(col.x * RFK_FIXED_SCALE));
                atomicAdd(b + 1, (unsigned long long)__float2ll_rn(col.y * RFK_FIXED_SCALE));
                atomicAdd(b + 2, (unsigned long long)__float2ll_rn(col.z * RFK_FIXED_SCALE));
                atomicAdd(b + 3, (unsigned long long)__float2ll_rn(fw * RFK_FIXED_SCALE));
  #elif RFK_L2_HINTS
                col.w = fw;  // the palette row lands in the vector operand of the reduction; only the density lane is patched
                if (hot) rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, col.w);
                else rfk_red_add_v4_evict_first(p.bins + idx, col.x, col.y, col.z, col.w, evict_first_policy);
  #else
                col.w = fw;  // the palette row lands in the vector operand of the reduction; only the density lane is patched
                rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, col.w);
  #endif
                binned++;
            }
#endif
        }
#if !RFK_PER_LANE_XFORM
  #if RFK_DEAL_PERIOD == 1
    #if RFK_STAGED_BINS
        if (DRAW) deal_load(); else deal();  // stored before the staging barrier above
    #else
        deal();
    #endif
  #else
        if (++since_deal == RFK_DEAL_PERIOD) { since_deal = 0; deal(); }
  #endif
#endif
      }
#if !RFK_PER_LANE_XFORM
      it = block_end;
#endif
    }

    p.particles[slot] = make_float4(x, y, c, 0.0f);
    p.rng[slot] = rs;  // flame.glsl:89

#if RFK_STAGED_BINS
    if (DRAW) {  // the last chunk of every region is handed over with what it holds
        __syncthreads();
        for (unsigned int r = tid; r < (unsigned int)p.stage_regions; r += RFK_BLOCK) {
            const unsigned int total = st_fill[r];
            if (!total) continue;
            const unsigned int last = (total - 1u) / RFK_STAGE_CHUNK, first = st_chunk[r][last & 3u];
            if (first != RFK_STAGE_DEAD) p.stage_fill[first / RFK_STAGE_CHUNK] = total - last * RFK_STAGE_CHUNK;
        }
    }
#endif
    if (DRAW) {
        // flame.glsl:85 does one same-address atomic per sample; one per warp here
        for (int o = 16; o > 0; o >>= 1) binned += __shfl_xor_sync(0xffffffffu, binned, o);
        if (lane == 0 && binned) atomicAdd(p.counters, (unsigned long long)binned);
    }
#if RFK_COUNT_XFORMS
    __syncthreads();
    if (tid < RFK_NUM_XFORMS && xcount[tid]) atomicAdd(p.counters + 1 + tid, (unsigned long long)xcount[tid]);
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// Two particles per thread (kernel option pair_particles, the default where no other option asks for the single-particle
// body): a CTA of RFK_BLOCK / 2 threads owns the same pool of RFK_BLOCK particles; thread t holds the particles of slots
// t and t + RFK_BLOCK / 2 and their two JSF32 states. One pick, one walk to the xform's case, one fetch of its rotated
// coefficients, one re-deal key, one barrier and one trip of the loop serve two iterations of the chaos game (~20 of the
// 135 instructions of an iteration are such per-trip overhead), and the two independent dependency chains overlap each
// other's MUFU / shared-memory latency. 40 registers per thread at 1536 resident threads: 3072 particles per SM (2048 with
// one particle per thread). Shipped genome: 135.0 -> 121.6 warp instructions per iteration.
// The warp's pick now covers 64 particles per iteration instead of 32 (the reference: 256).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RFK_PAIRS
#define RFK_PAIRS 0  // set by the host for the value-specialised build (flame::variant_source); the generic module keeps one particle per thread
#endif
#ifndef RFK_PAIRS_MIN_BLOCKS
#define RFK_PAIRS_MIN_BLOCKS (3072 / RFK_BLOCK)  // 40 registers at 1536 resident threads: the best of 64 / 48 / 40 / 36 / 32 registers measured
#endif
#define RFK_PAIRS_AVAILABLE (RFK_PAIRS && !RFK_PER_LANE_XFORM && !RFK_WARP_AGGREGATE && !RFK_DETERMINISTIC && !RFK_COUNT_XFORMS && !RFK_L2_HINTS && !RFK_STAGED_BINS && RFK_DEAL_PERIOD == 1)
#if RFK_PAIRS_AVAILABLE
template <bool first_run>
__device__ __forceinline__ void dispatch2(vec3 v0, vec3 v1, int xform, rfk_rng& rs0, rfk_rng& rs1, vec4& o0, vec4& o1) {
    dispatch2_a<first_run>(v0, v1, xform, rs0, rs1, rfk_aff[xform + 1], o0, o1);
}

template <bool DRAW>
__device__ __forceinline__ void rfk_iterate_pairs(const rfk_iter_params& p) {
    constexpr unsigned int HALF = RFK_BLOCK / 2;  // threads per CTA
    __shared__ float4 pal[256];
    __shared__ float4 ex[2][RFK_BLOCK];

    const unsigned int tid = threadIdx.x;
    const unsigned int lane = tid & 31u;
    const unsigned int blocks_per_ts = (unsigned int)p.ppt / RFK_BLOCK;
    const unsigned int ts = blockIdx.x / blocks_per_ts;
    const size_t slot0 = (size_t)blockIdx.x * RFK_BLOCK + tid, slot1 = slot0 + HALF;

    rfk_stage_params(p.fp_inflated + (size_t)ts * RFK_TOTAL_PARAMS);
    if (DRAW) for (int i = tid; i < 256; i += HALF) pal[i] = p.palette[i];

    rfk_rng rs0 = p.rng[slot0], rs1 = p.rng[slot1];
    float x0, y0, c0, x1, y1, c1;
    __syncthreads();

    unsigned int binned = 0;
    int pick_pool = 0, pick_count = 0;
    const unsigned int tid16 = tid << 4;
    unsigned int deal_key = rfk_hash32(p.deal_seed ^ (blockIdx.x * 0x9E3779B9u));
    unsigned int ex_cur = (unsigned int)__cvta_generic_to_shared(&ex[0][0]);
    const unsigned int ex_both = ex_cur + (unsigned int)__cvta_generic_to_shared(&ex[1][0]);
    // the bijection j = (slot * a + b) mod RFK_BLOCK of rfk_deal_offset for both slots of the thread: slot1 = slot0 + HALF and
    // a is odd, so j1 = j0 + HALF mod RFK_BLOCK — the second offset is the first with its top bit flipped
    auto deal = [&]() {
        deal_key = deal_key * 1664525u + 1013904223u;  // CTA-uniform
        const unsigned int off0 = rfk_deal_offset(tid16, deal_key), at0 = ex_cur + off0, at1 = ex_cur + (off0 ^ (HALF << 4));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %3};" ::"r"(at0), "f"(x0), "f"(y0), "f"(c0) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %3};" ::"r"(at1), "f"(x1), "f"(y1), "f"(c1) : "memory");
        __syncthreads();
        float unused;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(y0), "=f"(c0), "=f"(unused) : "r"(ex_cur + tid16) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x1), "=f"(y1), "=f"(c1), "=f"(unused) : "r"(ex_cur + tid16 + (HALF << 4)) : "memory");
        ex_cur = ex_both - ex_cur;
    };

    if (!DRAW && p.first_run) {
        // flame.glsl:58-65 for both slots; the draws of a slot come from that slot's generator, in the single-particle order
        const unsigned int gid0 = (blockIdx.x % blocks_per_ts) * RFK_BLOCK + tid, gid1 = gid0 + HALF;
        pick_pool = get_xform_id(rfk_randf(rs0));
        const int xid = __shfl_sync(0xffffffffu, pick_pool, 0);
        pick_count = 1;
        const float2 s0 = rfk_sample_point(gid0, p.hammersley_bits, p.hammersley_inv_max), s1 = rfk_sample_point(gid1, p.hammersley_bits, p.hammersley_inv_max);
        const float r00 = rfk_randf(rs0), r01 = rfk_randf(rs0), r10 = rfk_randf(rs1), r11 = rfk_randf(rs1);
        const vec2 sc0 = sincos(sqrtf(r01)), sc1 = sincos(sqrtf(r11));
        const float m0 = r00 * .1f * PI * 2.0f, m1 = r10 * .1f * PI * 2.0f;
        vec4 q0, q1;
        dispatch2<true>(vec3(s0.x + m0 * sc0.x, s0.y + m0 * sc0.y, 0.0f), vec3(s1.x + m1 * sc1.x, s1.y + m1 * sc1.y, 0.0f), xid, rs0, rs1, q0, q1);
        x0 = q0.x; y0 = q0.y; c0 = q0.z; x1 = q1.x; y1 = q1.y; c1 = q1.z;
        deal();
    } else {
        const float4 a = p.particles[slot0], b = p.particles[slot1];
        x0 = a.x; y0 = a.y; c0 = a.z; x1 = b.x; y1 = b.y; c1 = b.z;
    }

    auto bin = [&](float fx, float fy, float fc, float fw) {  // flame.glsl:78-86
        unsigned int cx, cy;
        if (rfk_bin_test(fx, fy, fw, p.ss_affine, p.bin_w, p.bin_h, cx, cy)) {
            const int idx = rfk_bin_of(cx, cy, p.bin_w, p.bin_h);
#if RFK_EXPERIMENT == 1  // timing only: rows forced into distinct bank groups per quarter warp = the cost of eight swizzled copies
            const unsigned int prow = (unsigned int)__cvta_generic_to_shared(&pal[(rfk_palette_index(fc) & ~7u) | (lane & 7u)]);
            float4 col;
            asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(col.x), "=f"(col.y), "=f"(col.z), "=f"(col.w) : "r"(prow));
#elif RFK_EXPERIMENT == 2  // timing only: one 128-bit load of the random row
            const unsigned int prow = (unsigned int)__cvta_generic_to_shared(&pal[rfk_palette_index(fc)]);
            float4 col;
            asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(col.x), "=f"(col.y), "=f"(col.z), "=f"(col.w) : "r"(prow));
#elif RFK_EXPERIMENT == 3  // timing only: no palette
            float4 col = make_float4(fc, fc, fc, fw);
#else
            const unsigned int prow = (unsigned int)__cvta_generic_to_shared(&pal[rfk_palette_index(fc)]);
            float4 col;
            asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(col.x), "=f"(col.y) : "r"(prow));
            asm volatile("ld.volatile.shared.f32 %0, [%1+8];" : "=f"(col.z) : "r"(prow));
#endif
            col.w = fw;
#if RFK_EXPERIMENT == 4  // timing only: no reduction (the palette row is still read: the loads are volatile)
            if (col.x == -1.0f) rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, col.w);
#else
            rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, col.w);
#endif
            binned++;
        }
    };

    for (int it = 0; it < p.num_iter;) {
        if ((pick_count & 31) == 0) pick_pool = get_xform_id(rfk_randf(rs0));
        int pick_lane = pick_count & 31;
        const int block_len = ::min(32 - pick_lane, p.num_iter - it);
        pick_count += block_len;
        const int pick_lane_end = pick_lane + block_len;
#if RFK_EXPERIMENT == 7  // A/B: the pick and the coefficients fetched at the top of the iteration that uses them
        for (; pick_lane != pick_lane_end; ++pick_lane) {
            const int xid = __shfl_sync(0xffffffffu, pick_pool, pick_lane);
            __builtin_assume(xid >= 0 && xid < (RFK_NUM_XFORMS > 0 ? RFK_NUM_XFORMS : 1));
            vec4 r0, r1;
            dispatch2<false>(vec3(x0, y0, c0), vec3(x1, y1, c1), xid, rs0, rs1, r0, r1);
#else
        // the pick of an iteration and the rotated coefficients of its xform are fetched one iteration ahead, right after the
        // dispatch: the shuffle and the shared-memory load then complete while the iteration bins and re-deals, instead of
        // standing between the barrier and the first arithmetic of the next one (1.208 -> 1.195 ms per call on the shipped
        // genome, 2.28 -> 2.20 on the stress genome; RFK_EXPERIMENT=7 is the old order)
        int xid = __shfl_sync(0xffffffffu, pick_pool, pick_lane);
        __builtin_assume(xid >= 0 && xid < (RFK_NUM_XFORMS > 0 ? RFK_NUM_XFORMS : 1));
        float4 coeff = rfk_aff[xid + 1];
        while (pick_lane != pick_lane_end) {
            vec4 r0, r1;
            dispatch2_a<false>(vec3(x0, y0, c0), vec3(x1, y1, c1), xid, rs0, rs1, coeff, r0, r1);
            ++pick_lane;
            xid = __shfl_sync(0xffffffffu, pick_pool, pick_lane & 31);  // past the block's end: lane 0's pick, not used
            __builtin_assume(xid >= 0 && xid < (RFK_NUM_XFORMS > 0 ? RFK_NUM_XFORMS : 1));
            coeff = rfk_aff[xid + 1];
#endif
            x0 = r0.x; y0 = r0.y; c0 = r0.z; x1 = r1.x; y1 = r1.y; c1 = r1.z;  // flame.glsl:72
            if (DRAW) {
#if RFK_HAS_FINAL
                vec4 q0, q1;  // src/flame.cpp:23: result = dispatch(result.xyz, -1) * vec4(1, 1, 1, result.w)
                dispatch2<false>(vec3(r0.x, r0.y, r0.z), vec3(r1.x, r1.y, r1.z), -1, rs0, rs1, q0, q1);
                bin(q0.x, q0.y, q0.z, q0.w * r0.w);
                bin(q1.x, q1.y, q1.z, q1.w * r1.w);
#else
                bin(r0.x, r0.y, r0.z, r0.w);
                bin(r1.x, r1.y, r1.z, r1.w);
#endif
            }
            deal();
        }
        it += block_len;
    }

    p.particles[slot0] = make_float4(x0, y0, c0, 0.0f);
    p.particles[slot1] = make_float4(x1, y1, c1, 0.0f);
    p.rng[slot0] = rs0;  // flame.glsl:89
    p.rng[slot1] = rs1;
    if (DRAW) {
        for (int o = 16; o > 0; o >>= 1) binned += __shfl_xor_sync(0xffffffffu, binned, o);
        if (lane == 0 && binned) atomicAdd(p.counters, (unsigned long long)binned);
    }
}
#endif  // RFK_PAIRS_AVAILABLE

#ifndef RFK_DRAW_ONLY
#define RFK_DRAW_ONLY 0  // 1: the module holds rfk_draw alone (the staged kernels of the automatic mode, built on first use)
#endif
#if RFK_PAIRS_AVAILABLE
// same names, half the threads per CTA (the host reads RFK_PAIRS back from the options it compiled with)
extern "C" __global__ void __launch_bounds__(RFK_BLOCK / 2, RFK_PAIRS_MIN_BLOCKS) rfk_draw(const __grid_constant__ rfk_iter_params p) { rfk_iterate_pairs<true>(p); }
#else
extern "C" __global__ void RFK_LAUNCH_BOUNDS rfk_draw(const __grid_constant__ rfk_iter_params p) { rfk_iterate_body<true>(p); }
#endif
#ifndef RFK_HOT_ONLY
#define RFK_HOT_ONLY 0   // 1: rfk_warm, rfk_draw and rfk_single_step only (the value-specialised build of kernel option `specialize`)
#endif
#if !RFK_DRAW_ONLY
#if RFK_PAIRS_AVAILABLE
extern "C" __global__ void __launch_bounds__(RFK_BLOCK / 2, RFK_PAIRS_MIN_BLOCKS) rfk_warm(const __grid_constant__ rfk_iter_params p) { rfk_iterate_pairs<false>(p); }
#else
extern "C" __global__ void RFK_LAUNCH_BOUNDS rfk_warm(const __grid_constant__ rfk_iter_params p) { rfk_iterate_body<false>(p); }
#endif

#if !RFK_HOT_ONLY
// The reference's own dispatch structure, restated for the GPU: shaders/flame.glsl:41-90 as ONE iteration per launch
// on a (PPT / 256, TS) grid, particle and RNG state through global memory, one xform per 256-thread workgroup picked
// by its first thread, shuffle-buffer gather / scatter (flame.glsl:31-37, :55-61). Not the product path: it is the
// same-hardware baseline the register-resident kernels are measured against, and — fed the oracle's shuffle tables and
// pass ids — it reproduces the oracle's RNG states bit for bit and its particle buffers to rounding.
struct rfk_pass_params {
    const float4* pos_in;               // binding 0
    float4* pos_out;                    // binding 1
    uint4* rng;                         // binding 11
    const unsigned int* shuf_buf;       // binding 4: [num_shuffle][PPT]
    const float* fp_inflated;           // binding 9
    const float4* palette;              // binding 6
    float4* bins;                       // binding 8
    unsigned long long* counters;       // binding 10
    float ss_affine[6];
    int bin_w, bin_h;
    float bin_wf, bin_hf;
    int ppt;
    unsigned int shuf_buf_idx_in, shuf_buf_idx_out;
    int random_read, random_write, first_run, do_draw;
};

extern "C" __global__ void __launch_bounds__(256) rfk_reference_pass(const __grid_constant__ rfk_pass_params p) {
    __shared__ int xid;
    __shared__ float4 pal[256];
    const unsigned int gid = blockIdx.x * 256 + threadIdx.x;        // gl_GlobalInvocationID.x
    const size_t base = (size_t)blockIdx.y * p.ppt;                 // gl_WorkGroupID.y * gl_WorkGroupSize.x * gl_NumWorkGroups.x
    rfk_rng rs = p.rng[base + gid];                                 // load_random_state()
    rfk_stage_params(p.fp_inflated + (size_t)blockIdx.y * RFK_TOTAL_PARAMS);
    pal[threadIdx.x] = p.palette[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) xid = get_xform_id(rfk_randf(rs));      // flame.glsl:51-53
    const unsigned int i_idx = p.random_read ? p.shuf_buf[gid + (size_t)p.ppt * p.shuf_buf_idx_in] : gid;
    const unsigned int o_idx = p.random_write ? p.shuf_buf[gid + (size_t)p.ppt * p.shuf_buf_idx_out] : gid;
    const float4 part = p.pos_in[(p.first_run ? 0 : base) + i_idx];
    float x = part.x, y = part.y;
    if (p.first_run) {
        float r0 = rfk_randf(rs);
        float r1 = rfk_randf(rs);
        vec2 sc = sincos(sqrtf(r1));
        float m = r0 * .1f * PI * 2.0f;
        x += m * sc.x; y += m * sc.y;
    }
    __syncthreads();
    vec4 r = p.first_run ? dispatch<true>(vec3(x, y, part.z), xid, rs) : dispatch<false>(vec3(x, y, part.z), xid, rs);
    p.pos_out[base + o_idx] = make_float4(r.x, r.y, r.z, 0.0f);
    if (p.do_draw) {
        float fx = r.x, fy = r.y, fc = r.z, fw = r.w;
#if RFK_HAS_FINAL
        {
            vec4 q = dispatch<false>(vec3(r.x, r.y, r.z), -1, rs);
            fx = q.x; fy = q.y; fc = q.z; fw = q.w * r.w;
        }
#endif
        const int idx = rfk_bin_index(fx, fy, fw, p.ss_affine, p.bin_w, p.bin_h);
        if (idx >= 0) {
            float4 col = pal[rfk_palette_index(fc)];
            rfk_red_add_v4(p.bins + idx, col.x, col.y, col.z, fw);  // the reference's racy +=, made atomic
            atomicAdd(p.counters, 1ull);                             // flame.glsl:85
        }
    }
    p.rng[base + gid] = rs;                                         // save_random_state()
}

#endif  // !RFK_HOT_ONLY

// Test hook: one dispatch(v, xid) per thread on caller-supplied particles, RNG states
// and parameter block — the single-step level of the parity contract.
extern "C" __global__ void rfk_single_step(int n, const float* __restrict__ xyz, const int* __restrict__ xid, uint4* rng,
                                           const float* __restrict__ fp, int first_run, float4* out) {
    rfk_stage_params(fp);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rfk_rng rs = rng[i];
    vec3 v(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    vec4 r = first_run ? dispatch<true>(v, xid[i], rs) : dispatch<false>(v, xid[i], rs);
    out[i] = make_float4(r.x, r.y, r.z, r.w);
    rng[i] = rs;
}

#if !RFK_HOT_ONLY
// Test hook: xform selection for caller-supplied ratios (xform_select.tpl.glsl).
extern "C" __global__ void rfk_select_xform(int n, const float* __restrict__ ratio, const float* __restrict__ fp, int* out) {
    rfk_stage_params(fp);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = get_xform_id(ratio[i]);  // the generated if-chain, as the hot kernels call it
}

struct rfk_bucket_params { float ss_affine[6]; int bin_w, bin_h; };

// Test hook: the bucket index of flame.glsl:78-84 for caller-supplied (x, y, w) and colour.
extern "C" __global__ void rfk_bucket_index(int n, const float* __restrict__ xyzw, const __grid_constant__ rfk_bucket_params bp,
                                            int* idx_out, int* pal_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx_out[i] = rfk_bin_index(xyzw[4 * i], xyzw[4 * i + 1], xyzw[4 * i + 3], bp.ss_affine, bp.bin_w, bp.bin_h);
    pal_out[i] = (int)rfk_palette_index(xyzw[4 * i + 2]);
}
#endif  // !RFK_HOT_ONLY
#endif  // !RFK_DRAW_ONLY
