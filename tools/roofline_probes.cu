// Micro-benchmarks for the rooflines MEASURED_PEAKS.json does not hold: FP32 FMA issue rate,
// MUFU (SFU) rate, plain instruction issue rate, and red.global.add.v4.f32 throughput at random 16-byte-aligned
// addresses over the three histogram footprints (14.7 MB, 132.7 MB, 2.12 GB). Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/roofline_probes tools/roofline_probes.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void ffma_kernel(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void mufu_kernel(float* out, float a) {
    float x0 = threadIdx.x * 1e-3f + a, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        x0 = __sinf(x0); x1 = __sinf(x1); x2 = __sinf(x2); x3 = __sinf(x3);  // FMUL + MUFU.SIN each
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

// mixed FMA-pipe / ALU-pipe stream: what the issue port sustains when both pipes are fed
__global__ void mixed_kernel(float* out, float a, float b, unsigned int k) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    unsigned int u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        x0 = fmaf(x0, a, b); u0 = (u0 ^ k) + (u0 >> 3);
        x1 = fmaf(x1, a, b); u1 = (u1 ^ k) + (u1 >> 3);
        x2 = fmaf(x2, a, b); u2 = (u2 ^ k) + (u2 >> 3);
        x3 = fmaf(x3, a, b); u3 = (u3 ^ k) + (u3 >> 3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + (float)(u0 + u1 + u2 + u3);
}

__device__ __forceinline__ unsigned int hash32(unsigned int h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// one red.global.add.v4.f32 per thread per iteration at a pseudo-random float4 slot of [0, nbins)
__global__ void red_v4_kernel(float4* bins, unsigned int nbins, int iters, unsigned int seed) {
    unsigned int s = hash32(seed ^ (blockIdx.x * blockDim.x + threadIdx.x));
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        unsigned int idx = (unsigned int)(((unsigned long long)hash32(s) * nbins) >> 32);
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(bins + idx), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(1.0f) : "memory");
    }
}

// four scalar atomicAdd(float) for comparison
__global__ void red_scalar_kernel(float4* bins, unsigned int nbins, int iters, unsigned int seed) {
    unsigned int s = hash32(seed ^ (blockIdx.x * blockDim.x + threadIdx.x));
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        unsigned int idx = (unsigned int)(((unsigned long long)hash32(s) * nbins) >> 32);
        float* p = reinterpret_cast<float*>(bins + idx);
        atomicAdd(p, 1.0f); atomicAdd(p + 1, 0.5f); atomicAdd(p + 2, 0.25f); atomicAdd(p + 3, 1.0f);
    }
}

template <typename F>
float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    launch();
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) launch();
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, threads = 1024, blocks = sms * 2;
    float* out;
    CHECK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)));
    const double warps = (double)blocks * threads / 32;

    float ms = time_ms([&] { ffma_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 5);
    const double ffma_winst = warps * ITERS * 8 / (ms * 1e-3);
    ms = time_ms([&] { mufu_kernel<<<blocks, threads>>>(out, 0.1f); }, 5);
    const double mufu_winst = warps * ITERS * 4 / (ms * 1e-3);
    ms = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f, 0x9E3779B9u); }, 5);
    const double mixed_winst = warps * ITERS * 4 * 3 / (ms * 1e-3);  // FFMA + LOP3 + LEA.HI per stream step (checked in the SASS)

    printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.l2CacheSize, prop.clockRate);
    printf(" \"ffma_warp_inst_per_s\": %.4g, \"ffma_tflops\": %.4g,\n", ffma_winst, ffma_winst * 64 / 1e12);
    printf(" \"mufu_warp_inst_per_s\": %.4g,\n \"mixed_fma_alu_warp_inst_per_s\": %.4g,\n", mufu_winst, mixed_winst);
    printf(" \"issue_peak_warp_inst_per_s_at_max_clock\": %.4g,\n", (double)sms * 4 * prop.clockRate * 1e3);

    const size_t footprints[3] = {1280ull * 720, 3840ull * 2160, 15360ull * 8640};
    const char* names[3] = {"14.7MB_720p", "132.7MB_4K", "2.12GB_15360x8640"};
    printf(" \"red_v4_f32\": {");
    for (int f = 0; f < 3; f++) {
        float4* bins;
        CHECK(cudaMalloc(&bins, footprints[f] * sizeof(float4)));
        CHECK(cudaMemset(bins, 0, footprints[f] * sizeof(float4)));
        const int iters = 64, rblocks = sms * 8, rthreads = 256;
        const double n = (double)rblocks * rthreads * iters;
        float v4 = time_ms([&] { red_v4_kernel<<<rblocks, rthreads>>>(bins, (unsigned int)footprints[f], iters, 12345u); }, 10);
        float sc = time_ms([&] { red_scalar_kernel<<<rblocks, rthreads>>>(bins, (unsigned int)footprints[f], iters, 12345u); }, 10);
        printf("%s\"%s\": {\"v4_gred_per_s\": %.4g, \"v4_algorithmic_gb_per_s\": %.4g, \"scalar_x4_gupdates_per_s\": %.4g}", f ? ", " : "", names[f],
               n / (v4 * 1e-3) / 1e9, n * 16 / (v4 * 1e-3) / 1e9, n / (sc * 1e-3) / 1e9);
        CHECK(cudaFree(bins));
    }
    printf("}}\n");
    return 0;
}
