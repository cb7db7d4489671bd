// Does an L2 eviction-priority hint on red.global.add.v4.f32 help a histogram far larger than L2?
// Address stream modelled on config 3 (2.12 GB float4 histogram): a fraction `hot_share` of the reductions goes to a
// "hot" set of 4 KB blocks scattered over the whole histogram (hot_bytes in total), the rest uniformly anywhere.
// Variants: no hint; hot -> evict_last + cold -> evict_first; hot -> evict_last only; cold -> evict_first only / no_allocate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_policy_probe tools/l2_policy_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned int hash32(unsigned int h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

template <int MODE>
__global__ void probe(float4* bins, unsigned int nbins, unsigned int hot_blocks, unsigned int block_stride, unsigned int hot_threshold, int iters, unsigned int seed) {
    unsigned int s = hash32(seed ^ (blockIdx.x * blockDim.x + threadIdx.x));
    unsigned long long pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        unsigned int h = hash32(s), h2 = hash32(s ^ 0x9E3779B9u);
        bool hot = h2 < hot_threshold;
        unsigned int idx;
        if (hot) {  // a 256-bin (4 KB) block out of hot_blocks, spread with stride block_stride over the histogram
            unsigned int b = (unsigned int)(((unsigned long long)h * hot_blocks) >> 32);
            idx = b * block_stride + (h2 & 255u);
        } else {
            idx = (unsigned int)(((unsigned long long)h * nbins) >> 32);
        }
        float4* a = bins + idx;
        if (MODE == 0 || (MODE == 2 && !hot) || (MODE == 3 && hot)) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(1.0f) : "memory");
        } else if (hot) {
            asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(1.0f), "l"(pol_last) : "memory");
        } else {
            asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(1.0f), "l"(pol_first) : "memory");
        }
    }
}

template <int MODE>
double run(float4* bins, unsigned int nbins, unsigned int hot_blocks, unsigned int stride, double hot_share) {
    const int blocks = 148 * 16, threads = 256, iters = 512;
    unsigned int thr = (unsigned int)(hot_share * 4294967295.0);
    probe<MODE><<<blocks, threads>>>(bins, nbins, hot_blocks, stride, thr, 64, 1);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) probe<MODE><<<blocks, threads>>>(bins, nbins, hot_blocks, stride, thr, iters, 7 + r);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return 3.0 * blocks * threads * iters / (ms * 1e-3) / 1e9;
}

int main() {
    const unsigned int nbins = 15360u * 8640u;
    float4* bins; CHECK(cudaMalloc(&bins, (size_t)nbins * 16)); CHECK(cudaMemset(bins, 0, (size_t)nbins * 16));
    printf("{\"footprint_gb\": %.2f, \"results\": [\n", nbins * 16.0 / 1e9);
    const double shares[] = {0.0, 0.42, 0.52, 0.68};
    const double hot_mb[] = {0.0, 64.0, 106.0, 212.0};
    for (int k = 0; k < 4; k++) {
        unsigned int hot_blocks = hot_mb[k] > 0 ? (unsigned int)(hot_mb[k] * 1e6 / 4096.0) : 1;
        unsigned int stride = (nbins / hot_blocks) & ~255u;
        double a = run<0>(bins, nbins, hot_blocks, stride, shares[k]);
        double b = run<1>(bins, nbins, hot_blocks, stride, shares[k]);
        double c = run<2>(bins, nbins, hot_blocks, stride, shares[k]);
        double d = run<3>(bins, nbins, hot_blocks, stride, shares[k]);
        printf("  {\"hot_share\": %.2f, \"hot_mb\": %.0f, \"gred_no_hint\": %.2f, \"gred_last_and_first\": %.2f, \"gred_hot_last_only\": %.2f, \"gred_cold_first_only\": %.2f}%s\n",
               shares[k], hot_mb[k], a, b, c, d, k < 3 ? "," : "");
    }
    printf("]}\n");
    return 0;
}
