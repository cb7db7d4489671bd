// Does the packed FP32 instruction FFMA2 (fma.rn.f32x2, sm_100) free issue slots? Four kernels, all with 8 independent
// dependency chains per thread: (A) FFMA only, (B) FFMA2 only, (C) FFMA + integer ALU work, (D) FFMA2 + the same ALU work
// with the same FP32 flops as (C). Prints warp-instructions/s and FP32 flop/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu && tools/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int n, float a, float b, unsigned k) {
    float x[CHAINS]; float2 y[CHAINS]; unsigned u[CHAINS];
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = make_float2(x[i], x[i] + 0.5f); u[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);
            if (MODE == 2) { x[i] = fmaf(x[i], a, b); if (i < CHAINS / 2) u[i] = (u[i] ^ k) + (u[i] >> 3); }
            if (MODE == 3) { if (i < CHAINS / 2) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ k) + (u[i] >> 3); } }
        }
    }
    float s = 0; unsigned t = 0;
    for (int i = 0; i < CHAINS; i++) { s += x[i] + y[i].x + y[i].y; t += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <int MODE>
static void run(const char* name, double fp_inst_per_iter, double flop_per_iter, double alu_inst_per_iter) {
    const int blocks = 148 * 8, threads = 256, n = 4096;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
    probe<MODE><<<blocks, threads>>>(out, 64, 1.0001f, 0.5f, 0x9E3779B9u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, n, 1.0001f, 0.5f, 0x9E3779B9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32, s = ms * 1e-3;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"fp_warp_inst_per_s\": %.4g, \"alu_warp_inst_per_s\": %.4g, \"total_warp_inst_per_s\": %.4g, \"fp32_tflops\": %.2f}\n", name, ms,
           warps * n * fp_inst_per_iter / s, warps * n * alu_inst_per_iter / s, warps * n * (fp_inst_per_iter + alu_inst_per_iter) / s, warps * 32 * n * flop_per_iter / s / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("ffma", 8, 16, 0);
    run<1>("ffma2", 8, 32, 0);
    run<2>("ffma+alu", 8, 16, 12);    // 4 chains x (LOP3 + SHF + IADD3) — the compiler may fuse; see SASS
    run<3>("ffma2+alu", 4, 16, 12);
    return 0;
}
