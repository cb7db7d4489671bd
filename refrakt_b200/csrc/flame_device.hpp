// Device-side helpers shared between flame_device.cpp and the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "flame.hpp"

namespace rfk {

cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
std::uint64_t kernel_launch_count();  // kernels launched by this library since load
void count_launch(unsigned n);

void release_sim_buffers();
std::size_t sim_total_particles();
std::size_t sim_temporal_samples();
const uint4* sim_rng_states();
uint4* sim_rng_states_mutable();
std::size_t sim_shuffle_count();

// NVRTC: CUDA source -> sm_100a cubin (no GPU needed). Throws with the compile log on failure.
std::vector<char> compile_cubin(const std::string& source, const kernel_options& opt, std::string* log_out, int auto_min_blocks = 0);

// test hooks: run the generated device functions on host-supplied vectors
void flame_single_step(flame& f, int n, const float* xyz, const int* xid, std::uint32_t* rng, const float* fp, int first_run, float* out);
void flame_select_xform(flame& f, int n, const float* ratio, const float* fp, int* out);
void flame_bucket_index(flame& f, int n, const float* xyzw, const float ss_affine[6], int W, int H, int* idx_out, int* pal_out);
void flame_animate_host(flame& f, float tss_width, int temporal_samples, float* out);
void flame_kernel_info(flame& f, const char* kernel, int* regs, int* smem_bytes, int* blocks_per_sm);
void flame_read_counters(flame& f, unsigned long long* out, int n);
bool flame_uses_baked(const flame& f);
// kernel option pair_particles: 0 = generic kernels in use / not measured, 1 = two particles per thread, 2 = one; ms_out = the measurement
int flame_pairs_state(const flame& f, float ms_out[2]);  // the last warmup chose the value-specialised kernels (kernel option specialize)
const unsigned long long* flame_binned_counter_dev(flame& f);  // device address of the binned-samples counter (counters[0])
void flame_copy_particles(flame& f, float* out);  // the flame's particle buffer (P x float4), to host

}  // namespace rfk
