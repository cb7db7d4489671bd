// Multi-GPU plumbing of the render path: one process per GPU, NCCL for bootstrap, barriers and the fallback data
// path, and peer memory over NVLink (CUDA IPC) for the histogram reduce-scatter and the image gather of a frame whose
// particle streams are sharded over the GPUs of one box (the reference is single-GPU and has no
// counterpart). NCCL is loaded at run time (dlopen), so the library still loads on a box without it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace rfk::comm {

constexpr int kIdBytes = 128;  // NCCL_UNIQUE_ID_BYTES
constexpr int kMaxWorld = 16;  // slab_reduce sums at most 16 sources

// Output rows [y0, y1) of rank `rank` when `height` rows are cut into `world` slabs (the first height % world slabs take
// one row more), and the source rows [src_y0, src_y1) a density estimation of radius `halo` reads for them.
struct row_slab { int y0 = 0, y1 = 0, src_y0 = 0, src_y1 = 0; };
row_slab slab_of(int height, int halo, int rank, int world);
// passes of this rank when `total` passes are split over `world` ranks (the first total % world ranks take one more)
std::uint64_t pass_share(std::uint64_t total, int rank, int world);

void unique_id(unsigned char out[kIdBytes]);                       // ncclGetUniqueId
void init(const unsigned char id[kIdBytes], int rank, int world);  // ncclCommInitRank on the current device; collective
void destroy();
bool active();
int rank();
int world();
bool p2p();  // the peers' frame buffers are mapped (CUDA IPC over NVLink); false = every exchange goes through NCCL

// Stream-ordered barrier that also sums `values` (n <= 8 counters) over the ranks: one small ncclAllReduce. When it has
// completed on a rank, everything the other ranks enqueued before their call has completed too.
void barrier_sum(std::uint64_t* values_dev, int n, cudaStream_t s);
// in-place sum of the per-rank histograms: onto `root`, or onto every rank when root < 0 (ncclReduce / ncclAllReduce)
void reduce_histogram(float4* bins, std::size_t count, int root, cudaStream_t s);

// Buffers of a sharded frame that the peers read or write. `exchange` (collective) publishes this rank's base pointers
// and maps the peers'; it is called again whenever a rank re-allocated (all ranks re-allocate in lockstep: the sizes
// follow from the frame request) — the caller says so by passing a new `generation`.
struct peer_buffers {
    float4* bins[kMaxWorld] = {};     // every rank's private histogram (read by the slab owners)
    uchar4* root_rgba8 = nullptr;     // rank 0's output images (written by the slab owners)
    float4* root_image = nullptr;
};
const peer_buffers& exchange(float4* my_bins, uchar4* my_rgba8, float4* my_image, std::uint64_t generation, cudaStream_t s);
void release_peers();  // unmaps the peers' buffers (before the owners free them)

// Fallback data path (NCCL): reduce-scatter of row slabs with halo = one ncclReduce per destination rank, grouped;
// gather of the finished rows on rank 0 = grouped send / receive.
// `slabs[r]`: rows of rank r (source rows for the reduce, output rows for the gather); `slab` / `my_rows` hold this rank's.
void reduce_scatter_slabs_nccl(const float4* bins, float4* slab, int W, int H, const std::vector<row_slab>& slabs, cudaStream_t s);
void gather_slabs_nccl(const void* my_rows, void* root_full, std::size_t bytes_per_row, const std::vector<row_slab>& slabs, cudaStream_t s);

}  // namespace rfk::comm
