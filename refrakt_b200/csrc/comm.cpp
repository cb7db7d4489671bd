// One process per GPU: NCCL (loaded at run time) for bootstrap, barriers and the fallback data path; CUDA IPC peer
// mappings over NVLink for the data path of a sharded frame. See comm.hpp.
#include "comm.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace rfk::comm {

namespace {

struct nccl_api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

// A process that already carries an NCCL (PyTorch bundles one) gets that copy: dlopen by soname returns the loaded object.
const nccl_api& nccl() {
    static nccl_api api = [] {
        nccl_api a;
        const char* override_path = std::getenv("RFK_NCCL_LIBRARY");
        for (const char* name : {override_path, "libnccl.so.2", "libnccl.so"}) {
            if (!name) continue;
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) {
            const char* why = dlerror();  // one call: it clears the message
            throw std::runtime_error("multi-GPU: libnccl.so.2 cannot be loaded (set RFK_NCCL_LIBRARY): " + std::string(why ? why : ""));
        }
        auto get = [&](const char* sym, auto& fn) {
            void* p = dlsym(a.handle, sym);
            if (!p) throw std::runtime_error(std::string("multi-GPU: NCCL symbol missing: ") + sym);
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
        };
        get("ncclGetUniqueId", a.GetUniqueId);
        get("ncclCommInitRank", a.CommInitRank);
        get("ncclCommDestroy", a.CommDestroy);
        get("ncclGetErrorString", a.GetErrorString);
        get("ncclAllReduce", a.AllReduce);
        get("ncclReduce", a.Reduce);
        get("ncclAllGather", a.AllGather);
        get("ncclSend", a.Send);
        get("ncclRecv", a.Recv);
        get("ncclGroupStart", a.GroupStart);
        get("ncclGroupEnd", a.GroupEnd);
        return a;
    }();
    return api;
}

void nccl_ok(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(r));
}
void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

struct ipc_entry { cudaIpcMemHandle_t handle; unsigned long long offset; unsigned long long valid; };
constexpr int kSlots = 3;  // bins, rgba8, image

struct state {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool p2p = false, p2p_tried = false;
    unsigned char* scratch = nullptr;        // device staging for the small collectives
    peer_buffers peers;
    void* opened[kMaxWorld][kSlots] = {};    // base pointers returned by cudaIpcOpenMemHandle
    void* published[kSlots] = {};            // what this rank published last
    std::uint64_t generation = ~0ull;        // the caller's allocation generation of that exchange
};
state g;

constexpr std::size_t kScratchBytes = 64 * 1024;

}  // namespace

row_slab slab_of(int height, int halo, int rank, int world) {
    if (height < 0 || halo < 0 || world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("row slab: bad arguments");
    const int base = height / world, rem = height % world;
    row_slab s;
    s.y0 = rank * base + (rank < rem ? rank : rem);
    s.y1 = s.y0 + base + (rank < rem ? 1 : 0);
    s.src_y0 = s.y0 - halo < 0 ? 0 : s.y0 - halo;
    s.src_y1 = s.y1 + halo > height ? height : s.y1 + halo;
    if (s.y1 == s.y0) s.src_y0 = s.src_y1 = s.y0;  // more ranks than rows: an empty slab reads nothing
    return s;
}

std::uint64_t pass_share(std::uint64_t total, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("pass share: bad arguments");
    return total / world + ((std::uint64_t)rank < total % world ? 1 : 0);
}

void unique_id(unsigned char out[kIdBytes]) {
    static_assert(sizeof(ncclUniqueId) == kIdBytes, "NCCL unique id size");
    ncclUniqueId id;
    nccl_ok(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, &id, kIdBytes);
}

void init(const unsigned char id_bytes[kIdBytes], int rank, int world) {
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) throw std::invalid_argument("rfk_comm_init: rank / world out of range (at most 16 ranks)");
    if (g.comm) destroy();
    cuda_ok(cudaFree(nullptr), "CUDA initialisation (is a GPU visible?)");
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, kIdBytes);
    nccl_ok(nccl().CommInitRank(&g.comm, world, id, rank), "ncclCommInitRank");
    g.rank = rank;
    g.world = world;
    g.p2p = false;
    g.p2p_tried = false;
    cuda_ok(cudaMalloc(&g.scratch, kScratchBytes), "cudaMalloc(comm scratch)");
}

void release_peers() {
    for (int r = 0; r < kMaxWorld; r++)
        for (int k = 0; k < kSlots; k++)
            if (g.opened[r][k]) { cudaIpcCloseMemHandle(g.opened[r][k]); g.opened[r][k] = nullptr; }
    g.peers = peer_buffers{};
    for (auto& p : g.published) p = nullptr;
}

void destroy() {
    release_peers();
    if (g.comm) { nccl().CommDestroy(g.comm); g.comm = nullptr; }
    cudaFree(g.scratch); g.scratch = nullptr;
    g.rank = 0; g.world = 1; g.p2p = false; g.p2p_tried = false;
}

bool active() { return g.comm != nullptr; }
int rank() { return g.rank; }
int world() { return g.world; }
bool p2p() { return g.p2p; }

static void need_comm() {
    if (!g.comm) throw std::runtime_error("multi-GPU: rfk_comm_init has not been called");
}

void barrier_sum(std::uint64_t* values_dev, int n, cudaStream_t s) {
    need_comm();
    if (n < 1 || n > 8) throw std::invalid_argument("barrier_sum: 1 to 8 counters");
    nccl_ok(nccl().AllReduce(values_dev, values_dev, (size_t)n, ncclUint64, ncclSum, g.comm, s), "ncclAllReduce(barrier)");
}

void reduce_histogram(float4* bins, std::size_t count, int root, cudaStream_t s) {
    need_comm();
    if (g.world == 1) return;
    if (root >= g.world) throw std::invalid_argument("reduce_histogram: no such root");
    if (root < 0) nccl_ok(nccl().AllReduce(bins, bins, count * 4, ncclFloat32, ncclSum, g.comm, s), "ncclAllReduce(histogram)");
    else nccl_ok(nccl().Reduce(bins, bins, count * 4, ncclFloat32, ncclSum, root, g.comm, s), "ncclReduce(histogram)");
}

// Publishes (handle, offset inside the allocation) of this rank's three buffers and maps the peers'. cudaIpcGetMemHandle
// names the whole cudaMalloc allocation a pointer lies in, so the offset travels with it. Every rank learns through the
// all-gather whether ALL mappings worked; one failure switches every rank to the NCCL data path.
const peer_buffers& exchange(float4* my_bins, uchar4* my_rgba8, float4* my_image, std::uint64_t generation, cudaStream_t s) {
    need_comm();
    void* mine[kSlots] = {my_bins, my_rgba8, my_image};
    if (g.p2p_tried && generation == g.generation) return g.peers;  // nothing re-allocated since the last exchange, on any rank
    g.generation = generation;
    release_peers();
    g.p2p_tried = true;
    const char* env = std::getenv("RFK_COMM_P2P");
    bool ok = !(env && std::atoi(env) == 0) && g.world > 1;

    std::vector<ipc_entry> table((size_t)g.world * kSlots);
    ipc_entry local[kSlots];
    std::memset(local, 0, sizeof local);
    for (int k = 0; k < kSlots && ok; k++) {
        if (!mine[k]) continue;
        void* base = nullptr;
        size_t size = 0;
        // the allocation's base: the IPC handle maps the allocation, not the pointer
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, mine[k]) != cudaSuccess || attr.type != cudaMemoryTypeDevice) { ok = false; cudaGetLastError(); break; }
        typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
        static range_fn get_range = [] {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
            return reinterpret_cast<range_fn>(p);
        }();
        CUdeviceptr b = 0;
        if (!get_range || get_range(&b, &size, (CUdeviceptr)mine[k]) != CUDA_SUCCESS) { ok = false; break; }
        base = (void*)b;
        if (cudaIpcGetMemHandle(&local[k].handle, base) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        local[k].offset = (unsigned long long)((char*)mine[k] - (char*)base);
        local[k].valid = 1;
    }
    if (!ok) std::memset(local, 0, sizeof local);

    // all-gather of the tables (through the device scratch; NCCL has no host collectives)
    const size_t bytes = sizeof(local);
    if ((size_t)g.world * bytes + bytes > kScratchBytes) throw std::runtime_error("comm scratch too small");
    unsigned char* send = g.scratch + (size_t)g.world * bytes;
    cuda_ok(cudaMemcpyAsync(send, local, bytes, cudaMemcpyHostToDevice, s), "publish ipc handles");
    nccl_ok(nccl().AllGather(send, g.scratch, bytes, ncclUint8, g.comm, s), "ncclAllGather(ipc handles)");
    cuda_ok(cudaMemcpyAsync(table.data(), g.scratch, (size_t)g.world * bytes, cudaMemcpyDeviceToHost, s), "read ipc handles");
    cuda_ok(cudaStreamSynchronize(s), "exchange ipc handles");

    // every rank needs: all bins; rank 0's rgba8 / image when rank 0 has them
    bool all_ok = g.world > 1;
    for (int r = 0; r < g.world; r++) all_ok = all_ok && table[(size_t)r * kSlots].valid;
    if (all_ok) {
        for (int r = 0; r < g.world && all_ok; r++) {
            for (int k = 0; k < kSlots && all_ok; k++) {
                const ipc_entry& e = table[(size_t)r * kSlots + k];
                if (k > 0 && r != 0) continue;  // only rank 0's output images are written remotely
                if (!e.valid) continue;
                void* p = nullptr;
                if (r == g.rank) {
                    p = mine[k];
                } else {
                    void* base = nullptr;
                    if (cudaIpcOpenMemHandle(&base, e.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); break; }
                    g.opened[r][k] = base;
                    p = (char*)base + e.offset;
                }
                if (k == 0) g.peers.bins[r] = (float4*)p;
                else if (k == 1) g.peers.root_rgba8 = (uchar4*)p;
                else g.peers.root_image = (float4*)p;
            }
        }
    }
    // agree on the outcome (one failed mapping anywhere = NCCL data path everywhere)
    std::uint64_t flag = all_ok ? 0 : 1;
    std::uint64_t* flag_dev = reinterpret_cast<std::uint64_t*>(g.scratch);
    cuda_ok(cudaMemcpyAsync(flag_dev, &flag, sizeof flag, cudaMemcpyHostToDevice, s), "publish p2p outcome");
    barrier_sum(flag_dev, 1, s);
    cuda_ok(cudaMemcpyAsync(&flag, flag_dev, sizeof flag, cudaMemcpyDeviceToHost, s), "read p2p outcome");
    cuda_ok(cudaStreamSynchronize(s), "agree on p2p");
    g.p2p = flag == 0 && g.world > 1;
    if (!g.p2p) {
        for (int r = 0; r < kMaxWorld; r++)
            for (int k = 0; k < kSlots; k++)
                if (g.opened[r][k]) { cudaIpcCloseMemHandle(g.opened[r][k]); g.opened[r][k] = nullptr; }
        g.peers = peer_buffers{};
        g.peers.bins[g.rank] = my_bins;
        if (g.rank == 0) { g.peers.root_rgba8 = my_rgba8; g.peers.root_image = my_image; }
    }
    for (int k = 0; k < kSlots; k++) g.published[k] = mine[k];
    return g.peers;
}

void reduce_scatter_slabs_nccl(const float4* bins, float4* slab, int W, int H, const std::vector<row_slab>& slabs, cudaStream_t s) {
    need_comm();
    nccl_ok(nccl().GroupStart(), "ncclGroupStart");
    for (int r = 0; r < g.world; r++) {
        const row_slab& sl = slabs[r];
        const size_t rows = (size_t)(sl.src_y1 - sl.src_y0);
        if (!rows) continue;
        // source rows [src_y0, src_y1) are the histogram rows [H - src_y1, H - src_y0)
        const float4* send = bins + (size_t)(H - sl.src_y1) * W;
        nccl_ok(nccl().Reduce(send, r == g.rank ? (void*)slab : nullptr, rows * W * 4, ncclFloat32, ncclSum, r, g.comm, s), "ncclReduce(row slab)");
    }
    nccl_ok(nccl().GroupEnd(), "ncclGroupEnd");
}

void gather_slabs_nccl(const void* my_rows, void* root_full, std::size_t bytes_per_row, const std::vector<row_slab>& slabs, cudaStream_t s) {
    need_comm();
    nccl_ok(nccl().GroupStart(), "ncclGroupStart");
    if (g.rank == 0) {
        for (int r = 1; r < g.world; r++) {
            const size_t rows = (size_t)(slabs[r].y1 - slabs[r].y0);
            if (rows) nccl_ok(nccl().Recv((char*)root_full + (size_t)slabs[r].y0 * bytes_per_row, rows * bytes_per_row, ncclUint8, r, g.comm, s), "ncclRecv(rows)");
        }
    } else {
        const size_t rows = (size_t)(slabs[g.rank].y1 - slabs[g.rank].y0);
        if (rows) nccl_ok(nccl().Send(my_rows, rows * bytes_per_row, ncclUint8, 0, g.comm, s), "ncclSend(rows)");
    }
    nccl_ok(nccl().GroupEnd(), "ncclGroupEnd");
}

}  // namespace rfk::comm
