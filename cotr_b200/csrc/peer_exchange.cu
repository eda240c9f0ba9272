// Result exchange between the ranks of one node over NVLink peer memory.
//
// The path shards by pairs: every rank runs whole forwards and the only data that ever crosses GPUs is the (b, Q, 2)
// fp32 block of predictions each rank produces (8 KB at the headline shape).  An NCCL all-gather makes every rank's
// kernel wait for every other rank inside the step; here nobody waits to SEND:
//   * every rank owns one symmetric buffer  data[slot][rank][block_bytes] + flag[slot][rank]  (cudaMalloc, exported
//     with cudaIpcGetMemHandle, opened by the other ranks with cudaIpcOpenMemHandle -> NVLink / NVSwitch stores);
//   * push (exchange_push_kernel): step number s, slot s % slots.  One column of CTAs per destination copies the block
//     straight into that peer's data[slot][my rank] (16-byte stores over NVLink), then the last CTA of the column
//     publishes flag[slot][my rank] = s in the peer's memory (fence.sys before the flag: the block is visible before it);
//   * wait (exchange_wait_kernel): polls the LOCAL flags of the slot until all ranks show step s (ld.acquire.sys, bounded),
//     copies the slot into the caller's tensor with L2 loads (remote stores land in L2, never in this SM's L1) and
//     re-checks the flags: a flag that moved on means a writer lapped the reader (status 2) - only possible when a
//     caller lets more than `slots - 1` pushes go by before waiting for an older one.
// A rank that alternates push / wait (the engines, the end-to-end call) can never be lapped with >= 2 slots: a peer can
// only push step s + 1 after its own wait(s) returned, and it can only push step s + 2 after it saw MY push of s + 1,
// which my stream issues after my wait(s) finished copying.
#include <cstring>

#include "../../include/cotr_b200.h"
#include "common.cuh"

namespace cotr {

namespace {

constexpr int kMaxRanks = 16;
constexpr int kPushThreads = 256;
constexpr int kPushChunkBytes = 32 * 1024;       // one CTA of a destination column moves up to this much per trip
constexpr long long kSpinBudget = 6000000000ll;  // clock64 ticks (~3 s): a peer that never shows up is an error, not a hang

struct PeerTable {
    unsigned char* base[kMaxRanks];
};
struct SizeTable {
    unsigned long long bytes[kMaxRanks];          // block size of every rank in this step
    unsigned long long dst_off[kMaxRanks];        // where it goes in the gathered tensor
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid (world, chunks): column x pushes this rank's block into peer x
__global__ void __launch_bounds__(kPushThreads) exchange_push_kernel(const uint4* __restrict__ src, unsigned long long bytes, PeerTable peers,
                                                                    unsigned long long data_off, unsigned long long flag_off,
                                                                    unsigned long long seq, unsigned int* __restrict__ column_done) {
    const int peer = blockIdx.x;
    uint4* dst = reinterpret_cast<uint4*>(peers.base[peer] + data_off);
    const unsigned long long n16 = bytes >> 4;
    for (unsigned long long i = (unsigned long long)blockIdx.y * kPushThreads + threadIdx.x; i < n16; i += (unsigned long long)gridDim.y * kPushThreads)
        dst[i] = __ldcg(src + i);                  // the block was written by the previous kernel of this stream: L2 is current
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                    // this CTA's stores before its arrival
        const unsigned int arrived = atomicAdd(&column_done[peer], 1u);
        if (arrived == gridDim.y - 1) {            // last CTA of the column: every chunk is on its way
            column_done[peer] = 0;
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[peer] + flag_off), seq);
        }
    }
}

__global__ void __launch_bounds__(kPushThreads) exchange_wait_kernel(const unsigned char* __restrict__ slot_data, const unsigned long long* __restrict__ flags,
                                                                    int world, unsigned long long seq, unsigned long long block_stride,
                                                                    SizeTable sizes, unsigned char* __restrict__ dst, int* __restrict__ status) {
    __shared__ int failed;
    if (threadIdx.x == 0) failed = 0;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) < seq) {
            if (clock64() - t0 > kSpinBudget) { failed = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (failed) {
        if (threadIdx.x == 0 && blockIdx.x == 0) *status = 1;        // a rank never published this step
        return;
    }
    if (dst != nullptr) {
        for (int r = 0; r < world; ++r) {
            const uint4* s = reinterpret_cast<const uint4*>(slot_data + (unsigned long long)r * block_stride);
            uint4* d = reinterpret_cast<uint4*>(dst + sizes.dst_off[r]);
            const unsigned long long n16 = sizes.bytes[r] >> 4;
            for (unsigned long long i = (unsigned long long)blockIdx.x * kPushThreads + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * kPushThreads)
                d[i] = __ldcg(s + i);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < world && ld_acquire_sys(flags + threadIdx.x) != seq) *status = 2;      // lapped by a writer
}

}  // namespace

}  // namespace cotr

using namespace cotr;

struct cotr_exchange {
    int device = 0, rank = 0, world = 1, slots = 0;
    size_t block_bytes = 0, slot_bytes = 0, flags_off = 0, total_bytes = 0;
    unsigned char* local = nullptr;
    unsigned char* peer[kMaxRanks] = {};
    bool ipc_opened[kMaxRanks] = {};
    bool connected = false;
    unsigned int* column_done = nullptr;     // device: arrival counters of the push columns
    int* status_host = nullptr;              // pinned + mapped: written by the wait kernel
    int* status_dev = nullptr;
    long long seq = 0;
};

extern "C" {

int cotr_exchange_create(int device, int rank, int world, size_t block_bytes, int slots, cotr_exchange** out) {
    COTR_CHECK(out != nullptr, "cotr_exchange_create: null output pointer");
    *out = nullptr;
    COTR_CHECK(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "cotr_exchange_create: rank %d of %d (at most %d ranks)", rank, world, kMaxRanks);
    COTR_CHECK(slots >= 2 && slots <= 64, "cotr_exchange_create: %d slots (2..64)", slots);
    COTR_CHECK(block_bytes > 0 && block_bytes % 16 == 0, "cotr_exchange_create: block size %zu is not a positive multiple of 16 bytes", block_bytes);
    COTR_CHECK_CUDA(cudaSetDevice(device));
    cotr_exchange* ex = new cotr_exchange();
    ex->device = device; ex->rank = rank; ex->world = world; ex->slots = slots;
    ex->block_bytes = block_bytes;
    ex->slot_bytes = block_bytes * (size_t)world;
    ex->flags_off = ex->slot_bytes * (size_t)slots;
    ex->total_bytes = ex->flags_off + sizeof(unsigned long long) * (size_t)slots * (size_t)world;
    cudaError_t e = cudaMalloc(&ex->local, ex->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->total_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&ex->column_done, sizeof(unsigned int) * kMaxRanks);
    if (e == cudaSuccess) e = cudaMemset(ex->column_done, 0, sizeof(unsigned int) * kMaxRanks);
    if (e == cudaSuccess) e = cudaHostAlloc(&ex->status_host, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *ex->status_host = 0; e = cudaHostGetDevicePointer(&ex->status_dev, ex->status_host, 0); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("cotr_exchange_create: %s (%zu bytes)", cudaGetErrorString(e), ex->total_bytes);
        cotr_exchange_destroy(ex);
        return 1;
    }
    ex->peer[rank] = ex->local;
    ex->connected = world == 1;
    *out = ex;
    return 0;
}

int cotr_exchange_handle(cotr_exchange* ex, void* handle_out) {
    COTR_CHECK(ex && handle_out, "cotr_exchange_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == COTR_EXCHANGE_HANDLE_BYTES, "handle size");
    COTR_CHECK_CUDA(cudaSetDevice(ex->device));
    cudaIpcMemHandle_t h;
    COTR_CHECK_CUDA(cudaIpcGetMemHandle(&h, ex->local));
    std::memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int cotr_exchange_connect(cotr_exchange* ex, const void* handles) {
    COTR_CHECK(ex && handles, "cotr_exchange_connect: null argument");
    COTR_CHECK(!ex->connected, "cotr_exchange_connect: already connected");
    COTR_CHECK_CUDA(cudaSetDevice(ex->device));
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cotr_exchange_connect: cannot map the buffer of rank %d: %s", r, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return 1;
        }
        ex->peer[r] = static_cast<unsigned char*>(p);
        ex->ipc_opened[r] = true;
    }
    ex->connected = true;
    return 0;
}

int cotr_exchange_connect_local(cotr_exchange* ex, cotr_exchange* const* all) {
    COTR_CHECK(ex && all, "cotr_exchange_connect_local: null argument");
    COTR_CHECK(!ex->connected, "cotr_exchange_connect_local: already connected");
    COTR_CHECK_CUDA(cudaSetDevice(ex->device));
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        const cotr_exchange* o = all[r];
        COTR_CHECK(o && o->rank == r && o->world == ex->world && o->slots == ex->slots && o->block_bytes == ex->block_bytes,
                   "cotr_exchange_connect_local: entry %d does not describe rank %d of the same exchange", r, r);
        if (o->device != ex->device) {
            int can = 0;
            COTR_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, ex->device, o->device));
            COTR_CHECK(can, "cotr_exchange_connect_local: device %d cannot address device %d", ex->device, o->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) COTR_CHECK_CUDA(e);
            (void)cudaGetLastError();
        }
        ex->peer[r] = o->local;
    }
    ex->connected = true;
    return 0;
}

long long cotr_exchange_push(cotr_exchange* ex, const void* block_dev, size_t bytes, void* cuda_stream) {
    if (!ex || !ex->connected) { set_error("cotr_exchange_push: exchange is not connected"); return -1; }
    if (bytes > ex->block_bytes || bytes % 16 != 0 || (bytes > 0 && !block_dev) || (reinterpret_cast<uintptr_t>(block_dev) & 15)) {
        set_error("cotr_exchange_push: block of %zu bytes (capacity %zu; size and address must be multiples of 16)", bytes, ex->block_bytes);
        return -1;
    }
    if (cudaSetDevice(ex->device) != cudaSuccess) { set_error("cotr_exchange_push: cudaSetDevice failed"); return -1; }
    const long long seq = ++ex->seq;
    const int slot = (int)(seq % ex->slots);
    PeerTable t{};
    for (int r = 0; r < ex->world; ++r) t.base[r] = ex->peer[r];
    const unsigned long long data_off = (unsigned long long)slot * ex->slot_bytes + (unsigned long long)ex->rank * ex->block_bytes;
    const unsigned long long flag_off = ex->flags_off + sizeof(unsigned long long) * ((unsigned long long)slot * ex->world + ex->rank);
    int chunks = (int)((bytes + kPushChunkBytes - 1) / kPushChunkBytes);
    chunks = chunks < 1 ? 1 : (chunks > 16 ? 16 : chunks);
    exchange_push_kernel<<<dim3(ex->world, chunks), kPushThreads, 0, (cudaStream_t)cuda_stream>>>(
        static_cast<const uint4*>(block_dev), bytes, t, data_off, flag_off, (unsigned long long)seq, ex->column_done);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("cotr_exchange_push: launch failed: %s", cudaGetErrorString(e)); return -1; }
    return seq;
}

int cotr_exchange_wait(cotr_exchange* ex, long long seq, void* dst_dev, const size_t* bytes_per_rank, void* cuda_stream) {
    COTR_CHECK(ex && ex->connected, "cotr_exchange_wait: exchange is not connected");
    COTR_CHECK(seq >= 1 && seq <= ex->seq, "cotr_exchange_wait: step %lld was never pushed (last push: %lld)", seq, ex->seq);
    COTR_CHECK(seq + ex->slots > ex->seq, "cotr_exchange_wait: step %lld has been overwritten (last push %lld, %d slots)", seq, ex->seq, ex->slots);
    COTR_CHECK((reinterpret_cast<uintptr_t>(dst_dev) & 15) == 0, "cotr_exchange_wait: destination must be 16-byte aligned");
    COTR_CHECK_CUDA(cudaSetDevice(ex->device));
    SizeTable sz{};
    unsigned long long off = 0, most = 0;
    for (int r = 0; r < ex->world; ++r) {
        const size_t b = bytes_per_rank ? bytes_per_rank[r] : ex->block_bytes;
        COTR_CHECK(b <= ex->block_bytes && b % 16 == 0, "cotr_exchange_wait: rank %d block of %zu bytes (capacity %zu, multiple of 16)", r, b, ex->block_bytes);
        sz.bytes[r] = b; sz.dst_off[r] = off; off += b;
        most = b > most ? b : most;
    }
    const int slot = (int)(seq % ex->slots);
    int ctas = dst_dev ? (int)((most + kPushChunkBytes - 1) / kPushChunkBytes) : 1;
    ctas = ctas < 1 ? 1 : (ctas > 32 ? 32 : ctas);
    exchange_wait_kernel<<<ctas, kPushThreads, 0, (cudaStream_t)cuda_stream>>>(
        ex->local + (size_t)slot * ex->slot_bytes, reinterpret_cast<const unsigned long long*>(ex->local + ex->flags_off) + (size_t)slot * ex->world,
        ex->world, (unsigned long long)seq, ex->block_bytes, sz, static_cast<unsigned char*>(dst_dev), ex->status_dev);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int cotr_exchange_status(const cotr_exchange* ex) {
    if (!ex) return -1;
    return *static_cast<volatile int*>(ex->status_host);
}

void cotr_exchange_destroy(cotr_exchange* ex) {
    if (!ex) return;
    (void)cudaSetDevice(ex->device);
    (void)cudaDeviceSynchronize();
    for (int r = 0; r < ex->world; ++r)
        if (ex->ipc_opened[r] && ex->peer[r]) (void)cudaIpcCloseMemHandle(ex->peer[r]);
    if (ex->local) (void)cudaFree(ex->local);
    if (ex->column_done) (void)cudaFree(ex->column_done);
    if (ex->status_host) (void)cudaFreeHost(ex->status_host);
    (void)cudaGetLastError();
    delete ex;
}

}  // extern "C"
