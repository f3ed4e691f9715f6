// xchg.cu — K8 for one-process-per-GPU deployments: the exchange of per-shard top-k over NVLink peer memory.
//
// SURVEY §8e prescribes "ncclAllGather of Q*k*(dist,key), then merge".  The payload is tiny (12*k bytes per query
// per rank, 1.2 MB for 10 000 queries), so the collective is pure latency: an NCCL launch, two stream hand-offs
// and a copy.  Here the collective is fused into the producer/consumer kernels instead:
//   push  : this rank's [q][k] keys + distances are stored straight into EVERY rank's gather buffer (CUDA-IPC
//           mapped peer memory; plain coalesced 16-byte stores over NVLink), then the last CTA to finish raises
//           this rank's step flag in every peer (fence.sys + st.release.sys);
//   merge : one CTA spins (ld.acquire.sys) until all ranks' flags carry the current step, then the ordinary K8
//           merge reads the local gather buffer.
// Two parities of the gather buffer make the scheme safe without any further synchronisation: a rank can only be
// one step ahead of the slowest rank, because its own merge of step s needs every rank's flag for step s.
// A watchdog turns a peer that never arrives into VSB_ENCCL instead of a hung GPU.
#include <cstring>
#include <new>
#include <vector>

#include "index_impl.h"

using vsbi::fail;

namespace {

// one block at the buffer head: result flags[8] (u64) @0, aux flags[8] (u64) @64, result push counter @128, error word
// @132, aux push counter @136
constexpr uint32_t kFlagBytes = 1024;
constexpr uint32_t kAuxFlagOff = 64, kCounterOff = 128, kErrorOff = 132, kAuxCounterOff = 136;

struct Layout {
    uint64_t cap;     // max_q * max_k entries per rank part
    uint32_t world;
    uint64_t aux;     // bytes per rank of the auxiliary all-gather region (query slices), multiple of 16
    size_t part_keys() const { return (size_t)cap * 8; }
    size_t part_dists() const { return (size_t)cap * 4; }
    size_t parity_bytes() const { return (size_t)world * (part_keys() + part_dists()); }
    size_t aux_off(uint32_t parity, uint32_t part) const { return kFlagBytes + 2 * parity_bytes() + ((size_t)parity * world + part) * aux; }
    size_t total() const { return kFlagBytes + 2 * parity_bytes() + 2 * (size_t)world * aux; }
    size_t keys_off(uint32_t parity, uint32_t part) const { return kFlagBytes + parity * parity_bytes() + (size_t)part * part_keys(); }
    size_t dists_off(uint32_t parity, uint32_t part) const {
        return kFlagBytes + parity * parity_bytes() + (size_t)world * part_keys() + (size_t)part * part_dists();
    }
};

struct PeerTable {
    uint8_t* base[8];
};

__global__ void __launch_bounds__(256) xchg_push_kernel(PeerTable peers, uint32_t world, uint32_t rank, size_t keys_off,
                                                        size_t dists_off, const uint64_t* __restrict__ keys,
                                                        const float* __restrict__ dists, size_t n_entries,
                                                        unsigned long long step, uint32_t* counter) {
    // n_entries is even (checked on the host): keys move as 16-byte, distances as 8-byte words
    const size_t n2 = n_entries / 2;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    const uint2* d2 = reinterpret_cast<const uint2*>(dists);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 kv = k4[i];
        const uint2 dv = d2[i];
        for (uint32_t p = 0; p < world; ++p) {
            reinterpret_cast<uint4*>(peers.base[p] + keys_off)[i] = kv;
            reinterpret_cast<uint2*>(peers.base[p] + dists_off)[i] = dv;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            for (uint32_t p = 0; p < world; ++p) {
                unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.base[p]) + rank;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(step) : "memory");
            }
        }
    }
}

// all-gather of an opaque byte block per rank (the e2e path's query slices): same push / flag protocol
__global__ void __launch_bounds__(256) xchg_push_bytes_kernel(PeerTable peers, uint32_t world, uint32_t rank, size_t dst_off,
                                                              const uint4* __restrict__ src, size_t n16,
                                                              unsigned long long step, uint32_t* counter) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (uint32_t p = 0; p < world; ++p) reinterpret_cast<uint4*>(peers.base[p] + dst_off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            for (uint32_t p = 0; p < world; ++p) {
                unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.base[p] + kAuxFlagOff) + rank;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(step) : "memory");
            }
        }
    }
}

// one CTA, one thread per rank: wait until every rank's flag has reached `step`
__global__ void xchg_wait_kernel(const unsigned long long* flags, uint32_t world, unsigned long long step,
                                 long long timeout_cycles, uint32_t* error) {
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        unsigned long long v = 0;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= step) break;
            if (clock64() - t0 > timeout_cycles) {
                atomicExch(error, 1u + threadIdx.x);
                break;
            }
            __nanosleep(100);
        }
    }
}

}  // namespace

struct vsb_xchg {
    int device = 0;
    Layout lay{};
    uint32_t rank = 0;
    uint64_t max_q = 0;
    uint32_t max_k = 0;
    uint8_t* local = nullptr;
    uint32_t* counter = nullptr;  // device words inside the local flag block
    uint32_t* aux_counter = nullptr;
    uint32_t* error = nullptr;
    uint32_t* h_error = nullptr;  // pinned mirror, polled without a sync
    std::vector<uint8_t*> peer;
    std::vector<bool> opened;
    unsigned long long step = 0, aux_step = 0;
    long long timeout_cycles = 4000000000ll;  // ~2 s at 2 GHz
};

extern "C" {

vsb_status vsb_xchg_create(int32_t device, uint32_t world, uint32_t rank, uint64_t max_queries, uint32_t max_k,
                           uint64_t aux_bytes_per_rank, vsb_xchg** out) {
    if (!out) return fail(VSB_EINVAL, "null out");
    *out = nullptr;
    if (world < 1 || world > 8 || rank >= world) return fail(VSB_EINVAL, "world must be 1..8 and rank < world");
    if (max_queries == 0 || max_k == 0 || (uint64_t)world * max_k > 2048) return fail(VSB_EINVAL, "world*max_k must be in [1, 2048]");
    if (device < 0) CU(cudaGetDevice(&device));
    CU(cudaSetDevice(device));
    vsb_xchg* x = new (std::nothrow) vsb_xchg();
    if (!x) return fail(VSB_EOOM, "host allocation failed");
    x->device = device;
    x->rank = rank;
    x->max_q = max_queries;
    x->max_k = max_k;
    x->lay.cap = (max_queries * max_k + 1) & ~1ull;
    x->lay.world = world;
    x->lay.aux = (aux_bytes_per_rank + 15) & ~15ull;
    cudaError_t e = cudaMalloc(&x->local, x->lay.total());  // plain cudaMalloc: CUDA IPC cannot export pool/VMM memory
    if (e != cudaSuccess) {
        delete x;
        return fail(VSB_EOOM, "cudaMalloc(%zu): %s", x->lay.total(), cudaGetErrorString(e));
    }
    cudaMemset(x->local, 0, kFlagBytes);
    x->counter = reinterpret_cast<uint32_t*>(x->local + kCounterOff);
    x->error = reinterpret_cast<uint32_t*>(x->local + kErrorOff);
    x->aux_counter = reinterpret_cast<uint32_t*>(x->local + kAuxCounterOff);
    x->peer.assign(world, nullptr);
    x->opened.assign(world, false);
    x->peer[rank] = x->local;
    *out = x;
    return VSB_OK;
}

void vsb_xchg_destroy(vsb_xchg* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (size_t r = 0; r < x->peer.size(); ++r)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local) cudaFree(x->local);
    delete x;
}

vsb_status vsb_xchg_local_handle(vsb_xchg* x, void* handle_out) {
    if (!x || !handle_out) return fail(VSB_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == VSB_XCHG_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, x->local);
    if (e != cudaSuccess) return fail(VSB_ENCCL, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    std::memcpy(handle_out, &h, sizeof h);
    return VSB_OK;
}

vsb_status vsb_xchg_open(vsb_xchg* x, const void* handles) {
    if (!x || !handles) return fail(VSB_EINVAL, "null argument");
    CU(cudaSetDevice(x->device));
    for (uint32_t r = 0; r < x->lay.world; ++r) {
        if (r == x->rank || x->opened[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * VSB_XCHG_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(VSB_ENCCL, "cudaIpcOpenMemHandle(rank %u): %s", r, cudaGetErrorString(e));
        }
        x->peer[r] = static_cast<uint8_t*>(p);
        x->opened[r] = true;
    }
    return VSB_OK;
}

vsb_status vsb_xchg_allgather_merge(vsb_xchg* x, const uint64_t* d_keys, const float* d_dists, uint64_t q, uint32_t k,
                                    uint64_t* d_out_keys, float* d_out_dists, uint32_t* d_out_counts, void* stream_) {
    if (!x || !d_keys || !d_dists || !d_out_keys || !d_out_dists) return fail(VSB_EINVAL, "null argument");
    if (q == 0 || k == 0) return VSB_OK;
    if (q * k > x->lay.cap) return fail(VSB_EINVAL, "q*k = %llu exceeds the exchange capacity %llu", (unsigned long long)(q * k), (unsigned long long)x->lay.cap);
    if ((q * k) % 2) return fail(VSB_EINVAL, "q*k must be even (16-byte peer stores)");
    for (uint32_t r = 0; r < x->lay.world; ++r)
        if (!x->peer[r]) return fail(VSB_ENCCL, "rank %u's buffer is not mapped: call vsb_xchg_open first", r);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CU(cudaSetDevice(x->device));
    uint32_t err = 0;
    x->step += 1;
    const uint32_t parity = (uint32_t)(x->step & 1);
    PeerTable t{};
    for (uint32_t r = 0; r < x->lay.world; ++r) t.base[r] = x->peer[r];
    const size_t n = (size_t)q * k;
    const unsigned grid = (unsigned)std::min<size_t>((n / 2 + 255) / 256, 148 * 4);
    xchg_push_kernel<<<grid, 256, 0, s>>>(t, x->lay.world, x->rank, x->lay.keys_off(parity, x->rank), x->lay.dists_off(parity, x->rank),
                                          d_keys, d_dists, n, x->step, x->counter);
    xchg_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(x->local), x->lay.world, x->step,
                                      x->timeout_cycles, x->error);
    // the K8 merge over the local gather buffer: parts are cap entries apart
    vsb::launch_merge_topk(reinterpret_cast<const uint64_t*>(x->local + x->lay.keys_off(parity, 0)),
                           reinterpret_cast<const float*>(x->local + x->lay.dists_off(parity, 0)), x->lay.world, q, k, d_out_keys,
                           d_out_dists, d_out_counts, s, x->lay.cap, x->lay.cap);
    vsb::g_kernel_launches += 2;
    CU(cudaGetLastError());
    (void)err;
    return VSB_OK;
}

vsb_status vsb_xchg_allgather_bytes(vsb_xchg* x, const void* d_src, uint64_t bytes_per_rank, void** d_gathered,
                                    void* d_copy_out, void* stream_) {
    if (!x || !d_src || (!d_gathered && !d_copy_out)) return fail(VSB_EINVAL, "null argument");
    if (bytes_per_rank == 0 || bytes_per_rank % 16 || bytes_per_rank > x->lay.aux)
        return fail(VSB_EINVAL, "bytes_per_rank must be a multiple of 16 and <= the aux capacity %llu", (unsigned long long)x->lay.aux);
    for (uint32_t r = 0; r < x->lay.world; ++r)
        if (!x->peer[r]) return fail(VSB_ENCCL, "rank %u's buffer is not mapped: call vsb_xchg_open first", r);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CU(cudaSetDevice(x->device));
    x->aux_step += 1;
    const uint32_t parity = (uint32_t)(x->aux_step & 1);
    PeerTable t{};
    for (uint32_t r = 0; r < x->lay.world; ++r) t.base[r] = x->peer[r];
    const size_t n16 = (size_t)bytes_per_rank / 16;
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 148 * 4);
    // the gathered blocks are lay.aux bytes apart (capacity stride): callers that need them contiguous create the
    // exchange with aux_bytes_per_rank == bytes_per_rank (bench.py does)
    xchg_push_bytes_kernel<<<grid, 256, 0, s>>>(t, x->lay.world, x->rank, x->lay.aux_off(parity, x->rank),
                                                static_cast<const uint4*>(d_src), n16, x->aux_step, x->aux_counter);
    xchg_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(x->local + kAuxFlagOff), x->lay.world,
                                      x->aux_step, x->timeout_cycles, x->error);
    vsb::g_kernel_launches += 2;
    CU(cudaGetLastError());
    if (d_gathered) *d_gathered = x->local + x->lay.aux_off(parity, 0);
    if (d_copy_out)  // out of the window (peers overwrite it two gathers later) into the caller's block, rank by rank
        CU(cudaMemcpy2DAsync(d_copy_out, bytes_per_rank, x->local + x->lay.aux_off(parity, 0), x->lay.aux, bytes_per_rank,
                             x->lay.world, cudaMemcpyDeviceToDevice, s));
    return VSB_OK;
}

/* 0 = healthy; r+1 = the wait kernel gave up on rank r (synchronises the stream it is given) */
vsb_status vsb_xchg_check(vsb_xchg* x, void* stream_) {
    if (!x) return fail(VSB_EINVAL, "null argument");
    CU(cudaSetDevice(x->device));
    uint32_t err = 0;
    CU(cudaMemcpyAsync(&err, x->error, 4, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream_)));
    CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream_)));
    if (err) return fail(VSB_ENCCL, "exchange timed out waiting for rank %u", err - 1);
    return VSB_OK;
}

}  // extern "C"
