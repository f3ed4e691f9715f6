// green.cu — the mutators' stream lives in a GREEN CONTEXT that owns all but a few SMs of the device.
//
// Searches and mutators share one GPU (index_impl.h).  Stream priorities only decide which pending CTA is placed
// first when an SM frees up; a refinement pass keeps every SM full of long-running K4 CTAs (3 per SM, ~1-2 ms each),
// and a batch-1 search — ONE CTA of the CTA-per-query kernel, which needs most of an SM's register file — then waits
// until a whole SM drains: measured p99 23 ms inside a refinement pass (profiles/r2_c5_*).  CUDA 12.4 green contexts
// partition the SMs spatially: the mutator stream is created in a green context over 136 of the 148 SMs, the search
// stream stays in the primary context (all SMs), so 12 SMs never hold mutator work and a small search always finds a
// free one.  Memory, events and modules are shared with the primary context.  The stream is self-tested at creation;
// any failure (older driver, MIG, ...) falls back to the plain low-priority stream.
#include <cuda.h>

#include <cstdlib>

#include "index_impl.h"

namespace vsbi {

namespace {
template <class F>
F entry(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(p);
}
}  // namespace

bool make_green_stream(int device, unsigned reserve_sms, int priority, cudaStream_t* stream_out, void** ctx_out,
                       unsigned* sms_out) {
    *stream_out = nullptr;
    *ctx_out = nullptr;
    using GetDev = CUresult (*)(CUdevice*, int);
    using GetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
    using Split = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    using GenDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    using GCreate = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    using GStream = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
    using GDestroy = CUresult (*)(CUgreenCtx);
    auto get_dev = entry<GetDev>("cuDeviceGet");
    auto get_res = entry<GetRes>("cuDeviceGetDevResource");
    auto split = entry<Split>("cuDevSmResourceSplitByCount");
    auto gen = entry<GenDesc>("cuDevResourceGenerateDesc");
    auto gcreate = entry<GCreate>("cuGreenCtxCreate");
    auto gstream = entry<GStream>("cuGreenCtxStreamCreate");
    auto gdestroy = entry<GDestroy>("cuGreenCtxDestroy");
    if (!get_dev || !get_res || !split || !gen || !gcreate || !gstream || !gdestroy) return false;
    CUdevice dev;
    if (get_dev(&dev, device) != CUDA_SUCCESS) return false;
    CUdevResource all{}, group{}, rest{};
    if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    const unsigned total = all.sm.smCount;
    if (total <= reserve_sms + 8) return false;
    unsigned nb = 1;
    if (split(&group, &nb, &all, &rest, 0, total - reserve_sms) != CUDA_SUCCESS || nb != 1) return false;
    if (group.sm.smCount >= total) return false;  // the split could not leave anything free
    CUdevResourceDesc desc;
    if (gen(&desc, &group, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx g = nullptr;
    if (gcreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream s = nullptr;
    if (gstream(&s, g, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) {
        gdestroy(g);
        return false;
    }
    // self-test: everything the mutators do on their stream must work on this one
    bool ok = true;
    void* d = nullptr;
    uint64_t h[4] = {0, 0, 0, 0};
    cudaEvent_t ev = nullptr;
    ok = ok && cudaMalloc(&d, 64) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        vsb::launch_fill_empty(static_cast<uint64_t*>(d), reinterpret_cast<float*>(static_cast<uint8_t*>(d) + 32), nullptr, 2, 2, s);
        ok = ok && cudaGetLastError() == cudaSuccess;
        ok = ok && cudaEventRecord(ev, s) == cudaSuccess;
        ok = ok && cudaStreamWaitEvent(s, ev, 0) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
        ok = ok && h[0] == 0xFFFFFFFFFFFFFFFFull && h[3] == 0xFFFFFFFFFFFFFFFFull;
    }
    if (ev) cudaEventDestroy(ev);
    if (d) cudaFree(d);
    if (!ok) {
        cudaGetLastError();
        cudaStreamDestroy(s);
        gdestroy(g);
        return false;
    }
    *stream_out = s;
    *ctx_out = g;
    if (sms_out) *sms_out = group.sm.smCount;
    return true;
}

void destroy_green(void* ctx) {
    using GDestroy = CUresult (*)(CUgreenCtx);
    auto gdestroy = entry<GDestroy>("cuGreenCtxDestroy");
    if (ctx && gdestroy) gdestroy(static_cast<CUgreenCtx>(ctx));
}

}  // namespace vsbi
