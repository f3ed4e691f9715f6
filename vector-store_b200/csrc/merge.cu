// merge.cu — K8: k-way merge of per-part top-k lists by (distance, key).
//
// Used for (a) graph result ∪ brute-force tail inside one index and (b) the per-shard lists that
// an NCCL all-gather lays out as [parts][q][k].  No reference counterpart: the
// reference holds one index per node and never merges.
// One CTA per query; parts*k <= 2048 entries are bitonic-sorted in shared memory.
#include "kernels.h"

namespace vsb {

namespace {

constexpr int K8_THREADS = 128;

__device__ __forceinline__ bool dk_less(uint32_t da, uint64_t ka, uint32_t db, uint64_t kb) {
    return da < db || (da == db && ka < kb);
}

__global__ void __launch_bounds__(K8_THREADS) merge_topk_kernel(const uint64_t* __restrict__ keys,
                                                                const float* __restrict__ dists, uint32_t parts,
                                                                uint64_t key_part_stride, uint64_t dist_part_stride,
                                                                uint64_t nq, uint32_t k, uint32_t n_pow2,
                                                                uint64_t* __restrict__ out_keys,
                                                                float* __restrict__ out_dists,
                                                                uint32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* sd = reinterpret_cast<uint32_t*>(sk + n_pow2);
    const uint64_t q = blockIdx.x;
    const uint32_t total = parts * k;
    for (uint32_t i = threadIdx.x; i < n_pow2; i += K8_THREADS) {
        uint64_t key = 0xFFFFFFFFFFFFFFFFull;
        uint32_t d = 0xFFFFFFFFu;
        if (i < total) {
            const uint32_t p = i / k, j = i % k;
            const size_t in_part = (size_t)q * k + j;
            key = keys[(size_t)p * key_part_stride + in_part];
            d = key == 0xFFFFFFFFFFFFFFFFull ? 0xFFFFFFFFu : f32_to_ord(dists[(size_t)p * dist_part_stride + in_part]);
        }
        sk[i] = key;
        sd[i] = d;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= n_pow2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < n_pow2 / 2; t += K8_THREADS) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint32_t dl = sd[lo], dh = sd[hi];
                const uint64_t kl = sk[lo], kh = sk[hi];
                const bool swap = up ? dk_less(dh, kh, dl, kl) : dk_less(dl, kl, dh, kh);
                if (swap) {
                    sd[lo] = dh; sd[hi] = dl;
                    sk[lo] = kh; sk[hi] = kl;
                }
            }
            __syncthreads();
        }
    }
    // emit the first k DISTINCT keys: a row may legitimately arrive twice (graph result + brute-force tail while a
    // streaming insert is linking that row) and then carries the same canonical distance, so the copies are adjacent
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        uint32_t count = 0;
        for (uint32_t b = 0; b < n_pow2 && count < k; b += 32) {
            const uint32_t i = b + lane;
            const uint64_t key = i < n_pow2 ? sk[i] : 0xFFFFFFFFFFFFFFFFull;
            const bool valid = key != 0xFFFFFFFFFFFFFFFFull && !(i > 0 && sk[i - 1] == key);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
            const uint32_t pos = count + __popc(m & ((1u << lane) - 1));
            if (valid && pos < k) {
                out_keys[q * k + pos] = key;
                out_dists[q * k + pos] = ord_to_f32(sd[i]);
            }
            count += __popc(m);
        }
        if (count > k) count = k;
        for (uint32_t i = count + lane; i < k; i += 32) {
            out_keys[q * k + i] = 0xFFFFFFFFFFFFFFFFull;
            out_dists[q * k + i] = __int_as_float(0x7F800000);
        }
        if (out_counts != nullptr && lane == 0) out_counts[q] = count;
    }
}

__global__ void fill_empty_kernel(uint64_t* keys, float* dists, uint32_t* counts, uint64_t nq, uint32_t k) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * k) {
        keys[i] = 0xFFFFFFFFFFFFFFFFull;
        dists[i] = __int_as_float(0x7F800000);
    }
    if (counts != nullptr && i < nq) counts[i] = 0;
}

}  // namespace

void launch_merge_topk(const uint64_t* keys, const float* dists, uint32_t parts, uint64_t q, uint32_t k,
                       uint64_t* out_keys, float* out_dists, uint32_t* out_counts, cudaStream_t stream,
                       uint64_t key_part_stride, uint64_t dist_part_stride) {
    if (key_part_stride == 0) key_part_stride = q * k;    // parts stored back to back
    if (dist_part_stride == 0) dist_part_stride = q * k;
    if (q == 0 || k == 0) return;
    uint32_t n_pow2 = 2;
    while (n_pow2 < parts * k) n_pow2 <<= 1;
    const size_t smem = (size_t)n_pow2 * 12;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_topk_kernel<<<(unsigned)q, K8_THREADS, smem, stream>>>(keys, dists, parts, key_part_stride, dist_part_stride, q, k, n_pow2, out_keys,
                                                                  out_dists, out_counts);
    g_kernel_launches += 1;
}

void launch_fill_empty(uint64_t* keys, float* dists, uint32_t* counts, uint64_t q, uint32_t k, cudaStream_t stream) {
    const size_t n = q * k > q ? q * k : q;
    if (n == 0) return;
    fill_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, dists, counts, q, k);
    g_kernel_launches += 1;
}

}  // namespace vsb
