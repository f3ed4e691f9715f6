// select.cuh — K2: warp-shuffle bitonic top-k primitives.
//
// Lists are arrays of packed u64 (ordered-distance-bits << 32 | slot), ascending, whose length is a
// multiple of 32; one warp owns a list.  New candidates arrive 32 at a time (one per lane), are
// sorted with a shuffle bitonic network and folded into the list block by block ("carry merge"):
// after folding block b the block holds the 32 smallest of (old block ∪ carry) and the carry the
// 32 largest, which is exactly what the next block needs because blocks are ordered.
#pragma once
#include "common.cuh"

namespace vsb {

// Tie-break policies. Less(a,b) must be a strict weak order on packed values.
struct LessBySlot {
    __device__ __forceinline__ bool operator()(uint64_t a, uint64_t b) const { return a < b; }
};
// (distance, key): looks the u64 key up only when the distance bits are equal.
struct LessByKey {
    const uint64_t* __restrict__ keys;
    __device__ __forceinline__ bool operator()(uint64_t a, uint64_t b) const {
        uint32_t ha = packed_hi(a), hb = packed_hi(b);
        if (ha != hb) return ha < hb;
        uint32_t sa = packed_lo(a), sb = packed_lo(b);
        if (sa == sb) return false;
        if (sa == kInvalidSlot) return false;
        if (sb == kInvalidSlot) return true;
        if (keys == nullptr) return sa < sb;
        return keys[sa] < keys[sb];
    }
};

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    return __shfl_xor_sync(kFullMask, (unsigned long long)v, m);
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    return __shfl_sync(kFullMask, (unsigned long long)v, src);
}

template <class Less>
__device__ __forceinline__ uint64_t cmpx(uint64_t v, uint64_t o, bool take_min, Less less) {
    bool o_less = less(o, v);
    return take_min ? (o_less ? o : v) : (o_less ? v : o);
}

// full bitonic sort of one value per lane, ascending by lane
template <class Less>
__device__ __forceinline__ uint64_t warp_sort32(uint64_t v, int lane, Less less) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t o = shfl_xor_u64(v, j);
            bool up = (k == 32) ? true : ((lane & k) == 0);
            bool lower = (lane & j) == 0;
            v = cmpx(v, o, lower == up, less);
        }
    }
    return v;
}

// sorts a bitonic sequence (one value per lane) ascending
template <class Less>
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t v, int lane, Less less) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        uint64_t o = shfl_xor_u64(v, j);
        v = cmpx(v, o, (lane & j) == 0, less);
    }
    return v;
}

// Folds 32 sorted candidates (`carry`, ascending by lane) into the sorted list `list[0..len)`,
// len % 32 == 0.  The list stays sorted; the 32 largest of the union fall off the end.
template <class Less>
__device__ __forceinline__ void warp_list_merge(uint64_t* list, int len, uint64_t carry, int lane, Less less) {
    for (int b = 0; b < len; b += 32) {
        uint64_t blk = list[b + lane];
        // nothing in the carry beats this block's maximum -> block unchanged, carry unchanged
        uint64_t blk_max = shfl_u64(blk, 31);
        uint64_t carry_min = shfl_u64(carry, 0);
        if (!less(carry_min, blk_max)) continue;
        uint64_t rc = shfl_u64(carry, 31 - lane);
        bool rc_less = less(rc, blk);
        uint64_t lo = rc_less ? rc : blk;
        uint64_t hi = rc_less ? blk : rc;
        lo = warp_bitonic_merge32(lo, lane, less);
        hi = warp_bitonic_merge32(hi, lane, less);
        list[b + lane] = lo;
        carry = hi;
    }
    __syncwarp();
}

}  // namespace vsb
