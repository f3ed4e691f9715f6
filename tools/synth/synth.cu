// synth.cu — on-device generator of the synthetic corpora of SURVEY §8d (bench / test tooling, NOT part of libvsb200).
//
// C3 (10 M x 768) takes ~100 s to draw on the host and C4 (1 B x 128) must never exist in host memory, so rows are
// generated where they are consumed.  The generator is COUNTER-BASED: element (row r, column j) depends only on
// (seed, r, j) through a splitmix64 finaliser, so any shard of any world size produces exactly its rows, and
// vector_store_b200.host.datasets.embedding_mix() reproduces them bit for bit in NumPy (integer arithmetic for the
// randomness, individually rounded fp32 operations in a fixed order for the rest) — that twin feeds the CPU arm and
// the oracle.
//   cluster(r)  = h(seed ^ K_CLUSTER, r, 0) mod clusters
//   g(h)        = ((sum of the four 16-bit fields of h) - 131070) * INV_STD          (Irwin-Hall(4) ~ N(0,1))
//   x[j]        = g(h(centers_seed, cluster, j + 1)) + sigma * g(h(seed, r, j + 1))
//   row         = x / sqrt(sum x^2), sum taken lane-strided (j mod 32) then by the xor butterfly 16,8,4,2,1
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr uint64_t K_A = 0x9E3779B97F4A7C15ull, K_B = 0xD1B54A32D192ED03ull, K_CLUSTER = 0xC1057E7ull;
constexpr float INV_STD = 0x1.bb67aep-16f;  // 1 / sqrt(4 * (65536^2 - 1) / 12)

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t h3(uint64_t seed, uint64_t a, uint64_t b) { return mix64(seed + a * K_A + b * K_B); }
__device__ inline float gauss(uint64_t h) {
    const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48) - 131070;
    return __fmul_rn((float)s, INV_STD);
}

// one warp per row
__global__ void __launch_bounds__(256) embedding_mix_kernel(float* __restrict__ out, uint64_t row0, uint64_t n, uint32_t dim,
                                                            uint32_t clusters, float sigma, uint64_t seed,
                                                            uint64_t centers_seed) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint64_t r = row0 + i;
    const uint64_t c = h3(seed ^ K_CLUSTER, r, 0) % clusters;
    float acc = 0.0f;
    float* row = out + i * dim;
    for (uint32_t j = lane; j < dim; j += 32) {
        const float x = __fadd_rn(gauss(h3(centers_seed, c, j + 1)), __fmul_rn(sigma, gauss(h3(seed, r, j + 1))));
        row[j] = x;
        acc = __fadd_rn(acc, __fmul_rn(x, x));
    }
    for (int s = 16; s > 0; s >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, s));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(acc));
    __syncwarp();
    for (uint32_t j = lane; j < dim; j += 32) row[j] = __fmul_rn(row[j], inv);
}

}  // namespace

extern "C" int vsbsynth_embedding_mix(float* d_out, uint64_t row0, uint64_t n, uint32_t dim, uint32_t clusters, float sigma,
                                      uint64_t seed, uint64_t centers_seed, void* stream) {
    if (n == 0) return 0;
    if (clusters == 0 || dim == 0) return 1;
    const unsigned blocks = (unsigned)((n + 7) / 8);
    embedding_mix_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, row0, n, dim, clusters, sigma, seed,
                                                                                centers_seed);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// ---- random-row gather probe (tools/rowgather_probe.py): the HBM ceiling of K4's access pattern ---------------------
// Every warp streams whole rows chosen by a hash of (warp, step) with 128-bit loads, ROWS rows in flight per warp, and
// does nothing else with them (xor into a register).  What it reaches for a given row size is the bandwidth a beam
// search over rows of that size can at best reach: random rows open a DRAM page for a few hundred bytes each.
namespace {
template <int ROWS>
__global__ void __launch_bounds__(256) row_gather_kernel(const uint4* __restrict__ base, uint64_t n_rows, uint32_t row_chunks,
                                                         uint32_t steps, uint64_t seed, uint32_t* __restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (uint32_t s = 0; s < steps; ++s) {
        uint4 v[ROWS][6];
        const uint32_t cpl = (row_chunks + 31) / 32;  // <= 6
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t row = h3(seed, warp, (uint64_t)s * ROWS + r) % n_rows;
            const uint4* p = base + row * row_chunks + lane;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                v[r][j] = make_uint4(0, 0, 0, 0);
                if (j < (int)cpl && j * 32 + lane < (int)row_chunks)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[r][j].x), "=r"(v[r][j].y), "=r"(v[r][j].z), "=r"(v[r][j].w)
                                 : "l"(p + j * 32));
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc ^= v[r][j].x ^ v[r][j].y ^ v[r][j].z ^ v[r][j].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;  // keeps the loads alive
}
}  // namespace

// rows of row_bytes (multiple of 16, <= 3072) in a buffer of n_rows rows; every warp gathers steps * rows_in_flight rows
extern "C" int vsbsynth_row_gather(const void* d_rows, uint64_t n_rows, uint32_t row_bytes, uint32_t rows_in_flight,
                                   uint32_t warps, uint32_t steps, uint64_t seed, void* d_sink, void* stream) {
    if (row_bytes % 16 || row_bytes > 3072 || n_rows == 0) return 1;
    const uint32_t chunks = row_bytes / 16;
    const unsigned blocks = (warps + 7) / 8;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint4* b = static_cast<const uint4*>(d_rows);
    uint32_t* sink = static_cast<uint32_t*>(d_sink);
    switch (rows_in_flight) {
        case 1: row_gather_kernel<1><<<blocks, 256, 0, s>>>(b, n_rows, chunks, steps, seed, sink); break;
        case 2: row_gather_kernel<2><<<blocks, 256, 0, s>>>(b, n_rows, chunks, steps, seed, sink); break;
        case 4: row_gather_kernel<4><<<blocks, 256, 0, s>>>(b, n_rows, chunks, steps, seed, sink); break;
        case 8: row_gather_kernel<8><<<blocks, 256, 0, s>>>(b, n_rows, chunks, steps, seed, sink); break;
        default: return 1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
