// graph_search_i8.cu — K4 / K4b / seed-scan instantiations for the i8 storage scalar (see graph_search.cuh).
#include "graph_search.cuh"

namespace vsb {
void launch_k4_i8(const K4Args& a, int cpl, dim3 grid, size_t smem, cudaStream_t stream) {
    launch_k4_storage<VSB_ST_I8>(a, cpl, grid, smem, stream);
}
void launch_k4b_i8(const K4Args& a, int cpl, dim3 grid, size_t smem, cudaStream_t stream) {
    launch_k4b_storage<VSB_ST_I8>(a, cpl, grid, smem, stream);
}
void launch_seed_scan_i8(const SeedScanArgs& s, cudaStream_t stream) { launch_seed_scan_storage<VSB_ST_I8>(s, stream); }
}  // namespace vsb
