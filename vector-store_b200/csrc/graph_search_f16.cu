// graph_search_f16.cu — K4 / K4b / seed-scan instantiations for the f16 storage scalar (see graph_search.cuh).
#include "graph_search.cuh"

namespace vsb {
void launch_k4_f16(const K4Args& a, int cpl, dim3 grid, size_t smem, cudaStream_t stream) {
    launch_k4_storage<VSB_ST_F16>(a, cpl, grid, smem, stream);
}
void launch_k4b_f16(const K4Args& a, int cpl, dim3 grid, size_t smem, cudaStream_t stream) {
    launch_k4b_storage<VSB_ST_F16>(a, cpl, grid, smem, stream);
}
void launch_seed_scan_f16(const SeedScanArgs& s, cudaStream_t stream) { launch_seed_scan_storage<VSB_ST_F16>(s, stream); }
}  // namespace vsb
