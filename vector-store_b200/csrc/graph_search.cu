// graph_search.cu — host launcher of K4 (kernel in graph_search.cuh, instantiated per storage scalar).
#include "graph_search.cuh"

namespace vsb {

static int pick_cpl(int n_chunks) {
    const int need = (n_chunks + 31) / 32;
    const int set[7] = {1, 2, 3, 4, 6, 8, 12};
    for (int i = 0; i < 7; ++i)
        if (set[i] >= need) return set[i];
    return -1;
}

bool graph_search_supported(uint32_t row_bytes) { return pick_cpl((int)(row_bytes / 16)) > 0; }

void launch_graph_search(const SearchParams& p, cudaStream_t stream) {
    if (p.q.n == 0) return;
    K4Args a;
    a.q_rows = p.q.rows; a.q_nrm = p.q.nrm; a.nq = p.q.n; a.q_row_bytes = p.q.row_bytes;
    a.x_rows = p.x.rows; a.x_nrm = p.x.nrm; a.x_row_bytes = p.x.row_bytes;
    a.graph = p.graph; a.graph_stride = p.graph_stride; a.degree = p.degree;
    a.seed_lists = p.seed_lists; a.seed_splits = p.seed_stride / 32; a.n_seeds = p.n_seeds;
    a.seed_slots = p.seed_slots; a.deny = p.deny; a.keys = p.keys;
    a.itopk = ((p.itopk + 31) / 32) * 32;
    if (a.itopk < ((p.k + 31) / 32) * 32) a.itopk = ((p.k + 31) / 32) * 32;
    a.search_width = p.search_width < 1 ? 1 : (p.search_width > 4 ? 4 : p.search_width);
    a.max_iters = p.max_iters ? p.max_iters : (2 * a.itopk) / a.search_width + 8;
    a.k = p.k;
    // Visited hash: 16 KB per warp (4096 slots) = 3 CTAs x 4 query warps per SM, the same residency the
    // register budget allows; it is reset to "what is still in the list" when 3/4 full (forgotten nodes
    // cost a re-evaluation, never a duplicate).
    const uint32_t bits = 12;
    a.hash_bits = bits;
    const uint32_t deg_pad = ((p.degree + 31) / 32) * 32;
    a.queue_cap = a.search_width * deg_pad < 32 ? 32 : a.search_width * deg_pad;
    a.metric = p.metric;
    a.out_keys = p.out_keys; a.out_dists = p.out_dists; a.out_counts = p.out_counts; a.counters = p.counters;
    const int cpl = pick_cpl((int)(p.x.row_bytes / 16));
    dim3 grid((p.q.n + K4_WARPS - 1) / K4_WARPS);
    const size_t smem = (size_t)K4_WARPS * ((size_t)a.itopk * 8 + ((size_t)4 << bits) + (size_t)a.queue_cap * 8);
    switch (p.storage) {
        case VSB_ST_F32: launch_k4_f32(a, cpl, grid, smem, stream); break;
        case VSB_ST_F16: launch_k4_f16(a, cpl, grid, smem, stream); break;
        case VSB_ST_BF16: launch_k4_bf16(a, cpl, grid, smem, stream); break;
        case VSB_ST_I8: launch_k4_i8(a, cpl, grid, smem, stream); break;
        default: launch_k4_b1(a, cpl, grid, smem, stream); break;
    }
    g_kernel_launches += 1;
}

}  // namespace vsb
