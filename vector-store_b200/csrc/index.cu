// index.cu — host side of libvsb200: the index object behind the C ABI of include/vsb200.h.
//
// Mirrors what `ThreadedUsearchIndex` + `usearch::Index` are to the reference
// (crates/vector-store/src/vs_index/usearch.rs:162-251): opaque u64 keys, unique keys
// (multi=false), explicit capacity (`reserve`), add / remove / search, and a live count.
//
// HBM layout (all sized by `capacity`, grown by vsb_reserve with a stream-ordered copy):
//   rows   [cap][row_bytes]  storage-typed vectors, rows padded to 16 bytes
//   sq,nrm [cap] f32         canonical sum of squares and its sqrt (cosine / L2 norm trick)
//   keys   [cap] u64         slot -> key
//   deny   [cap/32] u32      tombstone bitmap (remove() never moves rows)
//   graph  [n_graphed][stride] u32   fixed-degree neighbour rows (built by vsb_build)
//   seed_* [S]               contiguous copy of the entry-point sample ("upper layer")
// Slots are append-only; slots >= n_graphed form the brute-force tail, so an added vector is
// searchable as soon as vsb_add returns (dropping AsyncInProgress promises that).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/vsb200.h"
#include "kernels.h"

namespace vsb {
std::atomic<uint64_t> g_kernel_launches{0};
std::atomic<uint64_t> g_tc_launches{0};
}

namespace {

thread_local std::string g_last_error;

vsb_status fail(vsb_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return st;
}

#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? VSB_EOOM : VSB_ECUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                  \
    } while (0)

#define ST(expr)                        \
    do {                                \
        vsb_status s_ = (expr);         \
        if (s_ != VSB_OK) return s_;    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) {
        o.p = nullptr;
        o.bytes = 0;
    }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            bytes = o.bytes;
            o.p = nullptr;
            o.bytes = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    // grow-only scratch (contents not preserved)
    cudaError_t ensure(size_t want) {
        if (want <= bytes) return cudaSuccess;
        release();
        size_t sz = want + want / 4;
        cudaError_t e = cudaMalloc(&p, sz);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        bytes = sz;
        return cudaSuccess;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

struct vsb_index {
    std::mutex mu;
    vsb_options opt{};
    uint32_t dim = 0, row_bytes = 0;
    int metric = 0, storage = 0, device = 0, sm_count = 148;
    uint32_t degree = 32, graph_stride = 32, k_init = 64;
    uint32_t itopk = 64, max_iters = 0, n_seeds = 32, min_graph_size = 4096, search_width = 1;
    bool instrumented = false;

    cudaStream_t stream = nullptr;
    cudaStream_t last_stream = nullptr;

    uint64_t capacity = 0;
    uint32_t n_slots = 0, n_graphed = 0, n_seed_rows = 0;
    uint64_t live = 0;
    std::atomic<uint64_t> live_atomic{0};
    std::atomic<uint64_t> capacity_atomic{0};
    bool any_tombstone = false;

    DevBuf rows, sq, nrm, keys, deny, graph;
    DevBuf rows16, sq16, nrm16;   // VSB_FLAG_BF16_TRAVERSAL: bf16 copy of the rows for K4
    bool trav16 = false;
    uint32_t row_bytes16 = 0;
    // VSB_FLAG_I8_TRAVERSAL (f32 storage, cosine): K4 walks a scaled-int8 copy (a quarter of the f32 bytes, dp4a),
    // K3 re-ranks rr_mult8 * k candidates on the f32 rows.  The bf16 copy stays for the build and the seed tiles.
    DevBuf rows8, sq8, nrm8, q8_rows, q8_sq, q8_nrm;
    DevBuf reach_state;             // sample_seeds: reachability of the graph from the seed set
    bool reach_fix = true;          // VSB_DISABLE_REACH_FIX=1 turns the extra seeds off
    uint32_t reach_budget = 1024;   // at most this many extra seeds (one per unreached component)
    uint32_t n_extra_seeds = 0;
    bool trav8 = false;
    uint32_t row_bytes8 = 0, rr_mult8 = 4;
    DevBuf rr_packed;             // K4 -> K3 hand-over of the traversal shadow path
    DevBuf seed_rows, seed_sq, seed_nrm, seed_slots;
    DevBuf seed16_rows, seed16_sq, seed16_nrm;   // bf16 shadow of the seed block (f32 storage only)
    DevBuf q16_rows, q16_sq, q16_nrm;            // bf16 shadow of the converted queries (f32 storage only)
    DevBuf q_in, q_rows, q_sq, q_nrm, part, seed_part, tmp_keys, tmp_dists, counters, add_in, allow;
    // certified TF32 candidate stage for exact search on f32 rows (exact_block)
    DevBuf cert_state, fb_map, fb_rows, fb_sq, fb_nrm;
    bool cert_enabled = true;     // VSB_DISABLE_CERT=1: exact f32 search always on the SIMT tiles
    uint32_t cert_kp = 128;       // candidate list length of the certified TF32 stage (VSB_CERT_KP)
    uint32_t cert_kp16 = 32;      // ... of the certified f16/bf16 tensor-core stage (VSB_CERT_KP16)
    uint64_t cert_ok = 0, cert_fallback = 0, cert_scanned = 0;
    std::unordered_map<uint64_t, uint32_t> key2slot;
    std::vector<uint32_t> h_deny;

    uint64_t last_evals = 0, last_parents = 0, last_queries = 0;

    // optional CUDA-event timing of the search phases
    enum Phase { PH_CONVERT = 0, PH_SEED, PH_GRAPH, PH_EXACT, PH_MERGE, PH_COUNT };
    bool timing = false;
    struct Timed { cudaEvent_t a, b; int phase; };
    std::vector<Timed> timed;
    uint64_t phase_ns[PH_COUNT] = {0, 0, 0, 0, 0};
    uint64_t phase_launches[PH_COUNT] = {0, 0, 0, 0, 0};
    void t_begin(int phase, cudaStream_t s) {
        if (!timing) return;
        Timed t;
        t.phase = phase;
        cudaEventCreate(&t.a);
        cudaEventCreate(&t.b);
        cudaEventRecord(t.a, s);
        timed.push_back(t);
    }
    void t_end(cudaStream_t s) {
        if (!timing) return;
        cudaEventRecord(timed.back().b, s);
    }
    void t_resolve() {
        for (auto& t : timed) {
            float ms = 0.f;
            if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
                phase_ns[t.phase] += (uint64_t)((double)ms * 1e6);
                phase_launches[t.phase] += 1;
            }
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
        timed.clear();
    }

    size_t hbm_bytes() const {
        const DevBuf* all[] = {&rows, &sq, &nrm, &keys, &deny, &graph, &rows16, &sq16, &nrm16, &rr_packed, &seed_rows, &seed_sq, &seed_nrm, &seed_slots,
                               &seed16_rows, &seed16_sq, &seed16_nrm, &q16_rows, &q16_sq, &q16_nrm,
                               &q_in, &q_rows, &q_sq, &q_nrm, &part, &seed_part, &tmp_keys, &tmp_dists, &counters,
                               &add_in, &allow, &cert_state, &fb_map, &fb_rows, &fb_sq, &fb_nrm, &rows8, &sq8, &nrm8, &q8_rows,
                               &q8_sq, &q8_nrm, &reach_state};
        size_t s = 0;
        for (auto* b : all) s += b->bytes;
        return s;
    }

    vsb::RowsView corpus_view() const {
        vsb::RowsView v;
        v.rows = rows.as<uint8_t>();
        v.sq = sq.as<float>();
        v.nrm = nrm.as<float>();
        v.row_bytes = row_bytes;
        v.n = n_slots;
        return v;
    }

    vsb_status use_stream(cudaStream_t s) {
        if (last_stream != nullptr && last_stream != s) CU(cudaStreamSynchronize(last_stream));
        last_stream = s;
        return VSB_OK;
    }

    vsb_status reserve(uint64_t cap);
    vsb_status add(const uint64_t* k, const float* r, uint64_t n);
    vsb_status remove(const uint64_t* k, uint64_t n, uint64_t* removed);
    vsb_status build();
    vsb_status exact_block(const vsb::RowsView& q, const vsb::RowsView& x, uint32_t x_lo, uint32_t x_hi,
                           const uint32_t* deny_bm, const uint64_t* key_arr, const uint32_t* allow_bm,
                           uint64_t allow_bits, uint32_t k, uint64_t* out_keys, float* out_dists,
                           uint32_t* out_counts, uint64_t* out_packed, int64_t self_base, cudaStream_t s,
                           bool approx_ok, const vsb::RowsView* shadow_q = nullptr,
                           const vsb::RowsView* shadow_x = nullptr);
    bool tc_enabled = true;       // tcgen05 path for the dense distance tiles (VSB_DISABLE_TC=1 turns it off)
    uint32_t tc_min_rows = 8192;  // below this the SIMT K1 is used (launch + pipeline fill dominate)
    vsb_status graph_block(const vsb::RowsView& qv, uint32_t nb, uint32_t k, uint32_t itopk_eff, uint64_t* g_keys,
                           float* g_dists, uint32_t* counts_out, uint64_t* packed_out, const vsb::RowsView* q16_in,
                           cudaStream_t s, long long self_base = -1);
    vsb_status graph_from_knn(const uint64_t* knn, uint32_t n, uint32_t kin, const uint32_t* deny_bm);
    vsb_status refine_graph();
    uint32_t allpairs_prefix = 131072;  // rows of the exact all-pairs pass when n > allpairs_max
    vsb_status stream_insert();
    bool in_build = false;  // vsb_build runs its own refinement after streaming
    vsb_status compact();
    vsb_status sample_seeds(uint32_t n_rows);
    uint32_t allpairs_max = 262144;   // up to this many rows the graph comes from exact all-pairs kNN lists
    uint32_t refine_passes = 1;       // refinement passes after a streamed build
    uint64_t churn_since_refine = 0;  // rows streamed in + rows removed since the graph was last (re)built / refined
    uint32_t stream_threshold = 4096;  // un-graphed tail rows that trigger an automatic streaming insert
    vsb_status search_dev(const float* d_q, uint64_t nq, uint32_t k, uint64_t* d_keys, float* d_dists,
                          uint32_t* d_counts, cudaStream_t s, bool exact, const uint32_t* d_allow,
                          uint64_t allow_bits);
    vsb_status search_host(const float* queries, uint64_t nq, uint32_t k, uint64_t* keys_out, float* dists_out,
                           uint32_t* counts_out, bool exact, const uint32_t* allow_bitmap, uint64_t allow_bits);
};

static uint32_t storage_row_bytes(int storage, uint32_t dim) {
    uint64_t bits = 0;
    switch (storage) {
        case VSB_F32: bits = (uint64_t)dim * 32; break;
        case VSB_F16:
        case VSB_BF16: bits = (uint64_t)dim * 16; break;
        case VSB_I8: bits = (uint64_t)dim * 8; break;
        default: bits = dim; break;
    }
    return (uint32_t)(((bits + 127) / 128) * 16);
}

vsb_status vsb_index::reserve(uint64_t cap) {
    if (cap <= capacity) return VSB_OK;
    if (cap >= (1ull << 28)) return fail(VSB_EINVAL, "capacity %llu exceeds the 2^28 rows one shard holds", (unsigned long long)cap);
    CU(cudaSetDevice(device));
    const uint64_t words = (cap + 31) / 32;
    DevBuf n_rows, n_sq, n_nrm, n_keys, n_deny, n_rows16, n_sq16, n_nrm16, n_rows8, n_sq8, n_nrm8;
    auto alloc = [&](DevBuf& b, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(&b.p, bytes ? bytes : 16);
        if (e == cudaSuccess) b.bytes = bytes ? bytes : 16; else b.p = nullptr;
        return e;
    };
    if (trav8) {
        CU(alloc(n_rows8, cap * row_bytes8));
        CU(alloc(n_sq8, cap * 4));
        CU(alloc(n_nrm8, cap * 4));
    }
    CU(alloc(n_rows, cap * row_bytes));
    CU(alloc(n_sq, cap * 4));
    CU(alloc(n_nrm, cap * 4));
    CU(alloc(n_keys, cap * 8));
    CU(alloc(n_deny, words * 4));
    if (trav16) {
        CU(alloc(n_rows16, cap * row_bytes16));
        CU(alloc(n_sq16, cap * 4));
        CU(alloc(n_nrm16, cap * 4));
    }
    ST(use_stream(stream));
    CU(cudaMemsetAsync(n_deny.p, 0, words * 4, stream));
    if (n_slots > 0) {
        CU(cudaMemcpyAsync(n_rows.p, rows.p, (size_t)n_slots * row_bytes, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n_sq.p, sq.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n_nrm.p, nrm.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n_keys.p, keys.p, (size_t)n_slots * 8, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n_deny.p, deny.p, (size_t)((n_slots + 31) / 32) * 4, cudaMemcpyDeviceToDevice, stream));
        if (trav16) {
            CU(cudaMemcpyAsync(n_rows16.p, rows16.p, (size_t)n_slots * row_bytes16, cudaMemcpyDeviceToDevice, stream));
            CU(cudaMemcpyAsync(n_sq16.p, sq16.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
            CU(cudaMemcpyAsync(n_nrm16.p, nrm16.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
        }
        if (trav8) {
            CU(cudaMemcpyAsync(n_rows8.p, rows8.p, (size_t)n_slots * row_bytes8, cudaMemcpyDeviceToDevice, stream));
            CU(cudaMemcpyAsync(n_sq8.p, sq8.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
            CU(cudaMemcpyAsync(n_nrm8.p, nrm8.p, (size_t)n_slots * 4, cudaMemcpyDeviceToDevice, stream));
        }
    }
    CU(cudaStreamSynchronize(stream));
    if (trav8) {
        std::swap(rows8, n_rows8);
        std::swap(sq8, n_sq8);
        std::swap(nrm8, n_nrm8);
    }
    std::swap(rows, n_rows);
    std::swap(sq, n_sq);
    std::swap(nrm, n_nrm);
    std::swap(keys, n_keys);
    std::swap(deny, n_deny);
    if (trav16) {
        std::swap(rows16, n_rows16);
        std::swap(sq16, n_sq16);
        std::swap(nrm16, n_nrm16);
    }
    h_deny.resize(words, 0u);
    capacity = cap;
    capacity_atomic.store(cap);
    key2slot.reserve((size_t)cap);
    return VSB_OK;
}

vsb_status vsb_index::add(const uint64_t* k, const float* r, uint64_t n) {
    if (n == 0) return VSB_OK;
    if (k == nullptr || r == nullptr) return fail(VSB_EINVAL, "null keys/rows");
    if (live + n > capacity || (uint64_t)n_slots + n > capacity)
        return fail(VSB_EFULL, "size %llu + %llu exceeds capacity %llu: reserve capacity ahead of insertions",
                    (unsigned long long)n_slots, (unsigned long long)n, (unsigned long long)capacity);
    {
        std::unordered_set<uint64_t> batch;
        if (n > 1) batch.reserve((size_t)n);
        for (uint64_t i = 0; i < n; ++i) {
            if (k[i] == 0xFFFFFFFFFFFFFFFFull) return fail(VSB_EINVAL, "key UINT64_MAX is reserved");
            if (key2slot.count(k[i]) || (n > 1 && !batch.insert(k[i]).second))
                return fail(VSB_EDUPKEY, "duplicate key %llu", (unsigned long long)k[i]);
        }
    }
    CU(cudaSetDevice(device));
    ST(use_stream(stream));
    const uint64_t chunk_rows = std::max<uint64_t>(1, (256ull << 20) / ((uint64_t)dim * 4));
    for (uint64_t b = 0; b < n; b += chunk_rows) {
        const uint64_t nb = std::min(chunk_rows, n - b);
        CU(add_in.ensure(nb * dim * 4));
        CU(cudaMemcpyAsync(add_in.p, r + b * dim, nb * dim * 4, cudaMemcpyHostToDevice, stream));
        const uint32_t s0 = n_slots + (uint32_t)b;
        vsb::launch_convert_rows(storage, add_in.as<float>(), (uint32_t)nb, dim, rows.as<uint8_t>() + (size_t)s0 * row_bytes,
                                 row_bytes, sq.as<float>() + s0, nrm.as<float>() + s0, stream);
        CU(cudaGetLastError());
        if (trav16) {
            vsb::launch_convert_rows(VSB_BF16, add_in.as<float>(), (uint32_t)nb, dim,
                                     rows16.as<uint8_t>() + (size_t)s0 * row_bytes16, row_bytes16, sq16.as<float>() + s0,
                                     nrm16.as<float>() + s0, stream);
            CU(cudaGetLastError());
        }
        if (trav8) {
            vsb::launch_convert_rows_i8s(add_in.as<float>(), (uint32_t)nb, dim, dim, rows8.as<uint8_t>() + (size_t)s0 * row_bytes8,
                                         row_bytes8, sq8.as<float>() + s0, nrm8.as<float>() + s0, stream);
            CU(cudaGetLastError());
        }
        CU(cudaMemcpyAsync(keys.as<uint64_t>() + s0, k + b, nb * 8, cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));  // add_in is reused by the next chunk
    }
    for (uint64_t i = 0; i < n; ++i) key2slot.emplace(k[i], n_slots + (uint32_t)i);
    n_slots += (uint32_t)n;
    live += n;
    live_atomic.store(live);
    if (n_graphed > 0 && stream_threshold > 0 && n_slots - n_graphed >= stream_threshold) ST(stream_insert());
    return VSB_OK;
}

vsb_status vsb_index::remove(const uint64_t* k, uint64_t n, uint64_t* removed) {
    uint64_t cnt = 0;
    uint32_t lo_word = 0xFFFFFFFFu, hi_word = 0;
    for (uint64_t i = 0; i < n; ++i) {
        auto it = key2slot.find(k[i]);
        if (it == key2slot.end()) continue;
        const uint32_t slot = it->second;
        h_deny[slot >> 5] |= 1u << (slot & 31);
        lo_word = std::min(lo_word, slot >> 5);
        hi_word = std::max(hi_word, slot >> 5);
        key2slot.erase(it);
        ++cnt;
    }
    if (cnt) {
        CU(cudaSetDevice(device));
        ST(use_stream(stream));
        CU(cudaMemcpyAsync(deny.as<uint32_t>() + lo_word, h_deny.data() + lo_word, (size_t)(hi_word - lo_word + 1) * 4,
                           cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        live -= cnt;
        live_atomic.store(live);
        any_tombstone = true;
        churn_since_refine += cnt;
    }
    if (removed) *removed = cnt;
    return VSB_OK;
}

// exact top-k of the queries `q` against rows [x_lo, x_hi) of `x` (K1 + K3)
vsb_status vsb_index::exact_block(const vsb::RowsView& q, const vsb::RowsView& x, uint32_t x_lo, uint32_t x_hi,
                                  const uint32_t* deny_bm, const uint64_t* key_arr, const uint32_t* allow_bm,
                                  uint64_t allow_bits, uint32_t k, uint64_t* out_keys, float* out_dists,
                                  uint32_t* out_counts, uint64_t* out_packed, int64_t self_base, cudaStream_t s,
                                  bool approx_ok, const vsb::RowsView* shadow_q, const vsb::RowsView* shadow_x) {
    vsb::ExactParams p;
    p.storage = storage;
    p.metric = metric;
    p.q = q;
    p.x = x;
    p.x_lo = x_lo;
    p.x_hi = x_hi;
    p.deny = deny_bm;
    p.keys = key_arr;
    p.allow = allow_bm;
    p.allow_bits = allow_bits;
    const uint32_t extra = std::max<uint32_t>(16, k / 4) + (self_base >= 0 ? 1 : 0);
    p.kp = round_up(k + extra, 32);
    if (p.kp > 256) return fail(VSB_EINVAL, "k=%u too large for the exact path (max 200)", k);
    // Tensor-core tiles: 16-bit storages multiply exactly, f32 rows run as TF32.  Any tiled stage (tensor core
    // or SIMT) sums in its own order, so its lists are candidate-grade; for exact results on float storages K3
    // CERTIFIES each query (no dropped row can reach or tie into the canonical top-k) and the queries it
    // cannot certify fall through: TF32 tiles -> fp32 SIMT tiles -> canonical scan (K1c, needs no certificate).
    // Integer storages (i8, b1) are exact in every stage and ordered by (distance, key) throughout.
    const bool tc_shape = tc_enabled && vsb::exact_tc_supported(storage, metric) && (x_hi - x_lo) >= tc_min_rows;
    const bool is_float = storage == VSB_F32 || storage == VSB_F16 || storage == VSB_BF16;
    const bool certify = is_float && !approx_ok && cert_enabled;
    bool tc = tc_shape && (approx_ok || storage != VSB_F32 || certify);
    const uint32_t kp_simt = p.kp;
    if (certify && tc) p.kp = std::min<uint32_t>(256, std::max<uint32_t>(p.kp, round_up(storage == VSB_F32 ? cert_kp : cert_kp16, 32)));
    p.n_splits = tc ? vsb::exact_tc_pick_splits(q.n, x_hi - x_lo, sm_count)
                    : vsb::exact_pick_splits(q.n, x_hi - x_lo, sm_count);
    CU(part.ensure(vsb::exact_part_elems(q.n, p.n_splits, p.kp) * 8));
    p.part = part.as<uint64_t>();
    if (tc) {
        if (shadow_q != nullptr && shadow_x != nullptr && !certify) {
            // candidate stage on the bf16 shadow (half the bytes, kind::f16 rate); K3 re-ranks on the real rows
            vsb::ExactParams pc = p;
            pc.storage = VSB_BF16;
            pc.q = *shadow_q;
            pc.x = *shadow_x;
            tc = vsb::launch_exact_candidates_tc(pc, s);
        } else {
            tc = vsb::launch_exact_candidates_tc(p, s);
        }
    }
    if (!tc) {
        if (p.kp != kp_simt) {  // the tensor-core launch was refused: plain SIMT with its own list length
            p.kp = kp_simt;
            p.n_splits = vsb::exact_pick_splits(q.n, x_hi - x_lo, sm_count);
            CU(part.ensure(vsb::exact_part_elems(q.n, p.n_splits, p.kp) * 8));
            p.part = part.as<uint64_t>();
        }
        vsb::launch_exact_candidates(p, s);
    }
    CU(cudaGetLastError());
    if (!certify) {
        vsb::launch_exact_rerank(p, k, out_keys, out_dists, out_counts, out_packed, self_base, s);
        CU(cudaGetLastError());
        return VSB_OK;
    }

    // cert_state: [0] = max row norm of the block, [1] = number of flagged queries, [2..] flags
    CU(cert_state.ensure((size_t)(q.n + 2) * 4));
    float* d_xmax = cert_state.as<float>();
    uint32_t* d_count = cert_state.as<uint32_t>() + 1;
    uint32_t* d_flags = cert_state.as<uint32_t>() + 2;
    vsb::launch_max_norm(x.nrm, x_lo, x_hi, d_xmax, s);
    vsb::ExactCert cert;
    cert.x_nrm_max = d_xmax;
    cert.flags = d_flags;
    cert.count = d_count;
    // An fp32 sum of `dim` products, in any order, is within dim * 2^-24 |q||x| of the real dot product (2^-23 per
    // add if the adder truncates); candidate and canonical evaluation together: dim * 2^-22 with margin.
    // TF32 additionally keeps only 10 mantissa bits of each operand: <= 2^-10 relative each, 2^-9 on the product.
    const float rel_fp32 = (float)dim * 0x1p-22f;
    const float rel_tf32 = 1.25f * 0x1p-9f + rel_fp32;
    cert.sum = (float)dim * 0x1p-23f;
    cert.rel = (tc && storage == VSB_F32) ? rel_tf32 : rel_fp32;

    // flagged queries of one stage -> compact map (indices into the caller's query block) + gathered rows
    std::vector<uint32_t> map, flags;
    auto collect = [&](uint32_t n_stage, const std::vector<uint32_t>* prev) -> vsb_status {
        flags.resize(n_stage);
        CU(cudaMemcpyAsync(flags.data(), d_flags, (size_t)n_stage * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        std::vector<uint32_t> next;
        for (uint32_t i = 0; i < n_stage; ++i)
            if (flags[i]) next.push_back(prev ? (*prev)[i] : i);
        map.swap(next);
        const uint32_t nf = (uint32_t)map.size();
        CU(fb_map.ensure((size_t)nf * 4));
        CU(fb_rows.ensure((size_t)nf * q.row_bytes));
        CU(fb_sq.ensure((size_t)nf * 4));
        CU(fb_nrm.ensure((size_t)nf * 4));
        CU(cudaMemcpyAsync(fb_map.p, map.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, s));
        vsb::launch_gather_rows(q.rows, q.row_bytes, q.sq, q.nrm, fb_map.as<uint32_t>(), nf, fb_rows.as<uint8_t>(),
                                fb_sq.as<float>(), fb_nrm.as<float>(), s);
        CU(cudaStreamSynchronize(s));  // `map` is pageable and is rebuilt by the next stage
        return VSB_OK;
    };
    auto read_count = [&](uint32_t* out) -> vsb_status {
        CU(cudaMemcpyAsync(out, d_count, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return VSB_OK;
    };

    CU(cudaMemsetAsync(d_count, 0, 4, s));
    vsb::launch_exact_rerank(p, k, out_keys, out_dists, out_counts, out_packed, self_base, s, &cert);
    CU(cudaGetLastError());
    uint32_t n_flagged = 0;
    ST(read_count(&n_flagged));
    cert_ok += q.n - n_flagged;
    cert_fallback += n_flagged;
    if (n_flagged == 0) return VSB_OK;
    ST(collect(q.n, nullptr));

    vsb::ExactParams pf = p;
    pf.q.rows = fb_rows.as<uint8_t>();
    pf.q.sq = fb_sq.as<float>();
    pf.q.nrm = fb_nrm.as<float>();
    pf.q.n = (uint32_t)map.size();
    if (tc && storage == VSB_F32) {
        // second stage: full fp32 products on the SIMT tiles, same certificate with the fp32 bound
        pf.kp = kp_simt;
        pf.n_splits = vsb::exact_pick_splits(pf.q.n, x_hi - x_lo, sm_count);
        CU(part.ensure(vsb::exact_part_elems(pf.q.n, pf.n_splits, pf.kp) * 8));
        pf.part = part.as<uint64_t>();
        vsb::launch_exact_candidates(pf, s);
        CU(cudaGetLastError());
        cert.rel = rel_fp32;
        CU(cudaMemsetAsync(d_count, 0, 4, s));
        vsb::launch_exact_rerank(pf, k, out_keys, out_dists, out_counts, out_packed, self_base, s, &cert,
                                 fb_map.as<uint32_t>());
        CU(cudaGetLastError());
        ST(read_count(&n_flagged));
        if (n_flagged == 0) return VSB_OK;
        const std::vector<uint32_t> prev = map;
        ST(collect(pf.q.n, &prev));
        pf.q.n = (uint32_t)map.size();
    }

    // last stage: canonical scan of the whole block for what is left
    cert_scanned += pf.q.n;
    pf.kp = round_up(k + (self_base >= 0 ? 1 : 0), 32);
    const uint32_t scan_splits = vsb::exact_scan_pick_splits(pf.q.n, x_hi - x_lo, sm_count);
    pf.n_splits = scan_splits;
    const uint32_t lists = vsb::exact_scan_lists_per_query(scan_splits);
    CU(part.ensure(vsb::exact_part_elems(pf.q.n, lists, pf.kp) * 8));
    pf.part = part.as<uint64_t>();
    vsb::launch_exact_scan(pf, k, s);
    CU(cudaGetLastError());
    pf.n_splits = lists;
    vsb::launch_exact_rerank(pf, k, out_keys, out_dists, out_counts, out_packed, self_base, s, nullptr,
                             fb_map.as<uint32_t>());
    CU(cudaGetLastError());
    return VSB_OK;
}

// Tombstone compaction (part of vsb_build): live rows are gathered to the front in slot order, the key map is
// rebuilt and the tombstone bitmap cleared, so a long-running delete/update stream does not leak HBM.
vsb_status vsb_index::compact() {
    std::vector<uint32_t> live_slots;
    live_slots.reserve((size_t)live);
    for (uint32_t i = 0; i < n_slots; ++i)
        if (!(h_deny[i >> 5] >> (i & 31) & 1u)) live_slots.push_back(i);
    const uint32_t m = (uint32_t)live_slots.size();
    if (m != n_slots) {
        DevBuf d_slots, n_rows, n_sq, n_nrm, n_keys, n_rows16, n_sq16, n_nrm16, n_rows8, n_sq8, n_nrm8;
        CU(d_slots.ensure(std::max<size_t>((size_t)m * 4, 16)));
        CU(cudaMemcpyAsync(d_slots.p, live_slots.data(), (size_t)m * 4, cudaMemcpyHostToDevice, stream));
        auto alloc = [&](DevBuf& b, size_t bytes) -> cudaError_t {
            cudaError_t e = cudaMalloc(&b.p, bytes ? bytes : 16);
            if (e == cudaSuccess) b.bytes = bytes ? bytes : 16; else b.p = nullptr;
            return e;
        };
        CU(alloc(n_rows, (size_t)capacity * row_bytes));
        CU(alloc(n_sq, (size_t)capacity * 4));
        CU(alloc(n_nrm, (size_t)capacity * 4));
        CU(alloc(n_keys, (size_t)capacity * 8));
        vsb::launch_gather_rows(rows.as<uint8_t>(), row_bytes, sq.as<float>(), nrm.as<float>(), d_slots.as<uint32_t>(), m,
                                n_rows.as<uint8_t>(), n_sq.as<float>(), n_nrm.as<float>(), stream);
        vsb::launch_gather_u64(keys.as<uint64_t>(), d_slots.as<uint32_t>(), m, n_keys.as<uint64_t>(), stream);
        if (trav16) {
            CU(alloc(n_rows16, (size_t)capacity * row_bytes16));
            CU(alloc(n_sq16, (size_t)capacity * 4));
            CU(alloc(n_nrm16, (size_t)capacity * 4));
            vsb::launch_gather_rows(rows16.as<uint8_t>(), row_bytes16, sq16.as<float>(), nrm16.as<float>(),
                                    d_slots.as<uint32_t>(), m, n_rows16.as<uint8_t>(), n_sq16.as<float>(),
                                    n_nrm16.as<float>(), stream);
        }
        if (trav8) {
            CU(alloc(n_rows8, (size_t)capacity * row_bytes8));
            CU(alloc(n_sq8, (size_t)capacity * 4));
            CU(alloc(n_nrm8, (size_t)capacity * 4));
            vsb::launch_gather_rows(rows8.as<uint8_t>(), row_bytes8, sq8.as<float>(), nrm8.as<float>(), d_slots.as<uint32_t>(),
                                    m, n_rows8.as<uint8_t>(), n_sq8.as<float>(), n_nrm8.as<float>(), stream);
        }
        CU(cudaGetLastError());
        std::vector<uint64_t> h_keys(m);
        CU(cudaMemcpyAsync(h_keys.data(), n_keys.p, (size_t)m * 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemsetAsync(deny.p, 0, deny.bytes, stream));
        CU(cudaStreamSynchronize(stream));
        std::swap(rows, n_rows);
        std::swap(sq, n_sq);
        std::swap(nrm, n_nrm);
        std::swap(keys, n_keys);
        if (trav16) {
            std::swap(rows16, n_rows16);
            std::swap(sq16, n_sq16);
            std::swap(nrm16, n_nrm16);
        }
        if (trav8) {
            std::swap(rows8, n_rows8);
            std::swap(sq8, n_sq8);
            std::swap(nrm8, n_nrm8);
        }
        std::fill(h_deny.begin(), h_deny.end(), 0u);
        key2slot.clear();
        for (uint32_t i = 0; i < m; ++i) key2slot.emplace(h_keys[i], i);
        n_slots = m;
        n_graphed = 0;
        n_seed_rows = 0;
    }
    any_tombstone = false;
    return VSB_OK;
}

vsb_status vsb_index::build() {
    CU(cudaSetDevice(device));
    ST(use_stream(stream));
    if (any_tombstone) ST(compact());
    struct Flag {
        bool& f;
        explicit Flag(bool& r) : f(r) { f = true; }
        ~Flag() { f = false; }
    } building(in_build);
    churn_since_refine = 0;
    if (live < min_graph_size || !vsb::graph_search_supported(row_bytes)) {
        n_graphed = 0;
        n_seed_rows = 0;
        return VSB_OK;
    }
    // All-pairs kNN lists cost 2*n^2*D flop: above `allpairs_max` rows only the first `allpairs_max` rows are
    // built that way and the remaining rows are linked in with the streaming insert (K7, O(n log n)).
    const uint32_t n = n_slots <= allpairs_max ? n_slots : std::min<uint32_t>(n_slots, allpairs_prefix);
    const uint32_t R = degree;
    const uint32_t kin = std::min<uint32_t>(k_init, 128);
    DevBuf knn;
    CU(knn.ensure((size_t)n * kin * 8));
    const vsb::RowsView x = corpus_view();
    const uint32_t* deny_bm = any_tombstone ? deny.as<uint32_t>() : nullptr;
    // f32 storage: the all-pairs candidate stage runs on a temporary bf16 copy (kNN lists only need
    // candidate-grade distances; K3 re-evaluates the survivors on the f32 rows in the canonical order)
    DevBuf sh_rows, sh_sq, sh_nrm;
    vsb::RowsView shx;
    const bool use_shadow = storage == VSB_F32 && tc_enabled && n >= tc_min_rows;
    if (use_shadow) {
        const uint32_t dim_pad = row_bytes / 4;
        shx.row_bytes = ((dim_pad * 2 + 15) / 16) * 16;
        shx.n = n;
        CU(sh_rows.ensure((size_t)n * shx.row_bytes));
        CU(sh_sq.ensure((size_t)n * 4));
        CU(sh_nrm.ensure((size_t)n * 4));
        vsb::launch_convert_rows(VSB_BF16, reinterpret_cast<const float*>(x.rows), n, dim_pad, sh_rows.as<uint8_t>(),
                                 shx.row_bytes, sh_sq.as<float>(), sh_nrm.as<float>(), stream);
        CU(cudaGetLastError());
        shx.rows = sh_rows.as<uint8_t>();
        shx.sq = sh_sq.as<float>();
        shx.nrm = sh_nrm.as<float>();
    }
    const bool btime = getenv("VSB_BUILD_TIMING") != nullptr;
    cudaEvent_t ev[2];
    if (btime) {
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], stream);
    }
    const uint32_t QB = 16384;
    for (uint32_t b0 = 0; b0 < n; b0 += QB) {
        vsb::RowsView q;
        q.n = std::min(QB, n - b0);
        q.rows = x.rows + (size_t)b0 * row_bytes;
        q.sq = x.sq + b0;
        q.nrm = x.nrm + b0;
        q.row_bytes = row_bytes;
        vsb::RowsView shq = shx;
        if (use_shadow) {
            shq.n = q.n;
            shq.rows = shx.rows + (size_t)b0 * shx.row_bytes;
            shq.sq = shx.sq + b0;
            shq.nrm = shx.nrm + b0;
        }
        ST(exact_block(q, x, 0, n, deny_bm, keys.as<uint64_t>(), nullptr, 0, kin, nullptr, nullptr, nullptr,
                       knn.as<uint64_t>() + (size_t)b0 * kin, (int64_t)b0, stream, true, use_shadow ? &shq : nullptr,
                       use_shadow ? &shx : nullptr));
    }
    sh_rows.release();
    sh_sq.release();
    sh_nrm.release();
    if (btime) {
        cudaEventRecord(ev[1], stream);
        cudaEventSynchronize(ev[1]);
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[0], ev[1]);
        fprintf(stderr, "[vsb200 build] exact all-pairs kNN lists (K1+K3) over %u rows: %.1f ms\n", n, t);
        cudaEventRecord(ev[0], stream);
    }
    ST(graph_from_knn(knn.as<uint64_t>(), n, kin, deny_bm));
    knn.release();
    ST(sample_seeds(n));
    if (btime) {
        cudaEventRecord(ev[1], stream);
        cudaEventSynchronize(ev[1]);
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[0], ev[1]);
        fprintf(stderr, "[vsb200 build] prune + reverse + merge + seeds: %.1f ms\n", t);
        cudaEventRecord(ev[0], stream);
    }
    if (n < n_slots) {
        // large index: link the remaining rows with the streaming insert (K7), then rebuild every row's
        // list from an ANN search over that navigable graph and prune it exactly like the all-pairs lists
        ST(stream_insert());
        ST(sample_seeds(n_slots));  // entry points drawn from every row, not only the all-pairs prefix
        if (btime) {
            cudaEventRecord(ev[1], stream);
            cudaEventSynchronize(ev[1]);
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[0], ev[1]);
            fprintf(stderr, "[vsb200 build] streaming insert of %u rows (K7): %.1f ms\n", n_slots - n, t);
            cudaEventRecord(ev[0], stream);
        }
        if (refine_passes > 0) {
            for (uint32_t pass = 0; pass < refine_passes; ++pass) ST(refine_graph());
            ST(sample_seeds(n_slots));  // reachability of the FINAL graph from the entry points
            if (btime) {
                cudaEventRecord(ev[1], stream);
                cudaEventSynchronize(ev[1]);
                float t = 0.f;
                cudaEventElapsedTime(&t, ev[0], ev[1]);
                fprintf(stderr, "[vsb200 build] %u refine pass(es) (K4 kNN lists + K6): %.1f ms\n", refine_passes, t);
            }
        }
    }
    if (btime)
        for (auto& e : ev) cudaEventDestroy(e);
    return VSB_OK;
}

// K6 pipeline: packed kNN lists [n][kin] -> fixed-degree graph rows (replaces `graph`, sets n_graphed = n)
vsb_status vsb_index::graph_from_knn(const uint64_t* knn, uint32_t n, uint32_t kin, const uint32_t* deny_bm) {
    const uint32_t R = degree;
    DevBuf fwd, rev, rev_cnt, scratch;
    CU(fwd.ensure((size_t)n * R * 4));
    CU(rev.ensure((size_t)n * R * 4));
    CU(rev_cnt.ensure((size_t)n * 4));
    vsb::launch_prune_detour(knn, n, kin, R, deny_bm, fwd.as<uint32_t>(), stream);
    CU(cudaGetLastError());
    const size_t sb = vsb::reverse_edges_scratch_bytes(n, R);
    CU(scratch.ensure(sb));
    vsb::launch_reverse_edges(fwd.as<uint32_t>(), n, R, rev.as<uint32_t>(), rev_cnt.as<uint32_t>(), scratch.p,
                              scratch.bytes, stream);
    CU(cudaGetLastError());
    scratch.release();
    DevBuf new_graph;
    CU(new_graph.ensure((size_t)std::max<uint64_t>(capacity, n) * graph_stride * 4));  // room for streamed rows
    vsb::launch_merge_graph(fwd.as<uint32_t>(), rev.as<uint32_t>(), rev_cnt.as<uint32_t>(), n, R,
                            new_graph.as<uint32_t>(), graph_stride, stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(stream));
    std::swap(graph, new_graph);
    n_graphed = n;
    return VSB_OK;
}

// One refinement pass over a complete (streamed) graph: every row searches the graph for its own
// k_init nearest rows (K4, beam = expansion_add) and the lists go through the K6 pipeline again.
vsb_status vsb_index::refine_graph() {
    const uint32_t n = n_graphed;
    if (n == 0) return VSB_OK;
    const uint32_t kin = std::min<uint32_t>(k_init, 128);
    const uint32_t ef_add = opt.expansion_add ? opt.expansion_add : 128;
    const uint32_t ef = std::min<uint32_t>(round_up(std::max(ef_add, kin + 1), 32), 512);
    const uint32_t* deny_bm = any_tombstone ? deny.as<uint32_t>() : nullptr;
    DevBuf knn;
    CU(knn.ensure((size_t)n * kin * 8));
    const uint32_t QB = 16384;
    for (uint32_t b0 = 0; b0 < n; b0 += QB) {
        const uint32_t nb = std::min(QB, n - b0);
        vsb::RowsView qv;
        qv.rows = rows.as<uint8_t>() + (size_t)b0 * row_bytes;
        qv.sq = sq.as<float>() + b0;
        qv.nrm = nrm.as<float>() + b0;
        qv.row_bytes = row_bytes;
        qv.n = nb;
        vsb::RowsView q16;
        if (trav16) {
            q16.rows = rows16.as<uint8_t>() + (size_t)b0 * row_bytes16;
            q16.sq = sq16.as<float>() + b0;
            q16.nrm = nrm16.as<float>() + b0;
            q16.row_bytes = row_bytes16;
            q16.n = nb;
        }
        ST(graph_block(qv, nb, kin, ef, nullptr, nullptr, nullptr, knn.as<uint64_t>() + (size_t)b0 * kin,
                       trav16 ? &q16 : nullptr, stream, (long long)b0));
    }
    ST(graph_from_knn(knn.as<uint64_t>(), n, kin, deny_bm));
    return VSB_OK;
}

// Entry-point sample ("upper layer"): a stride permutation of the live slots below n_rows
// (deterministic, seed-shifted), gathered into a contiguous block (+ bf16 shadow for f32 storage).
vsb_status vsb_index::sample_seeds(uint32_t n) {
    const vsb::RowsView x = corpus_view();
    uint64_t live_below = 0;
    for (uint32_t w = 0; w < (n + 31) / 32; ++w) live_below += __builtin_popcount(~h_deny[w]);
    if (n % 32) live_below -= 32 - (n % 32);
    // entry-point sample: a stride permutation of the live slots (deterministic, seed-shifted)
    uint32_t S = 256;
    const double target = 4.0 * std::sqrt((double)live_below);
    while (S < target && S < 8192) S <<= 1;
    if (S > live_below / 4) S = (uint32_t)std::max<uint64_t>(32, live_below / 4);
    std::vector<uint32_t> h_seeds;
    h_seeds.reserve(S);
    {
        uint64_t step = (uint64_t)((double)n * 0.6180339887498949) | 1ull;
        auto gcd = [](uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; };
        while (gcd(step, n) != 1) step += 2;
        uint64_t pos = opt.seed % n;
        for (uint32_t i = 0; i < n && h_seeds.size() < S; ++i) {
            const uint32_t slot = (uint32_t)pos;
            pos = (pos + step) % n;
            if (h_deny[slot >> 5] >> (slot & 31) & 1u) continue;
            h_seeds.push_back(slot);
        }
    }
    S = (uint32_t)h_seeds.size();
    // every live graph node must be reachable from the seed set: expand the frontier from the sample to a fixed
    // point, then promote the first unreached live node to an extra seed and continue, once per lost component
    uint32_t extra_seeds = 0;
    if (n_graphed >= n && n > 0 && S > 0 && reach_fix) {
        CU(reach_state.ensure((size_t)n + 16));
        uint8_t* st = reach_state.as<uint8_t>();
        uint32_t* flag = reinterpret_cast<uint32_t*>(st + (((size_t)n + 7) / 8) * 8);  // [0] changed, [1] first unreached
        CU(seed_slots.ensure((size_t)(S + reach_budget) * 4));
        CU(cudaMemsetAsync(st, 0, n, stream));
        CU(cudaMemcpyAsync(seed_slots.p, h_seeds.data(), (size_t)S * 4, cudaMemcpyHostToDevice, stream));
        vsb::launch_reach_mark(st, seed_slots.as<uint32_t>(), S, stream);
        const uint32_t* deny_bm = any_tombstone ? deny.as<uint32_t>() : nullptr;
        auto expand = [&]() -> vsb_status {
            for (int round = 0; round < 4096; ++round) {
                CU(cudaMemsetAsync(flag, 0, 4, stream));
                for (int i = 0; i < 4; ++i)
                    vsb::launch_reach_step(graph.as<uint32_t>(), n, graph_stride, degree, st, flag, stream);
                uint32_t changed = 0;
                CU(cudaMemcpyAsync(&changed, flag, 4, cudaMemcpyDeviceToHost, stream));
                CU(cudaStreamSynchronize(stream));
                if (!changed) break;
            }
            return VSB_OK;
        };
        ST(expand());
        while (extra_seeds < reach_budget) {
            CU(cudaMemsetAsync(flag + 1, 0xFF, 4, stream));
            vsb::launch_first_unreached(st, deny_bm, n, flag + 1, stream);
            uint32_t first = 0xFFFFFFFFu;
            CU(cudaMemcpyAsync(&first, flag + 1, 4, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            if (first == 0xFFFFFFFFu) break;
            h_seeds.push_back(first);
            ++extra_seeds;
            CU(cudaMemsetAsync(st + first, 1, 1, stream));
            ST(expand());
        }
        CU(cudaGetLastError());
        S = (uint32_t)h_seeds.size();
    }
    n_extra_seeds = extra_seeds;
    CU(seed_slots.ensure((size_t)S * 4));
    CU(seed_rows.ensure((size_t)S * row_bytes));
    CU(seed_sq.ensure((size_t)S * 4));
    CU(seed_nrm.ensure((size_t)S * 4));
    CU(cudaMemcpyAsync(seed_slots.p, h_seeds.data(), (size_t)S * 4, cudaMemcpyHostToDevice, stream));
    vsb::launch_gather_rows(x.rows, row_bytes, x.sq, x.nrm, seed_slots.as<uint32_t>(), S, seed_rows.as<uint8_t>(),
                            seed_sq.as<float>(), seed_nrm.as<float>(), stream);
    CU(cudaGetLastError());
    if (storage == VSB_F32) {
        const uint32_t dim_pad = row_bytes / 4;
        const uint32_t rb16 = ((dim_pad * 2 + 15) / 16) * 16;
        CU(seed16_rows.ensure((size_t)S * rb16));
        CU(seed16_sq.ensure((size_t)S * 4));
        CU(seed16_nrm.ensure((size_t)S * 4));
        vsb::launch_convert_rows(VSB_BF16, seed_rows.as<float>(), S, dim_pad, seed16_rows.as<uint8_t>(), rb16,
                                 seed16_sq.as<float>(), seed16_nrm.as<float>(), stream);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(stream));
    n_seed_rows = S;
    return VSB_OK;
}

// Seeds + K4 (+ K3 re-rank on the bf16-traversal path) for `nb` converted queries `qv`.
//   out_keys/out_dists/out_counts : user-facing top-k (nullable when packed_out is used)
//   packed_out                    : raw K4 list (packed ord(dist)<<32|slot, [nb][k]) — used by the streaming insert
//   q16_in                        : bf16 copy of the queries if the caller already has one (corpus rows), else built here
vsb_status vsb_index::graph_block(const vsb::RowsView& qv, uint32_t nb, uint32_t k, uint32_t itopk_eff,
                                  uint64_t* g_keys, float* g_dists, uint32_t* counts_out, uint64_t* packed_out,
                                  const vsb::RowsView* q16_in, cudaStream_t s, long long self_base) {
    const vsb::RowsView x = corpus_view();
    const uint32_t* deny_bm = any_tombstone ? deny.as<uint32_t>() : nullptr;
    const bool have_tail = counts_out == nullptr;  // the caller merges and counts later
    uint32_t* o_counts = counts_out;
    const bool rerank = trav16 && packed_out == nullptr;
    // ---- bf16 shadow of the queries (f32 storage: tensor-core seed layer and/or bf16 traversal) ----
    bool seed_tc = tc_enabled && vsb::exact_tc_supported(storage, metric) && nb >= 16;
    vsb::RowsView q16v;
    if (q16_in != nullptr) {
        q16v = *q16_in;
    } else if (storage == VSB_F32 && (seed_tc || trav16)) {
        CU(q16_rows.ensure((size_t)nb * row_bytes16));
        CU(q16_sq.ensure((size_t)nb * 4));
        CU(q16_nrm.ensure((size_t)nb * 4));
        vsb::launch_convert_rows(VSB_BF16, reinterpret_cast<const float*>(qv.rows), nb, row_bytes / 4, q16_rows.as<uint8_t>(),
                                 row_bytes16, q16_sq.as<float>(), q16_nrm.as<float>(), s);
        CU(cudaGetLastError());
        q16v.rows = q16_rows.as<uint8_t>();
        q16v.sq = q16_sq.as<float>();
        q16v.nrm = q16_nrm.as<float>();
        q16v.row_bytes = row_bytes16;
        q16v.n = nb;
    }
    // ---- seed layer: distances to the contiguous entry-point sample ----
    vsb::ExactParams sp;
    sp.storage = storage;
    sp.metric = metric;
    sp.q = qv;
    sp.x.rows = seed_rows.as<uint8_t>();
    sp.x.sq = seed_sq.as<float>();
    sp.x.nrm = seed_nrm.as<float>();
    sp.x.row_bytes = row_bytes;
    sp.x.n = n_seed_rows;
    sp.x_lo = 0;
    sp.x_hi = n_seed_rows;
    sp.keys = nullptr;  // ties fall back to the seed index (LessByKey with null keys)
    sp.kp = 32;
    const vsb::ExactParams sp_native = sp;
    if (seed_tc) {
        // tensor cores: one winner per 256-row tile per query (no list maintenance);
        // f32 storage multiplies the bf16 shadows of the queries and of the seed block
        sp.n_splits = std::max(vsb::exact_tc_pick_splits(nb, n_seed_rows, sm_count),
                               vsb::exact_tc_min_splits_tile_min(n_seed_rows, 32));
        if (storage == VSB_F32) {
            sp.storage = VSB_BF16;
            sp.q = q16v;
            sp.x.rows = seed16_rows.as<uint8_t>();
            sp.x.sq = seed16_sq.as<float>();
            sp.x.nrm = seed16_nrm.as<float>();
            sp.x.row_bytes = row_bytes16;
        }
    } else {
        sp.n_splits = vsb::exact_pick_splits(nb, n_seed_rows, sm_count);
    }
    const bool seed_scan = !seed_tc && nb <= vsb::graph_search_small_batch();
    if (seed_scan) sp.n_splits = vsb::seed_scan_blocks(n_seed_rows);
    CU(seed_part.ensure(vsb::exact_part_elems(nb, sp.n_splits, 32) * 8));
    sp.part = seed_part.as<uint64_t>();
    t_begin(PH_SEED, s);
    if (seed_tc) seed_tc = vsb::launch_exact_candidates_tc(sp, s, true);
    if (!seed_tc) {
        const uint32_t splits = sp.n_splits;
        sp = sp_native;
        sp.n_splits = splits;
        sp.part = seed_part.as<uint64_t>();
        if (seed_scan) {
            // tiny batch: one warp per 4 seed rows, one winner per CTA
            CU(cudaMemsetAsync(seed_part.p, 0xFF, vsb::exact_part_elems(nb, sp.n_splits, 32) * 8, s));
            vsb::launch_seed_scan(storage, metric, qv, sp.x, seed_part.as<uint64_t>(), s);
        } else {
            vsb::launch_exact_candidates(sp, s);
        }
    }
    t_end(s);
    CU(cudaGetLastError());
    // ---- K4 beam search (on the bf16 traversal copy when VSB_FLAG_BF16_TRAVERSAL is set) ----
    vsb::SearchParams gp;
    gp.storage = trav16 ? VSB_BF16 : storage;
    gp.metric = metric;
    gp.q = trav16 ? q16v : qv;
    gp.x = x;
    if (trav16) {
        gp.x.rows = rows16.as<uint8_t>();
        gp.x.sq = sq16.as<float>();
        gp.x.nrm = nrm16.as<float>();
        gp.x.row_bytes = row_bytes16;
    }
    const bool use8 = trav8 && rerank;  // searches only: the build keeps the bf16 traversal
    if (use8) {
        CU(q8_rows.ensure((size_t)nb * row_bytes8));
        CU(q8_sq.ensure((size_t)nb * 4));
        CU(q8_nrm.ensure((size_t)nb * 4));
        vsb::launch_convert_rows_i8s(reinterpret_cast<const float*>(qv.rows), nb, dim, row_bytes / 4, q8_rows.as<uint8_t>(),
                                     row_bytes8, q8_sq.as<float>(), q8_nrm.as<float>(), s);
        CU(cudaGetLastError());
        gp.storage = VSB_I8;
        gp.q.rows = q8_rows.as<uint8_t>();
        gp.q.sq = q8_sq.as<float>();
        gp.q.nrm = q8_nrm.as<float>();
        gp.q.row_bytes = row_bytes8;
        gp.q.n = nb;
        gp.x.rows = rows8.as<uint8_t>();
        gp.x.sq = sq8.as<float>();
        gp.x.nrm = nrm8.as<float>();
        gp.x.row_bytes = row_bytes8;
    }
    gp.graph = graph.as<uint32_t>();
    gp.graph_stride = graph_stride;
    gp.degree = degree;
    gp.n_graphed = n_graphed;
    gp.seed_lists = seed_part.as<uint64_t>();
    gp.seed_stride = sp.n_splits * 32;
    gp.n_seeds = n_seeds;
    gp.seed_slots = seed_slots.as<uint32_t>();
    gp.deny = deny_bm;
    gp.keys = keys.as<uint64_t>();
    gp.itopk = std::max(itopk_eff, k);
    gp.max_iters = max_iters;
    gp.search_width = search_width;
    gp.k = k;
    gp.out_keys = g_keys;
    gp.out_dists = g_dists;
    gp.out_counts = have_tail ? nullptr : o_counts;
    gp.self_base = self_base;
    uint32_t kr = 0;
    if (packed_out != nullptr) {
        gp.out_packed = packed_out;
        gp.out_counts = nullptr;
    }
    if (rerank) {
        // hand the best kr bf16-ranked candidates to K3 for the canonical fp32 re-rank
        // (2k for small k, k + 32 + k/4 for large k, at least k + 6; bf16 ranking errors only reorder
        // candidates near the k-th distance and are far smaller than that margin)
        // int8 traversal ranks more coarsely: rr_mult8 * k candidates go to the re-rank
        const uint32_t kv = use8 ? std::min<uint32_t>(std::max(rr_mult8 * k, k + 16), 256)
                                 : std::max(std::min(2 * k, k + 32 + k / 4), k + 6);
        kr = std::min<uint32_t>(round_up(kv, 32), 256);
        if (kr < k) return fail(VSB_EINVAL, "k=%u too large for the bf16-traversal re-rank (max 256)", k);
        CU(rr_packed.ensure((size_t)nb * kr * 8));
        gp.k = std::min(kv, kr);
        gp.out_stride = kr;
        gp.out_packed = rr_packed.as<uint64_t>();
        gp.out_counts = nullptr;
    }
    if (instrumented) {
        CU(counters.ensure(16));
        CU(cudaMemsetAsync(counters.p, 0, 16, s));
        gp.counters = counters.as<unsigned long long>();
    }
    t_begin(PH_GRAPH, s);
    vsb::launch_graph_search(gp, s);
    t_end(s);
    CU(cudaGetLastError());
    if (rerank) {
        vsb::ExactParams rp;
        rp.storage = storage;
        rp.metric = metric;
        rp.q = qv;
        rp.x = x;
        rp.keys = keys.as<uint64_t>();
        rp.part = rr_packed.as<uint64_t>();
        rp.kp = kr;
        rp.n_splits = 1;
        t_begin(PH_EXACT, s);
        vsb::launch_exact_rerank(rp, k, g_keys, g_dists, have_tail ? nullptr : o_counts, nullptr, -1, s);
        t_end(s);
        CU(cudaGetLastError());
    }
    if (instrumented) {
        unsigned long long h[2];
        CU(cudaMemcpyAsync(h, counters.p, 16, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        last_evals = h[0];
        last_parents = h[1];
        last_queries = nb;
    }
    return VSB_OK;
}

// K7: link every un-graphed tail row into the existing graph (batched HNSW-style insert:
// search with beam = expansion_add, connect to the R closest, add reverse edges).
vsb_status vsb_index::stream_insert() {
    if (n_graphed == 0 || n_graphed >= n_slots) return VSB_OK;
    CU(cudaSetDevice(device));
    ST(use_stream(stream));
    const size_t need = (size_t)capacity * graph_stride * 4;
    if (graph.bytes < need) {
        DevBuf g2;
        CU(g2.ensure(need));
        CU(cudaMemcpyAsync(g2.p, graph.p, (size_t)n_graphed * graph_stride * 4, cudaMemcpyDeviceToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        graph = std::move(g2);
    }
    const uint32_t R = degree;
    const uint32_t ef_add = opt.expansion_add ? opt.expansion_add : 128;
    const uint32_t ef = std::min<uint32_t>(round_up(std::max(ef_add, R), 32), 512);
    DevBuf cand;
    const uint32_t QB = 8192;
    while (n_graphed < n_slots) {
        const uint32_t t0 = n_graphed;
        const uint32_t nb = std::min(QB, n_slots - t0);
        vsb::RowsView qv;
        qv.rows = rows.as<uint8_t>() + (size_t)t0 * row_bytes;
        qv.sq = sq.as<float>() + t0;
        qv.nrm = nrm.as<float>() + t0;
        qv.row_bytes = row_bytes;
        qv.n = nb;
        vsb::RowsView q16;
        if (trav16) {
            q16.rows = rows16.as<uint8_t>() + (size_t)t0 * row_bytes16;
            q16.sq = sq16.as<float>() + t0;
            q16.nrm = nrm16.as<float>() + t0;
            q16.row_bytes = row_bytes16;
            q16.n = nb;
        }
        CU(cand.ensure((size_t)nb * R * 8));
        ST(graph_block(qv, nb, R, ef, nullptr, nullptr, nullptr, cand.as<uint64_t>(), trav16 ? &q16 : nullptr, stream));
        vsb::launch_stream_link(cand.as<uint64_t>(), nb, R, t0, R, graph.as<uint32_t>(), graph_stride, stream);
        CU(cudaGetLastError());
        n_graphed = t0 + nb;  // later batches may link to these rows (stream order)
        churn_since_refine += nb;
    }
    CU(cudaStreamSynchronize(stream));
    // Streamed links are a little worse than built ones and tombstoned rows keep occupying beam slots: once
    // 10 % of the graph has churned, one refinement pass (K4 kNN lists of every row -> K6) restores the
    // quality of a fresh build and drops the tombstoned rows from every list (~0.8 s per million rows).
    if (!in_build && refine_passes > 0 && churn_since_refine * 10 >= n_graphed && n_graphed >= min_graph_size) {
        ST(refine_graph());
        ST(sample_seeds(n_graphed));
        churn_since_refine = 0;
    }
    return VSB_OK;
}

vsb_status vsb_index::search_dev(const float* d_q, uint64_t nq, uint32_t k, uint64_t* d_keys, float* d_dists,
                                 uint32_t* d_counts, cudaStream_t s, bool exact, const uint32_t* d_allow,
                                 uint64_t allow_bits) {
    if (nq == 0) return VSB_OK;
    if (k == 0) return fail(VSB_EINVAL, "k must be > 0");
    if (d_q == nullptr || d_keys == nullptr || d_dists == nullptr) return fail(VSB_EINVAL, "null buffer");
    CU(cudaSetDevice(device));
    ST(use_stream(s));
    const uint64_t QCHUNK = 65536;
    for (uint64_t q0 = 0; q0 < nq; q0 += QCHUNK) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(QCHUNK, nq - q0);
        uint64_t* o_keys = d_keys + q0 * k;
        float* o_dists = d_dists + q0 * k;
        uint32_t* o_counts = d_counts ? d_counts + q0 : nullptr;
        if (n_slots == 0) {
            vsb::launch_fill_empty(o_keys, o_dists, o_counts, nb, k, s);
            CU(cudaGetLastError());
            continue;
        }
        CU(q_rows.ensure((size_t)nb * row_bytes));
        CU(q_sq.ensure((size_t)nb * 4));
        CU(q_nrm.ensure((size_t)nb * 4));
        t_begin(PH_CONVERT, s);
        vsb::launch_convert_rows(storage, d_q + q0 * dim, nb, dim, q_rows.as<uint8_t>(), row_bytes, q_sq.as<float>(),
                                 q_nrm.as<float>(), s);
        t_end(s);
        CU(cudaGetLastError());
        vsb::RowsView qv;
        qv.rows = q_rows.as<uint8_t>();
        qv.sq = q_sq.as<float>();
        qv.nrm = q_nrm.as<float>();
        qv.row_bytes = row_bytes;
        qv.n = nb;
        const vsb::RowsView x = corpus_view();
        const uint32_t* deny_bm = any_tombstone ? deny.as<uint32_t>() : nullptr;
        const bool use_graph = !exact && d_allow == nullptr && n_graphed > 0 && k <= 1024;
        const uint32_t tail_lo = use_graph ? n_graphed : 0, tail_hi = n_slots;
        const bool have_tail = tail_hi > tail_lo;
        uint64_t* g_keys = o_keys;
        float* g_dists = o_dists;
        uint64_t* t_keys = o_keys;
        float* t_dists = o_dists;
        if (use_graph && have_tail) {
            CU(tmp_keys.ensure((size_t)2 * nb * k * 8));
            CU(tmp_dists.ensure((size_t)2 * nb * k * 4));
            g_keys = tmp_keys.as<uint64_t>();
            g_dists = tmp_dists.as<float>();
            t_keys = g_keys + (size_t)nb * k;
            t_dists = g_dists + (size_t)nb * k;
        }
        if (use_graph) {
            ST(graph_block(qv, nb, k, itopk, g_keys, g_dists, have_tail ? nullptr : o_counts, nullptr, nullptr, s));
        }
        if (have_tail) {
            t_begin(PH_EXACT, s);
            vsb_status est = exact_block(qv, x, tail_lo, tail_hi, deny_bm, keys.as<uint64_t>(), d_allow, allow_bits, k,
                                         t_keys, t_dists, use_graph ? nullptr : o_counts, nullptr, -1, s,
                                         /*approx_ok=*/use_graph);  // the ANN tail needs no certificate (no host sync)
            t_end(s);
            ST(est);
        }
        if (use_graph && have_tail) {
            t_begin(PH_MERGE, s);
            vsb::launch_merge_topk(g_keys, g_dists, 2, nb, k, o_keys, o_dists, o_counts, s);
            t_end(s);
            CU(cudaGetLastError());
        }
    }
    return VSB_OK;
}

vsb_status vsb_index::search_host(const float* queries, uint64_t nq, uint32_t k, uint64_t* keys_out,
                                  float* dists_out, uint32_t* counts_out, bool exact, const uint32_t* allow_bitmap,
                                  uint64_t allow_bits) {
    if (nq == 0) return VSB_OK;
    if (k == 0) return fail(VSB_EINVAL, "k must be > 0");
    if (queries == nullptr || keys_out == nullptr || dists_out == nullptr) return fail(VSB_EINVAL, "null buffer");
    CU(cudaSetDevice(device));
    ST(use_stream(stream));
    DevBuf& d_in = q_in;
    const size_t in_bytes = (size_t)nq * dim * 4;
    const size_t keys_bytes = (size_t)nq * k * 8, dists_bytes = (size_t)nq * k * 4, counts_bytes = (size_t)nq * 4;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    CU(d_in.ensure(al(in_bytes) + al(keys_bytes) + al(dists_bytes) + al(counts_bytes)));
    uint8_t* base = d_in.as<uint8_t>();
    float* dq = reinterpret_cast<float*>(base);
    uint64_t* dk = reinterpret_cast<uint64_t*>(base + al(in_bytes));
    float* dd = reinterpret_cast<float*>(base + al(in_bytes) + al(keys_bytes));
    uint32_t* dc = reinterpret_cast<uint32_t*>(base + al(in_bytes) + al(keys_bytes) + al(dists_bytes));
    CU(cudaMemcpyAsync(dq, queries, in_bytes, cudaMemcpyHostToDevice, stream));
    const uint32_t* d_allow = nullptr;
    if (allow_bitmap != nullptr) {
        const size_t words = (size_t)((allow_bits + 31) / 32);
        CU(allow.ensure(std::max<size_t>(words * 4, 16)));
        CU(cudaMemcpyAsync(allow.p, allow_bitmap, words * 4, cudaMemcpyHostToDevice, stream));
        d_allow = allow.as<uint32_t>();
    }
    ST(search_dev(dq, nq, k, dk, dd, dc, stream, exact || allow_bitmap != nullptr, d_allow, allow_bits));
    CU(cudaMemcpyAsync(keys_out, dk, keys_bytes, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(dists_out, dd, dists_bytes, cudaMemcpyDeviceToHost, stream));
    if (counts_out) CU(cudaMemcpyAsync(counts_out, dc, counts_bytes, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return VSB_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

vsb_status vsb_create(const vsb_options* o, vsb_index** out) {
    if (o == nullptr || out == nullptr) return fail(VSB_EINVAL, "null options/out");
    *out = nullptr;
    if (o->dimensions == 0) return fail(VSB_EINVAL, "dimensions must be > 0");
    if (o->storage < VSB_F32 || o->storage > VSB_B1) return fail(VSB_EINVAL, "unknown storage scalar %d", o->storage);
    if (o->metric < VSB_L2SQ || o->metric > VSB_HAMMING) return fail(VSB_EINVAL, "unknown metric %d", o->metric);
    int metric = o->metric;
    // usearch.rs:450-464: B1 always uses Hamming; Hamming without B1 is rejected (usearch.rs:480-485)
    if (o->storage == VSB_B1) metric = VSB_HAMMING;
    else if (metric == VSB_HAMMING) return fail(VSB_EINVAL, "Binary space type requires B1 quantization.");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VSB_ECUDA, "no usable CUDA device (%s): vsb200 has no CPU fallback", cudaGetErrorString(e));
    int dev = o->device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(VSB_EINVAL, "device %d out of range (%d devices)", dev, ndev);
    CU(cudaSetDevice(dev));
    vsb_index* ix = new (std::nothrow) vsb_index();
    if (!ix) return fail(VSB_EOOM, "host allocation failed");
    ix->opt = *o;
    ix->dim = o->dimensions;
    ix->metric = metric;
    ix->storage = o->storage;
    ix->device = dev;
    ix->row_bytes = storage_row_bytes(o->storage, o->dimensions);
    const uint32_t M = o->connectivity ? o->connectivity : 16;
    const uint32_t ef_add = o->expansion_add ? o->expansion_add : 128;
    const uint32_t ef_search = o->expansion_search ? o->expansion_search : 64;
    ix->degree = std::min<uint32_t>(std::max<uint32_t>(2 * M, 8), 64);
    ix->graph_stride = round_up(ix->degree, 32);
    ix->k_init = std::min<uint32_t>(std::max<uint32_t>(ef_add / 2, ix->degree), 128);
    ix->itopk = std::min<uint32_t>(round_up(ef_search, 32), 1024);
    ix->trav16 = (o->flags & VSB_FLAG_BF16_TRAVERSAL) != 0 && o->storage == VSB_F32;
    ix->row_bytes16 = storage_row_bytes(VSB_BF16, o->dimensions);
    ix->trav8 = (o->flags & VSB_FLAG_I8_TRAVERSAL) != 0 && o->storage == VSB_F32 && o->metric == VSB_COS;
    if (ix->trav8) ix->trav16 = true;  // the bf16 copy feeds the build and the seed tiles
    ix->row_bytes8 = storage_row_bytes(VSB_I8, o->dimensions);
    if (const char* e = getenv("VSB_I8_RERANK_MULT")) ix->rr_mult8 = std::max<uint32_t>(2, (uint32_t)strtoul(e, nullptr, 10));
    if (const char* e = getenv("VSB_DISABLE_TC")) ix->tc_enabled = !(e[0] == '1');
    if (const char* e = getenv("VSB_DISABLE_CERT")) ix->cert_enabled = !(e[0] == '1');
    if (const char* e = getenv("VSB_DISABLE_REACH_FIX")) ix->reach_fix = !(e[0] == '1');
    if (const char* e = getenv("VSB_CERT_KP")) ix->cert_kp = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("VSB_CERT_KP16")) ix->cert_kp16 = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("VSB_TC_MIN_ROWS")) ix->tc_min_rows = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("VSB_ALLPAIRS_MAX")) ix->allpairs_max = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("VSB_ALLPAIRS_PREFIX")) ix->allpairs_prefix = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("VSB_REFINE_PASSES")) ix->refine_passes = (uint32_t)strtoul(e, nullptr, 10);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) ix->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ix;
        return fail(VSB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return VSB_OK;
}

void vsb_destroy(vsb_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();
    ix->t_resolve();
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

vsb_status vsb_reserve(vsb_index* ix, uint64_t capacity) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->reserve(capacity);
}

uint64_t vsb_capacity(const vsb_index* ix) { return ix ? ix->capacity_atomic.load() : 0; }
uint64_t vsb_size(const vsb_index* ix) { return ix ? ix->live_atomic.load() : 0; }

vsb_status vsb_add(vsb_index* ix, const uint64_t* keys, const float* rows, uint64_t n) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->add(keys, rows, n);
}

vsb_status vsb_remove(vsb_index* ix, const uint64_t* keys, uint64_t n, uint64_t* n_removed) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    if (n_removed) *n_removed = 0;
    if (n == 0) return VSB_OK;
    if (!keys) return fail(VSB_EINVAL, "null keys");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->remove(keys, n, n_removed);
}

int vsb_contains(const vsb_index* ix, uint64_t key) {
    if (!ix) return 0;
    std::lock_guard<std::mutex> g(const_cast<vsb_index*>(ix)->mu);
    return ix->key2slot.count(key) ? 1 : 0;
}

vsb_status vsb_build(vsb_index* ix) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->build();
}

vsb_status vsb_insert_pending(vsb_index* ix) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->stream_insert();
}

vsb_status vsb_export_graph(vsb_index* ix, uint32_t* rows_out, uint64_t* keys_out, uint64_t* n_graphed,
                            uint32_t* stride) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    if (n_graphed) *n_graphed = ix->n_graphed;
    if (stride) *stride = ix->graph_stride;
    CU(cudaSetDevice(ix->device));
    ST(ix->use_stream(ix->stream));
    if (rows_out && ix->n_graphed)
        CU(cudaMemcpyAsync(rows_out, ix->graph.p, (size_t)ix->n_graphed * ix->graph_stride * 4, cudaMemcpyDeviceToHost,
                           ix->stream));
    if (keys_out && ix->n_graphed)
        CU(cudaMemcpyAsync(keys_out, ix->keys.p, (size_t)ix->n_graphed * 8, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    return VSB_OK;
}

vsb_status vsb_set_search_params(vsb_index* ix, const vsb_search_params* p) {
    if (!ix || !p) return fail(VSB_EINVAL, "null argument");
    std::lock_guard<std::mutex> g(ix->mu);
    if (p->expansion_search) ix->itopk = std::min<uint32_t>(round_up(p->expansion_search, 32), 1024);
    if (p->max_iterations) ix->max_iters = p->max_iterations >= 1000000u ? 0 : p->max_iterations;  // >= 1e6: back to auto
    if (p->n_seeds) ix->n_seeds = std::min<uint32_t>(p->n_seeds, 32);
    if (p->min_graph_size) ix->min_graph_size = p->min_graph_size;
    if (p->search_width) ix->search_width = std::min<uint32_t>(p->search_width, 4);
    if (p->stream_threshold) ix->stream_threshold = p->stream_threshold == 0xFFFFFFFFu ? 0 : p->stream_threshold;
    return VSB_OK;
}

vsb_status vsb_set_instrumented(vsb_index* ix, int on) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    ix->instrumented = on != 0;
    return VSB_OK;
}

vsb_status vsb_set_kernel_timing(vsb_index* ix, int on) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    cudaSetDevice(ix->device);
    ix->t_resolve();
    ix->timing = on != 0;
    if (on) {
        for (int i = 0; i < vsb_index::PH_COUNT; ++i) ix->phase_ns[i] = ix->phase_launches[i] = 0;
    }
    return VSB_OK;
}

vsb_status vsb_get_stats(vsb_index* ix, vsb_stats* out) {
    if (!ix || !out) return fail(VSB_EINVAL, "null argument");
    std::lock_guard<std::mutex> g(ix->mu);
    out->kernel_launches = vsb::g_kernel_launches.load();
    out->distance_evals = ix->last_evals;
    out->parent_expansions = ix->last_parents;
    out->queries = ix->last_queries;
    out->n_slots = ix->n_slots;
    out->n_graphed = ix->n_graphed;
    out->graph_degree = ix->degree;
    out->row_bytes = ix->row_bytes;
    out->n_seed_rows = ix->n_seed_rows;
    out->hbm_bytes = ix->hbm_bytes();
    cudaSetDevice(ix->device);
    ix->t_resolve();
    out->convert_ns = ix->phase_ns[vsb_index::PH_CONVERT];
    out->seed_ns = ix->phase_ns[vsb_index::PH_SEED];
    out->graph_search_ns = ix->phase_ns[vsb_index::PH_GRAPH];
    out->exact_ns = ix->phase_ns[vsb_index::PH_EXACT];
    out->merge_ns = ix->phase_ns[vsb_index::PH_MERGE];
    out->convert_launches = ix->phase_launches[vsb_index::PH_CONVERT];
    out->seed_launches = ix->phase_launches[vsb_index::PH_SEED];
    out->graph_search_launches = ix->phase_launches[vsb_index::PH_GRAPH];
    out->exact_launches = ix->phase_launches[vsb_index::PH_EXACT];
    out->merge_launches = ix->phase_launches[vsb_index::PH_MERGE];
    out->tc_launches = vsb::g_tc_launches.load();
    out->exact_certified = ix->cert_ok;
    out->exact_fallback = ix->cert_fallback;
    out->exact_scanned = ix->cert_scanned;
    out->extra_seeds = ix->n_extra_seeds;
    return VSB_OK;
}

vsb_status vsb_search(vsb_index* ix, const float* queries, uint64_t q, uint32_t k, uint64_t* keys, float* distances,
                      uint32_t* counts) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->search_host(queries, q, k, keys, distances, counts, false, nullptr, 0);
}

vsb_status vsb_search_exact(vsb_index* ix, const float* queries, uint64_t q, uint32_t k, uint64_t* keys,
                            float* distances, uint32_t* counts) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->search_host(queries, q, k, keys, distances, counts, true, nullptr, 0);
}

vsb_status vsb_search_filtered(vsb_index* ix, const float* queries, uint64_t q, uint32_t k,
                               const uint32_t* allow_bitmap, uint64_t bitmap_bits, uint64_t* keys, float* distances,
                               uint32_t* counts) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    if (!allow_bitmap) return fail(VSB_EINVAL, "null bitmap");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->search_host(queries, q, k, keys, distances, counts, true, allow_bitmap, bitmap_bits);
}

vsb_status vsb_search_dev(vsb_index* ix, const float* d_queries, uint64_t q, uint32_t k, uint64_t* d_keys,
                          float* d_distances, uint32_t* d_counts, void* stream, int exact) {
    if (!ix) return fail(VSB_EINVAL, "null index");
    std::lock_guard<std::mutex> g(ix->mu);
    return ix->search_dev(d_queries, q, k, d_keys, d_distances, d_counts, static_cast<cudaStream_t>(stream),
                          exact != 0, nullptr, 0);
}

vsb_status vsb_merge_topk_dev(const uint64_t* d_keys, const float* d_distances, uint32_t parts, uint64_t q, uint32_t k,
                              uint64_t* d_out_keys, float* d_out_distances, uint32_t* d_out_counts, int device,
                              void* stream) {
    if (!d_keys || !d_distances || !d_out_keys || !d_out_distances) return fail(VSB_EINVAL, "null buffer");
    if (parts == 0 || k == 0 || (uint64_t)parts * k > 2048) return fail(VSB_EINVAL, "parts*k must be in [1, 2048]");
    if (device >= 0) CU(cudaSetDevice(device));
    vsb::launch_merge_topk(d_keys, d_distances, parts, q, k, d_out_keys, d_out_distances, d_out_counts,
                           static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return VSB_OK;
}

vsb_status vsb_merge_topk_strided_dev(const uint64_t* d_keys, const float* d_distances, uint32_t parts,
                                      uint64_t key_part_stride, uint64_t dist_part_stride, uint64_t q, uint32_t k,
                                      uint64_t* d_out_keys, float* d_out_distances, uint32_t* d_out_counts, int device,
                                      void* stream) {
    if (!d_keys || !d_distances || !d_out_keys || !d_out_distances) return fail(VSB_EINVAL, "null buffer");
    if (parts == 0 || k == 0 || (uint64_t)parts * k > 2048) return fail(VSB_EINVAL, "parts*k must be in [1, 2048]");
    if (device >= 0) CU(cudaSetDevice(device));
    vsb::launch_merge_topk(d_keys, d_distances, parts, q, k, d_out_keys, d_out_distances, d_out_counts,
                           static_cast<cudaStream_t>(stream), key_part_stride, dist_part_stride);
    CU(cudaGetLastError());
    return VSB_OK;
}

void vsb_f32_to_b1x8(const float* v, uint64_t n, uint8_t* out) {
    const uint64_t nb = (n + 7) / 8;
    for (uint64_t j = 0; j < nb; ++j) {
        uint8_t b = 0;
        for (uint64_t i = 0; i < 8 && 8 * j + i < n; ++i)
            if (v[8 * j + i] > 0.0f) b |= (uint8_t)(1u << i);
        out[j] = b;
    }
}

const char* vsb_last_error(void) { return g_last_error.c_str(); }
const char* vsb_version(void) { return "vsb200-0.1.0"; }

}  // extern "C"

// ---- N3: snapshot ------------------------------------------------------------------------------
namespace {
struct SnapHeader {
    char magic[8];  // "VSB200S1"
    vsb_options opt;
    uint64_t n_slots, n_graphed, capacity;
    uint32_t row_bytes, graph_stride, degree, reserved;
};

bool write_dev(FILE* f, const void* dptr, size_t bytes, std::vector<uint8_t>& stage, cudaStream_t s) {
    const size_t CH = stage.size();
    for (size_t off = 0; off < bytes; off += CH) {
        const size_t nb = std::min(CH, bytes - off);
        if (cudaMemcpyAsync(stage.data(), static_cast<const uint8_t*>(dptr) + off, nb, cudaMemcpyDeviceToHost, s) != cudaSuccess)
            return false;
        if (cudaStreamSynchronize(s) != cudaSuccess) return false;
        if (fwrite(stage.data(), 1, nb, f) != nb) return false;
    }
    return true;
}
bool read_dev(FILE* f, void* dptr, size_t bytes, std::vector<uint8_t>& stage, cudaStream_t s) {
    const size_t CH = stage.size();
    for (size_t off = 0; off < bytes; off += CH) {
        const size_t nb = std::min(CH, bytes - off);
        if (fread(stage.data(), 1, nb, f) != nb) return false;
        if (cudaMemcpyAsync(static_cast<uint8_t*>(dptr) + off, stage.data(), nb, cudaMemcpyHostToDevice, s) != cudaSuccess)
            return false;
        if (cudaStreamSynchronize(s) != cudaSuccess) return false;
    }
    return true;
}
}  // namespace

extern "C" vsb_status vsb_save(vsb_index* ix, const char* path) {
    if (!ix || !path) return fail(VSB_EINVAL, "null argument");
    std::lock_guard<std::mutex> g(ix->mu);
    CU(cudaSetDevice(ix->device));
    ST(ix->use_stream(ix->stream));
    FILE* f = fopen(path, "wb");
    if (!f) return fail(VSB_EINVAL, "cannot open %s for writing", path);
    SnapHeader h{};
    memcpy(h.magic, "VSB200S1", 8);
    h.opt = ix->opt;
    h.n_slots = ix->n_slots;
    h.n_graphed = ix->n_graphed;
    h.capacity = ix->capacity;
    h.row_bytes = ix->row_bytes;
    h.graph_stride = ix->graph_stride;
    h.degree = ix->degree;
    std::vector<uint8_t> stage((size_t)64 << 20);
    const size_t n = ix->n_slots;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    ok = ok && fwrite(ix->h_deny.data(), 4, (n + 31) / 32, f) == (n + 31) / 32;
    ok = ok && write_dev(f, ix->keys.p, n * 8, stage, ix->stream);
    ok = ok && write_dev(f, ix->rows.p, n * ix->row_bytes, stage, ix->stream);
    ok = ok && write_dev(f, ix->sq.p, n * 4, stage, ix->stream);
    ok = ok && write_dev(f, ix->nrm.p, n * 4, stage, ix->stream);
    ok = ok && write_dev(f, ix->graph.p, (size_t)ix->n_graphed * ix->graph_stride * 4, stage, ix->stream);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(VSB_ECUDA, "short write or copy failure while saving %s", path);
    return VSB_OK;
}

extern "C" vsb_status vsb_load(const char* path, int32_t device, vsb_index** out) {
    if (!path || !out) return fail(VSB_EINVAL, "null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(VSB_EINVAL, "cannot open %s", path);
    SnapHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "VSB200S1", 8) != 0) {
        fclose(f);
        return fail(VSB_EINVAL, "%s is not a vsb200 snapshot", path);
    }
    h.opt.device = device;
    vsb_index* ix = nullptr;
    vsb_status st = vsb_create(&h.opt, &ix);
    if (st != VSB_OK) {
        fclose(f);
        return st;
    }
    auto bail = [&](vsb_status code, const char* what) {
        fclose(f);
        vsb_destroy(ix);
        return fail(code, "%s while loading %s", what, path);
    };
    if (h.row_bytes != ix->row_bytes || h.graph_stride != ix->graph_stride) return bail(VSB_EINVAL, "layout mismatch");
    const size_t n = h.n_slots;  // nobody else holds this handle yet: no locking needed
    if (ix->reserve(std::max<uint64_t>(h.capacity, std::max<uint64_t>(n, 1))) != VSB_OK) return bail(VSB_EOOM, "reserve failed");
    std::vector<uint8_t> stage((size_t)64 << 20);
    std::vector<uint64_t> h_keys(n);
    bool ok = fread(ix->h_deny.data(), 4, (n + 31) / 32, f) == (n + 31) / 32;
    ok = ok && fread(h_keys.data(), 8, n, f) == n;
    if (!ok) return bail(VSB_EINVAL, "truncated file");
    if (cudaMemcpy(ix->keys.p, h_keys.data(), n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ix->deny.p, ix->h_deny.data(), ((n + 31) / 32) * 4, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(VSB_ECUDA, "upload failed");
    ok = read_dev(f, ix->rows.p, n * ix->row_bytes, stage, ix->stream);
    ok = ok && read_dev(f, ix->sq.p, n * 4, stage, ix->stream);
    ok = ok && read_dev(f, ix->nrm.p, n * 4, stage, ix->stream);
    if (ok && h.n_graphed) {
        if (ix->graph.ensure((size_t)std::max<uint64_t>(ix->capacity, h.n_graphed) * ix->graph_stride * 4) != cudaSuccess)
            return bail(VSB_EOOM, "graph allocation failed");
        ok = read_dev(f, ix->graph.p, (size_t)h.n_graphed * ix->graph_stride * 4, stage, ix->stream);
    }
    if (!ok) return bail(VSB_EINVAL, "truncated file or copy failure");
    fclose(f);
    ix->n_slots = (uint32_t)n;
    ix->n_graphed = (uint32_t)h.n_graphed;
    uint64_t live = 0;
    for (size_t i = 0; i < n; ++i) {
        if (ix->h_deny[i >> 5] >> (i & 31) & 1u) {
            ix->any_tombstone = true;
            continue;
        }
        ix->key2slot.emplace(h_keys[i], (uint32_t)i);
        ++live;
    }
    ix->live = live;
    ix->live_atomic.store(live);
    if (ix->trav16 && n) {  // the bf16 traversal copy is derived data: regenerate instead of storing it
        vsb::launch_convert_rows(VSB_BF16, ix->rows.as<float>(), (uint32_t)n, ix->row_bytes / 4, ix->rows16.as<uint8_t>(),
                                 ix->row_bytes16, ix->sq16.as<float>(), ix->nrm16.as<float>(), ix->stream);
    }
    if (ix->trav8 && n) {
        vsb::launch_convert_rows_i8s(ix->rows.as<float>(), (uint32_t)n, ix->dim, ix->row_bytes / 4, ix->rows8.as<uint8_t>(),
                                     ix->row_bytes8, ix->sq8.as<float>(), ix->nrm8.as<float>(), ix->stream);
    }
    if (ix->n_graphed && ix->sample_seeds(ix->n_graphed) != VSB_OK) {
        vsb_destroy(ix);
        return VSB_ECUDA;
    }
    cudaStreamSynchronize(ix->stream);
    *out = ix;
    return VSB_OK;
}
