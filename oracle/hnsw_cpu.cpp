// oracle/hnsw_cpu.cpp — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libvsb200).
//
// "USearch-equivalent CPU HNSW (restatement, not USearch 2.22.0)": the reference's index arithmetic
// lives in the crates.io `usearch` 2.22.0 C++ library (Cargo.toml:93), which cannot be built
// offline, so this file restates the published HNSW algorithm with USearch's parameter meaning
// as used by the reference (vs_index/usearch.rs:74-82; defaults lib.rs:394-437):
//   connectivity M on upper levels, 2M on level 0; level = floor(-ln(U) / ln(M));
//   expansion_add = beam width while inserting; expansion_search: ef = max(expansion_search, k);
//   neighbour selection = the HNSW heuristic (keep a candidate only if it is closer to the new
//   node than to every neighbour already kept); unique keys; results ascending by distance.
// Concurrency mirrors the reference's Insert family (usearch.rs:515-624): many inserts in flight,
// one per thread (OpenMP threads = the VECTOR_STORE_THREADS equivalent), per-node spin locks.
// It is the timed CPU baseline of bench.py (`cpu_baseline.kind = "port"`) and the recall yardstick
// of the ANN tests.  PARITY UNPINNED against real USearch: no reference test asserts a recall.
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <queue>
#include <random>
#include <vector>

namespace {

enum { M_L2SQ = 0, M_COS = 1, M_IP = 2 };

#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))

CLONES float dot_f32(const float* a, const float* b, int d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; ++i) s += a[i] * b[i];
    return s;
}
CLONES float l2_f32(const float* a, const float* b, int d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; ++i) {
        float t = a[i] - b[i];
        s += t * t;
    }
    return s;
}

struct SpinLock {
    std::atomic_flag f = ATOMIC_FLAG_INIT;
    void lock() { while (f.test_and_set(std::memory_order_acquire)) {} }
    void unlock() { f.clear(std::memory_order_release); }
};

typedef std::pair<float, uint32_t> DI;

struct Hnsw {
    int dim, metric, M, M0, ef_add, ef_search;
    double mult;
    size_t cap, n = 0;
    std::vector<float> data, inv_norm;
    std::vector<uint64_t> keys;
    std::vector<int> level;
    std::vector<uint32_t> links0;              // cap * (M0 + 1): [count, ids...]
    std::vector<std::vector<uint32_t>> linksU; // per node: level * (M + 1)
    std::vector<SpinLock> locks;
    std::mutex global;
    int max_level = -1;
    uint32_t entry = 0;
    std::mt19937_64 rng;
    std::vector<std::vector<uint32_t>> visited;  // per thread: epoch stamps
    std::vector<uint32_t> visit_epoch;

    Hnsw(int dim_, int metric_, int M_, int efa, int efs, size_t cap_, uint64_t seed)
        : dim(dim_), metric(metric_), M(M_), M0(2 * M_), ef_add(efa), ef_search(efs), mult(1.0 / std::log((double)M_)),
          cap(cap_), data(cap_ * (size_t)dim_), inv_norm(cap_, 1.f), keys(cap_), level(cap_, 0),
          links0(cap_ * (size_t)(2 * M_ + 1), 0), linksU(cap_), locks(cap_), rng(seed) {
        int t = omp_get_max_threads();
        visited.assign(t, std::vector<uint32_t>(cap_, 0));
        visit_epoch.assign(t, 0);
    }

    inline float dist(const float* q, float q_inv, uint32_t id) const {
        const float* x = &data[(size_t)id * dim];
        if (metric == M_L2SQ) return l2_f32(q, x, dim);
        float d = dot_f32(q, x, dim);
        if (metric == M_IP) return 1.f - d;
        float c = 1.f - d * q_inv * inv_norm[id];
        return c < 0.f ? 0.f : c;
    }
    uint32_t* nbrs(uint32_t id, int l) {
        return l == 0 ? &links0[(size_t)id * (M0 + 1)] : &linksU[id][(size_t)(l - 1) * (M + 1)];
    }

    // beam search on one level; returns up to ef closest as a max-heap
    std::priority_queue<DI> search_layer(const float* q, float q_inv, uint32_t ep, float ep_d, int ef, int l) {
        int tid = omp_get_thread_num();
        std::vector<uint32_t>& vis = visited[tid];
        uint32_t epoch = ++visit_epoch[tid];
        if (epoch == 0) { std::fill(vis.begin(), vis.end(), 0); epoch = ++visit_epoch[tid]; }
        std::priority_queue<DI> top;                                       // worst on top
        std::priority_queue<DI, std::vector<DI>, std::greater<DI>> cand;   // best on top
        top.emplace(ep_d, ep);
        cand.emplace(ep_d, ep);
        vis[ep] = epoch;
        uint32_t buf[257];
        while (!cand.empty()) {
            DI c = cand.top();
            if (c.first > top.top().first && (int)top.size() >= ef) break;
            cand.pop();
            locks[c.second].lock();
            uint32_t* nb = nbrs(c.second, l);
            uint32_t cnt = nb[0];
            std::memcpy(buf, nb + 1, cnt * 4);
            locks[c.second].unlock();
            // software prefetch of the neighbours' vectors, as hnswlib / usearch do: the loop is DRAM-latency bound
            for (uint32_t i = 0; i < cnt; ++i) {
                const char* p = reinterpret_cast<const char*>(&data[(size_t)buf[i] * dim]);
                __builtin_prefetch(p);
                __builtin_prefetch(p + 64);
            }
            for (uint32_t i = 0; i < cnt; ++i) {
                uint32_t v = buf[i];
                if (i + 1 < cnt) {
                    const char* p = reinterpret_cast<const char*>(&data[(size_t)buf[i + 1] * dim]);
                    for (int off = 128; off < dim * 4; off += 64) __builtin_prefetch(p + off);
                }
                if (vis[v] == epoch) continue;
                vis[v] = epoch;
                float d = dist(q, q_inv, v);
                if ((int)top.size() < ef || d < top.top().first) {
                    cand.emplace(d, v);
                    top.emplace(d, v);
                    if ((int)top.size() > ef) top.pop();
                }
            }
        }
        return top;
    }

    // HNSW heuristic: input ascending by distance to the base point
    void select(std::vector<DI>& sorted, int m, std::vector<uint32_t>& out) {
        out.clear();
        for (const DI& c : sorted) {
            if ((int)out.size() >= m) break;
            bool good = true;
            const float* cv = &data[(size_t)c.second * dim];
            float c_inv = inv_norm[c.second];
            for (uint32_t s : out) {
                if (dist(cv, c_inv, s) < c.first) { good = false; break; }
            }
            if (good) out.push_back(c.second);
        }
    }

    void add_one(uint32_t id, int lvl) {
        const float* q = &data[(size_t)id * dim];
        const float q_inv = inv_norm[id];
        std::unique_lock<std::mutex> gl(global);
        int ml = max_level;
        if (ml < 0) {  // first node
            max_level = lvl;
            entry = id;
            return;
        }
        if (lvl <= ml) gl.unlock();
        uint32_t cur = entry;
        float cur_d = dist(q, q_inv, cur);
        uint32_t buf[257];
        for (int l = ml; l > lvl; --l) {
            bool changed = true;
            while (changed) {
                changed = false;
                locks[cur].lock();
                uint32_t* nb = nbrs(cur, l);
                uint32_t cnt = nb[0];
                std::memcpy(buf, nb + 1, cnt * 4);
                locks[cur].unlock();
                for (uint32_t i = 0; i < cnt; ++i) {
                    float d = dist(q, q_inv, buf[i]);
                    if (d < cur_d) { cur_d = d; cur = buf[i]; changed = true; }
                }
            }
        }
        std::vector<DI> sorted;
        std::vector<uint32_t> sel, sel2;
        for (int l = std::min(lvl, ml); l >= 0; --l) {
            auto top = search_layer(q, q_inv, cur, cur_d, ef_add, l);
            sorted.clear();
            while (!top.empty()) { sorted.push_back(top.top()); top.pop(); }
            std::reverse(sorted.begin(), sorted.end());
            select(sorted, M, sel);
            const int mmax = l == 0 ? M0 : M;
            locks[id].lock();
            uint32_t* mine = nbrs(id, l);
            mine[0] = (uint32_t)sel.size();
            for (size_t i = 0; i < sel.size(); ++i) mine[1 + i] = sel[i];
            locks[id].unlock();
            for (uint32_t v : sel) {
                locks[v].lock();
                uint32_t* nb = nbrs(v, l);
                uint32_t cnt = nb[0];
                if ((int)cnt < mmax) {
                    nb[1 + cnt] = id;
                    nb[0] = cnt + 1;
                } else {
                    const float* vv = &data[(size_t)v * dim];
                    const float v_inv = inv_norm[v];
                    std::vector<DI> cands;
                    cands.reserve(cnt + 1);
                    cands.emplace_back(dist(vv, v_inv, id), id);
                    for (uint32_t i = 0; i < cnt; ++i) cands.emplace_back(dist(vv, v_inv, nb[1 + i]), nb[1 + i]);
                    std::sort(cands.begin(), cands.end());
                    select(cands, mmax, sel2);
                    nb[0] = (uint32_t)sel2.size();
                    for (size_t i = 0; i < sel2.size(); ++i) nb[1 + i] = sel2[i];
                }
                locks[v].unlock();
            }
            if (!sorted.empty()) { cur = sorted[0].second; cur_d = sorted[0].first; }
        }
        if (lvl > ml) {
            max_level = lvl;
            entry = id;
        }
    }

    void search_one(const float* q, int k, uint64_t* out_keys, float* out_d) {
        for (int i = 0; i < k; ++i) { out_keys[i] = UINT64_MAX; out_d[i] = INFINITY; }
        if (max_level < 0) return;
        float q_inv = 1.f;
        if (metric == M_COS) {
            float s = dot_f32(q, q, dim);
            q_inv = s > 0.f ? 1.f / std::sqrt(s) : 0.f;
        }
        uint32_t cur = entry;
        float cur_d = dist(q, q_inv, cur);
        for (int l = max_level; l > 0; --l) {
            bool changed = true;
            while (changed) {
                changed = false;
                uint32_t* nb = nbrs(cur, l);
                for (uint32_t i = 0; i < nb[0]; ++i) {
                    float d = dist(q, q_inv, nb[1 + i]);
                    if (d < cur_d) { cur_d = d; cur = nb[1 + i]; changed = true; }
                }
            }
        }
        auto top = search_layer(q, q_inv, cur, cur_d, std::max(ef_search, k), 0);
        while ((int)top.size() > k) top.pop();
        int c = (int)top.size();
        for (int i = c - 1; i >= 0; --i) { out_keys[i] = keys[top.top().second]; out_d[i] = top.top().first; top.pop(); }
    }
};

float round_bf16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

extern "C" {

void* hnsw_create(int dim, int metric, int M, int ef_add, int ef_search, uint64_t capacity, uint64_t seed) {
    if (M < 2) M = 2;
    if (M > 128) M = 128;
    return new Hnsw(dim, metric, M, ef_add, ef_search, (size_t)capacity, seed);
}
void hnsw_free(void* h) { delete static_cast<Hnsw*>(h); }
void hnsw_set_ef(void* h, int ef) { static_cast<Hnsw*>(h)->ef_search = ef; }
uint64_t hnsw_size(void* h) { return static_cast<Hnsw*>(h)->n; }

// storage: 0 = f32, 2 = bf16 (values rounded to bf16, math in f32)
int hnsw_add_batch(void* hp, const uint64_t* keys, const float* rows, uint64_t n, int storage, int threads) {
    Hnsw* h = static_cast<Hnsw*>(hp);
    if (h->n + n > h->cap) return 1;
    if (threads > 0) omp_set_num_threads(threads);
    if ((int)h->visited.size() < omp_get_max_threads()) {
        h->visited.assign(omp_get_max_threads(), std::vector<uint32_t>(h->cap, 0));
        h->visit_epoch.assign(omp_get_max_threads(), 0);
    }
    const size_t base = h->n;
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (uint64_t i = 0; i < n; ++i) {
        const size_t id = base + i;
        float* dst = &h->data[id * h->dim];
        const float* src = rows + i * h->dim;
        double s = 0;
        for (int d = 0; d < h->dim; ++d) {
            float v = storage == 2 ? round_bf16(src[d]) : src[d];
            dst[d] = v;
            s += (double)v * v;
        }
        h->inv_norm[id] = s > 0 ? (float)(1.0 / std::sqrt(s)) : 0.f;
        h->keys[id] = keys[i];
        double u = U(h->rng);
        if (u < 1e-300) u = 1e-300;
        int lvl = (int)std::floor(-std::log(u) * h->mult);
        h->level[id] = lvl;
        if (lvl > 0) h->linksU[id].assign((size_t)lvl * (h->M + 1), 0);
    }
    h->n += n;
    uint64_t start = 0;
    if (base == 0 && n > 0) {
        h->add_one((uint32_t)base, h->level[base]);
        start = 1;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = (int64_t)start; i < (int64_t)n; ++i) h->add_one((uint32_t)(base + i), h->level[base + i]);
    return 0;
}

void hnsw_search_batch(void* hp, const float* queries, uint64_t nq, int k, uint64_t* out_keys, float* out_d,
                       int storage, int threads) {
    Hnsw* h = static_cast<Hnsw*>(hp);
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
    {
        std::vector<float> qb(h->dim);
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < (int64_t)nq; ++i) {
            const float* q = queries + (size_t)i * h->dim;
            if (storage == 2) {
                for (int d = 0; d < h->dim; ++d) qb[d] = round_bf16(q[d]);
                q = qb.data();
            }
            h->search_one(q, k, out_keys + (size_t)i * k, out_d + (size_t)i * k);
        }
    }
}

// single query, single thread: the batch-1 latency path
void hnsw_search_one(void* hp, const float* query, int k, uint64_t* out_keys, float* out_d) {
    static_cast<Hnsw*>(hp)->search_one(query, k, out_keys, out_d);
}

int hnsw_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
