// batcher.cu — N2: the micro-batcher the reference's actor loop lacks.
//
// The reference serves ONE query per message (vs_index/actor.rs:38-56; `recv` at vs_index/mod.rs:30-45
// hands each `VsIndexSearch::Ann` to a worker thread that calls usearch once).  A GPU only pays off when
// concurrent callers are coalesced, so this is the piece a Rust shim would put in that loop, written here
// in C++ behind the same C ABI: any number of threads call vsb_batcher_search() with one query each; a
// single dispatcher thread drains the queue — up to `max_batch` requests or `max_wait_us` after the first —
// into ONE vsb_search() call and fans the rows back out (the oneshot replies of actor.rs:137-147).
// Searches keep priority over modifications exactly like the biased select: modifications go straight to
// vsb_add/vsb_remove on the caller's thread and only contend on the index mutex between batches.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vsb200.h"

namespace {
struct Request {
    const float* query;
    uint32_t k;
    uint64_t* keys;
    float* dists;
    uint32_t* count;
    vsb_status status = VSB_OK;
    bool done = false;
    std::condition_variable cv;
};
}  // namespace

struct vsb_batcher {
    vsb_index* index = nullptr;
    uint32_t dim = 0, max_batch = 1024, max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work;
    std::deque<Request*> queue;
    bool stop = false;
    std::thread worker;
    uint64_t n_queries = 0, n_batches = 0;
    std::string last_error;

    void run() {
        std::vector<Request*> batch;
        std::vector<float> q;
        std::vector<uint64_t> keys;
        std::vector<float> dists;
        std::vector<uint32_t> counts;
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            cv_work.wait(lk, [&] { return stop || !queue.empty(); });
            if (stop && queue.empty()) return;
            // first request is in: give followers max_wait_us to arrive (or until the batch is full)
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us);
            while (!stop && queue.size() < max_batch && cv_work.wait_until(lk, deadline) != std::cv_status::timeout) {
            }
            batch.clear();
            const uint32_t k = queue.front()->k;  // one k per batch; other k's wait for the next round
            for (auto it = queue.begin(); it != queue.end() && batch.size() < max_batch;) {
                if ((*it)->k == k) {
                    batch.push_back(*it);
                    it = queue.erase(it);
                } else {
                    ++it;
                }
            }
            lk.unlock();
            const size_t n = batch.size();
            q.resize(n * dim);
            keys.resize(n * k);
            dists.resize(n * k);
            counts.resize(n);
            for (size_t i = 0; i < n; ++i) std::memcpy(&q[i * dim], batch[i]->query, sizeof(float) * dim);
            const vsb_status st = vsb_search(index, q.data(), n, k, keys.data(), dists.data(), counts.data());
            for (size_t i = 0; i < n; ++i) {
                if (st == VSB_OK) {
                    std::memcpy(batch[i]->keys, &keys[i * k], sizeof(uint64_t) * k);
                    std::memcpy(batch[i]->dists, &dists[i * k], sizeof(float) * k);
                    if (batch[i]->count) *batch[i]->count = counts[i];
                }
            }
            lk.lock();
            n_queries += n;
            n_batches += 1;
            if (st != VSB_OK) last_error = vsb_last_error();
            for (Request* r : batch) {
                r->status = st;
                r->done = true;
                r->cv.notify_one();
            }
        }
    }
};

extern "C" {

vsb_status vsb_batcher_create(vsb_index* index, uint32_t dimensions, uint32_t max_batch, uint32_t max_wait_us,
                              vsb_batcher** out) {
    if (!index || !out || dimensions == 0) return VSB_EINVAL;
    vsb_batcher* b = new vsb_batcher();
    b->index = index;
    b->dim = dimensions;
    if (max_batch) b->max_batch = max_batch;
    b->max_wait_us = max_wait_us;
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return VSB_OK;
}

void vsb_batcher_destroy(vsb_batcher* b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
}

vsb_status vsb_batcher_search(vsb_batcher* b, const float* query, uint32_t k, uint64_t* keys, float* distances,
                              uint32_t* count) {
    if (!b || !query || !keys || !distances || k == 0) return VSB_EINVAL;
    Request r;
    r.query = query;
    r.k = k;
    r.keys = keys;
    r.dists = distances;
    r.count = count;
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stop) return VSB_EINVAL;
    b->queue.push_back(&r);
    b->cv_work.notify_all();
    r.cv.wait(lk, [&] { return r.done; });
    return r.status;
}

vsb_status vsb_batcher_stats(vsb_batcher* b, uint64_t* n_queries, uint64_t* n_batches) {
    if (!b) return VSB_EINVAL;
    std::lock_guard<std::mutex> g(b->mu);
    if (n_queries) *n_queries = b->n_queries;
    if (n_batches) *n_batches = b->n_batches;
    return VSB_OK;
}

}  // extern "C"
