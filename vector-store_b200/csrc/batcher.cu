// batcher.cu — N2: the micro-batcher the reference's actor loop lacks.
//
// The reference serves ONE query per message (vs_index/actor.rs:38-56; `recv` at vs_index/mod.rs:30-45
// hands each `VsIndexSearch::Ann` to a worker thread that calls usearch once).  A GPU only pays off when
// concurrent callers are coalesced, so this is the piece a Rust shim would put in that loop, written here
// in C++ behind the same C ABI: any number of threads call vsb_batcher_search() with one query each; a
// single dispatcher thread drains the queue — up to `max_batch` requests or `max_wait_us` after the first —
// into ONE vsb_search() call and fans the rows back out (the oneshot replies of actor.rs:137-147).
// Searches keep priority over modifications exactly like the biased select (vs_index/mod.rs:39-42): a pending
// search batch is always dispatched before staged adds are flushed.
// Single-vector adds (VsIndexModify::AddVector, one vector per message, monitor_items.rs:255-353) are coalesced the
// same way: vsb_batcher_add copies the row into a staging block and returns; the dispatcher applies a block with
// ONE vsb_add_each (one H2D + one stream sync for the block instead of one per row).
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
    // staged single-row adds
    std::vector<uint64_t> add_keys;
    std::vector<float> add_rows;
    std::chrono::steady_clock::time_point add_first;
    uint64_t adds_staged = 0, adds_applied = 0, adds_ok = 0, adds_failed = 0;
    std::condition_variable cv_flushed;

    // mu held on entry and exit; released around the index call
    void flush_adds(std::unique_lock<std::mutex>& lk) {
        std::vector<uint64_t> k;
        std::vector<float> r;
        k.swap(add_keys);
        r.swap(add_rows);
        const uint64_t n = k.size();
        if (n == 0) return;
        lk.unlock();
        uint64_t ok = 0;
        std::vector<int32_t> st((size_t)n, 0);
        vsb_status rc = vsb_add_each(index, k.data(), r.data(), n, st.data(), &ok);
        if (rc == VSB_EFULL) {
            // the reference's actor reserves ahead of the stream (usearch.rs:655-665); do the same and retry once
            const uint64_t cap = vsb_capacity(index);
            if (vsb_reserve(index, cap + std::max<uint64_t>(n, cap / 4) + 1024) == VSB_OK)
                rc = vsb_add_each(index, k.data(), r.data(), n, st.data(), &ok);
        }
        lk.lock();
        adds_applied += n;
        adds_ok += ok;
        adds_failed += n - ok;
        if (rc != VSB_OK) last_error = vsb_last_error();
        cv_flushed.notify_all();
    }

    void run() {
        std::vector<Request*> batch;
        std::vector<float> q;
        std::vector<uint64_t> keys;
        std::vector<float> dists;
        std::vector<uint32_t> counts;
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            if (add_keys.empty()) {
                cv_work.wait(lk, [&] { return stop || !queue.empty() || !add_keys.empty(); });
            } else {
                // staged adds wait for company until max_wait_us after the first of them, searches permitting
                const auto add_deadline = add_first + std::chrono::microseconds(max_wait_us);
                cv_work.wait_until(lk, add_deadline, [&] { return stop || !queue.empty() || add_keys.size() >= max_batch; });
                if (queue.empty() && (stop || add_keys.size() >= max_batch || std::chrono::steady_clock::now() >= add_deadline)) {
                    flush_adds(lk);
                    continue;
                }
            }
            if (stop && queue.empty()) {
                flush_adds(lk);
                return;
            }
            if (queue.empty()) continue;
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

vsb_status vsb_batcher_add(vsb_batcher* b, uint64_t key, const float* row) {
    if (!b || !row) return VSB_EINVAL;
    std::lock_guard<std::mutex> g(b->mu);
    if (b->stop) return VSB_EINVAL;
    if (b->add_keys.empty()) b->add_first = std::chrono::steady_clock::now();
    b->add_keys.push_back(key);
    b->add_rows.insert(b->add_rows.end(), row, row + b->dim);
    b->adds_staged += 1;
    if (b->add_keys.size() == 1 || b->add_keys.size() >= b->max_batch) b->cv_work.notify_all();
    return VSB_OK;
}

vsb_status vsb_batcher_flush(vsb_batcher* b, uint64_t* n_added, uint64_t* n_failed) {
    if (!b) return VSB_EINVAL;
    std::unique_lock<std::mutex> lk(b->mu);
    const uint64_t target = b->adds_staged;
    // make the dispatcher flush now instead of waiting for company
    b->add_first = std::chrono::steady_clock::now() - std::chrono::hours(1);
    b->cv_work.notify_all();
    b->cv_flushed.wait(lk, [&] { return b->adds_applied >= target || b->stop; });
    if (n_added) *n_added = b->adds_ok;
    if (n_failed) *n_failed = b->adds_failed;
    return VSB_OK;
}

vsb_status vsb_batcher_stats(vsb_batcher* b, uint64_t* n_queries, uint64_t* n_batches) {
    if (!b) return VSB_EINVAL;
    std::lock_guard<std::mutex> g(b->mu);
    if (n_queries) *n_queries = b->n_queries;
    if (n_batches) *n_batches = b->n_batches;
    return VSB_OK;
}

}  // extern "C"
