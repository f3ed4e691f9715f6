// index_set.cu — A13: the actor's partition state as a C++ object behind the C ABI.
//
// The reference's index actor keeps a map PartitionId -> usearch index plus a live counter per IndexId
// (vs_index/usearch.rs:626-895): partitions are created lazily by the first AddVector, grow their capacity by
// 1 000 000 (global index) or 1 000 (local index) slots whenever fewer than `free_threshold` slots are left
// (usearch.rs:442-443, 655-665), an Ann for a partition that does not exist answers with an empty result
// (usearch.rs:787-806), Count is served from the counter (usearch.rs:868-878), RemovePartition drops the index
// (usearch.rs:888-893), and add / remove errors are swallowed after logging (usearch.rs:1020-1050).
// `vsb_set` is that state for vsb200 indexes, so a host shim (integration/vs_index/gpu.rs keeps the same map on the
// Rust side) or the Python mirror does not have to re-implement it.  PartitionId layout: IndexId in the top 16 bits
// (bit 63 = global index), partition number in the low 48 (table/partition_id.rs:11-44).
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>

#include "index_impl.h"

using vsbi::fail;

namespace {
constexpr uint64_t kReserveIncrementGlobal = 1000000;  // usearch.rs:442
constexpr uint64_t kReserveIncrementLocal = 1000;      // usearch.rs:443

struct Partition {
    vsb_index* idx = nullptr;
    uint64_t increment = kReserveIncrementLocal;
    ~Partition() {
        if (idx) vsb_destroy(idx);
    }
};
}  // namespace

struct vsb_set {
    vsb_options opt{};
    uint64_t free_threshold = 64;
    mutable std::shared_mutex mu;  // the MAP only; the indexes synchronise themselves
    std::map<uint64_t, std::shared_ptr<Partition>> partitions;
    std::map<uint16_t, std::shared_ptr<std::atomic<uint64_t>>> sizes;  // IndexState::size per IndexId

    std::shared_ptr<Partition> find(uint64_t pid) const {
        std::shared_lock<std::shared_mutex> g(mu);
        auto it = partitions.find(pid);
        return it == partitions.end() ? nullptr : it->second;
    }
    std::shared_ptr<std::atomic<uint64_t>> size_of(uint16_t index_id, bool create) {
        {
            std::shared_lock<std::shared_mutex> g(mu);
            auto it = sizes.find(index_id);
            if (it != sizes.end()) return it->second;
        }
        if (!create) return nullptr;
        std::unique_lock<std::shared_mutex> g(mu);
        auto& s = sizes[index_id];
        if (!s) s = std::make_shared<std::atomic<uint64_t>>(0);
        return s;
    }
};

extern "C" {

vsb_status vsb_set_create(const vsb_options* options, uint32_t free_threshold, vsb_set** out) {
    if (!options || !out) return fail(VSB_EINVAL, "null argument");
    // validate the options once, on a throw-away empty index (no rows are allocated before the first reserve)
    vsb_index* probe = nullptr;
    const vsb_status st = vsb_create(options, &probe);
    if (st != VSB_OK) return st;
    vsb_destroy(probe);
    vsb_set* s = new vsb_set();
    s->opt = *options;
    if (free_threshold) s->free_threshold = free_threshold;
    *out = s;
    return VSB_OK;
}

void vsb_set_destroy(vsb_set* set) { delete set; }

/* AddVector x n for one partition.  Creates the partition on first use, reserves ahead of the insertion like
 * PartitionState::needs_more_capacity, inserts with per-row status (a duplicate key fails alone).  Returns VSB_OK
 * unless the partition could not be created or grown; *n_added counts the rows that went in. */
vsb_status vsb_set_add(vsb_set* set, uint64_t partition_id, const uint64_t* keys, const float* rows, uint64_t n,
                       uint64_t* n_added) {
    if (n_added) *n_added = 0;
    if (!set) return fail(VSB_EINVAL, "null set");
    if (n == 0) return VSB_OK;
    std::shared_ptr<Partition> p = set->find(partition_id);
    if (!p) {
        std::unique_lock<std::shared_mutex> g(set->mu);
        auto& slot = set->partitions[partition_id];
        if (!slot) {
            auto np = std::make_shared<Partition>();
            const vsb_status st = vsb_create(&set->opt, &np->idx);
            if (st != VSB_OK) {
                set->partitions.erase(partition_id);
                return st;
            }
            np->increment = (partition_id >> 63) ? kReserveIncrementGlobal : kReserveIncrementLocal;
            slot = np;
        }
        p = slot;
    }
    const uint64_t cap = vsb_capacity(p->idx), size = vsb_size(p->idx);
    if (cap - size < set->free_threshold + n) {
        const uint64_t want = cap + std::max<uint64_t>(p->increment, n + set->free_threshold);
        const vsb_status st = vsb_reserve(p->idx, want);
        if (st != VSB_OK) return st;  // the reference logs and goes on; the add below would fail with VSB_EFULL
    }
    uint64_t added = 0;
    const vsb_status st = vsb_add_each(p->idx, keys, rows, n, nullptr, &added);
    if (st != VSB_OK) return st;
    set->size_of((uint16_t)(partition_id >> 48), true)->fetch_add(added);
    if (n_added) *n_added = added;
    return VSB_OK;
}

/* RemoveVector x n: unknown partitions and unknown keys are not errors (usearch.rs:880-886, 1037-1049). */
vsb_status vsb_set_remove(vsb_set* set, uint64_t partition_id, const uint64_t* keys, uint64_t n, uint64_t* n_removed) {
    if (n_removed) *n_removed = 0;
    if (!set) return fail(VSB_EINVAL, "null set");
    std::shared_ptr<Partition> p = set->find(partition_id);
    if (!p || n == 0) return VSB_OK;
    uint64_t removed = 0;
    const vsb_status st = vsb_remove(p->idx, keys, n, &removed);
    if (st != VSB_OK) return st;
    if (auto s = set->size_of((uint16_t)(partition_id >> 48), false)) s->fetch_sub(removed);
    if (n_removed) *n_removed = removed;
    return VSB_OK;
}

/* RemovePartition: the index dies with its last user (searches in flight keep their reference). */
vsb_status vsb_set_remove_partition(vsb_set* set, uint64_t partition_id) {
    if (!set) return fail(VSB_EINVAL, "null set");
    std::shared_ptr<Partition> victim;
    {
        std::unique_lock<std::shared_mutex> g(set->mu);
        auto it = set->partitions.find(partition_id);
        if (it == set->partitions.end()) return VSB_OK;
        victim = it->second;
        set->partitions.erase(it);
    }
    // the reference does not touch IndexState::size here either (usearch.rs:888-893)
    return VSB_OK;
}

/* Ann / FilteredAnn x q for one partition.  allow_bitmap == NULL is a plain Ann — also what a FilteredAnn whose
 * restrictions were consumed by the partition key is downgraded to (usearch.rs:844-862).  A partition that does not
 * exist answers every query with zero hits. */
vsb_status vsb_set_search(vsb_set* set, uint64_t partition_id, const float* queries, uint64_t q, uint32_t k,
                          const uint32_t* allow_bitmap, uint64_t bitmap_bits, uint64_t* keys, float* distances,
                          uint32_t* counts) {
    if (!set) return fail(VSB_EINVAL, "null set");
    if (q == 0) return VSB_OK;
    if (k == 0) return fail(VSB_EINVAL, "k must be > 0");
    if (!queries || !keys || !distances) return fail(VSB_EINVAL, "null buffer");
    std::shared_ptr<Partition> p = set->find(partition_id);
    if (!p) {
        for (uint64_t i = 0; i < q * k; ++i) {
            keys[i] = 0xFFFFFFFFFFFFFFFFull;
            distances[i] = __builtin_inff();
        }
        if (counts)
            for (uint64_t i = 0; i < q; ++i) counts[i] = 0;
        return VSB_OK;
    }
    if (allow_bitmap) return vsb_search_filtered(p->idx, queries, q, k, allow_bitmap, bitmap_bits, keys, distances, counts);
    return vsb_search(p->idx, queries, q, k, keys, distances, counts);
}

/* Count: live vectors over all partitions of one IndexId (= partition_id >> 48); 0 for an unknown index. */
uint64_t vsb_set_count(const vsb_set* cset, uint16_t index_id) {
    if (!cset) return 0;
    auto s = const_cast<vsb_set*>(cset)->size_of(index_id, false);
    return s ? s->load() : 0;
}

uint64_t vsb_set_partitions(const vsb_set* set) {
    if (!set) return 0;
    std::shared_lock<std::shared_mutex> g(set->mu);
    return set->partitions.size();
}

/* The partition's index handle (NULL if it does not exist) for calls this wrapper does not forward: vsb_build,
 * vsb_set_search_params, vsb_get_stats, vsb_save.  Owned by the set; do not destroy; invalid after
 * vsb_set_remove_partition / vsb_set_destroy. */
vsb_index* vsb_set_index(vsb_set* set, uint64_t partition_id) {
    if (!set) return nullptr;
    std::shared_ptr<Partition> p = set->find(partition_id);
    return p ? p->idx : nullptr;
}

}  // extern "C"
