/* abi_client.c — a plain C consumer of include/vsb200.h (no Python, no C++): replays the reference's
 * `add_or_replace_size_ann` scenario (crates/vector-store/src/vs_index/usearch.rs:1298-1458) through the C ABI.
 * Exit code 0 = scenario passed, 77 = no CUDA device (vsb_create returned VSB_ECUDA), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "vsb200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        vsb_status s_ = (call);                                                       \
        if (s_ != VSB_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)s_, vsb_last_error());      \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static int nearest(vsb_index* ix, const float* q, uint64_t* key, float* dist) {
    uint32_t count = 0;
    vsb_status s = vsb_search(ix, q, 1, 1, key, dist, &count);
    if (s != VSB_OK) {
        fprintf(stderr, "vsb_search -> %d: %s\n", (int)s, vsb_last_error());
        return -1;
    }
    return (int)count;
}

/* the actor's partition state through the C ABI (vsb_set_*, usearch.rs:626-895) */
static int partition_set_scenario(const vsb_options* opt) {
    vsb_set* set = NULL;
    CHECK(vsb_set_create(opt, 8, &set));
    const uint64_t local = ((uint64_t)5 << 48) | 1, global = (uint64_t)0x8002 << 48;
    const uint64_t keys[3] = {1, 2, 3};
    const float rows[9] = {1.f, 1.f, 1.f, 2.f, -2.f, 2.f, 3.f, 3.f, 3.f};
    const float q[3] = {2.2f, -2.2f, 2.2f};
    uint64_t key = 0, n = 0;
    float dist = 0.f;
    uint32_t count = 7;
    CHECK(vsb_set_search(set, local, q, 1, 1, NULL, 0, &key, &dist, &count)); /* unknown partition: empty answer */
    if (count != 0 || vsb_set_partitions(set) != 0) return 20;
    CHECK(vsb_set_add(set, local, keys, rows, 3, &n));
    if (n != 3 || vsb_set_count(set, 5) != 3 || vsb_capacity(vsb_set_index(set, local)) != 1000) return 21;
    CHECK(vsb_set_add(set, global, keys, rows, 2, &n));
    if (n != 2 || vsb_set_count(set, 0x8002) != 2 || vsb_capacity(vsb_set_index(set, global)) != 1000000) return 22;
    CHECK(vsb_set_add(set, local, keys, rows, 1, &n)); /* duplicate key: swallowed, nothing added */
    if (n != 0 || vsb_set_count(set, 5) != 3) return 23;
    CHECK(vsb_set_search(set, local, q, 1, 1, NULL, 0, &key, &dist, &count));
    if (count != 1 || key != 2 || fabsf(dist - 0.12f) > 1e-5f) return 24;
    const uint32_t allow = 1u << 3; /* only row id 3 is admissible */
    CHECK(vsb_set_search(set, local, q, 1, 1, &allow, 32, &key, &dist, &count));
    if (count != 1 || key != 3) return 25;
    CHECK(vsb_set_remove(set, local, keys + 1, 1, &n));
    if (n != 1 || vsb_set_count(set, 5) != 2) return 26;
    CHECK(vsb_set_remove(set, ((uint64_t)5 << 48) | 9, keys, 1, &n)); /* unknown partition: not an error */
    if (n != 0) return 27;
    CHECK(vsb_set_remove_partition(set, local));
    CHECK(vsb_set_search(set, local, q, 1, 1, NULL, 0, &key, &dist, &count));
    if (count != 0 || vsb_set_partitions(set) != 1) return 28;
    vsb_set_destroy(set);
    return 0;
}

/* usage: abi_client [n_devices]  — n_devices >= 2 runs the scenario on ONE handle sharded over devices 0..n-1 */
int main(int argc, char** argv) {
    vsb_options opt;
    memset(&opt, 0, sizeof opt);
    opt.dimensions = 3;
    opt.metric = VSB_L2SQ;
    opt.storage = VSB_F32;
    opt.device = -1;
    if (argc > 1) {
        int n = 0;
        sscanf(argv[1], "%d", &n);
        if (n >= 2 && n <= 8) {
            opt.n_devices = n;
            for (int i = 0; i < n; ++i) opt.device_ids[i] = i;
        }
    }
    vsb_index* ix = NULL;
    vsb_status st = vsb_create(&opt, &ix);
    if (st == VSB_ECUDA) {
        printf("no CUDA device: %s\n", vsb_last_error());
        return 77;
    }
    if (st != VSB_OK) return 1;
    printf("engine %s\n", vsb_version());
    CHECK(vsb_reserve(ix, 1000));
    const uint64_t keys[3] = {1, 2, 3};
    const float rows[9] = {1.f, 1.f, 1.f, 2.f, -2.f, 2.f, 3.f, 3.f, 3.f};
    CHECK(vsb_add(ix, keys, rows, 3));
    if (vsb_size(ix) != 3 || vsb_capacity(ix) < 1000) return 2;
    if (vsb_add(ix, keys, rows, 1) != VSB_EDUPKEY) return 3; /* multi = false */
    const float q[3] = {2.2f, -2.2f, 2.2f};
    uint64_t key = 0;
    float dist = 0.f;
    if (nearest(ix, q, &key, &dist) != 1 || key != 2) return 4;
    if (fabsf(dist - 0.12f) > 1e-5f) return 5; /* 3 * 0.2^2 */
    uint64_t removed = 0;
    const uint64_t k3 = 3;
    CHECK(vsb_remove(ix, &k3, 1, &removed));
    if (removed != 1 || vsb_size(ix) != 2) return 6;
    const float row3b[3] = {2.1f, -2.1f, 2.1f};
    CHECK(vsb_add(ix, &k3, row3b, 1));
    if (nearest(ix, q, &key, &dist) != 1 || key != 3) return 7;
    CHECK(vsb_remove(ix, &k3, 1, &removed));
    if (nearest(ix, q, &key, &dist) != 1 || key != 2 || vsb_size(ix) != 2) return 8;
    const float bad[2] = {1.f, 2.f};
    (void)bad;
    vsb_destroy(ix);
    if (opt.n_devices == 0) {
        const int rc = partition_set_scenario(&opt);
        if (rc != 0) return rc;
    }
    printf("abi_client: scenario passed%s\n", opt.n_devices ? " (sharded handle)" : "");
    return 0;
}
