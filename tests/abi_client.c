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

int main(void) {
    vsb_options opt;
    memset(&opt, 0, sizeof opt);
    opt.dimensions = 3;
    opt.metric = VSB_L2SQ;
    opt.storage = VSB_F32;
    opt.device = -1;
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
    printf("abi_client: scenario passed\n");
    return 0;
}
