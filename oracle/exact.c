/*
 * oracle/exact.c — TEST INFRASTRUCTURE ONLY (never linked into libvsb200, never on a product path).
 *
 * CPU restatement of the distance arithmetic and exact k-NN of the index path of
 * scylladb/vector-store.  The arithmetic itself lives in the third-party crate `usearch` 2.22.0
 * (Cargo.toml:93, Cargo.lock:5920-5927; C++/SimSIMD, not vendored, not buildable offline), so this
 * file restates its published semantics and is anchored on the reference's own call sites and
 * golden vectors:
 *   - metric mapping Cos / L2sq (squared, no sqrt) / IP / Hamming:  vs_index/usearch.rs:450-501
 *   - L2sq = sum (a-b)^2            pinned by tests/integration/vs_index.rs:1795-1798 (0,1,9)
 *   - Cos  = 1 - a.b/(|a||b|), 0 if both norms 0, 1 if one is 0, clamped to [0,2]
 *            (range required by distance.rs:66-69)
 *   - IP   = 1 - a.b                (distance.rs:77)
 *   - Hamming(b1x8) = popcount(a xor b);  packing per vs_index/usearch.rs:1179-1205
 *   - storage casts on add AND query: f16/bf16 round-to-nearest-even,
 *     i8 = round(clamp(x,-1,1)*127)  pinned by tests/integration/quantization.rs:35-39
 * Parity status: exact distances PINNED by goldens G1-G7,G9 (tests/test_oracle_golden.py);
 * i8 IP scaling (1 - a.b/127^2) and the half-away rounding of exact .5 cases are UNPINNED.
 *
 * Summation order is the "canonical order" the CUDA kernels mirror bit for bit
 * (vector-store_b200/csrc/common.cuh): 16-byte chunks, chunk c -> lane c%32, one fmaf accumulator
 * per lane, xor-butterfly 16,8,4,2,1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { M_L2SQ = 0, M_COS = 1, M_IP = 2, M_HAMMING = 3 };
enum { S_F32 = 0, S_F16 = 1, S_BF16 = 2, S_I8 = 3, S_B1 = 4 };

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* f32 -> bf16 bits, round-to-nearest-even (NaN quieted) */
static uint16_t f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x0040u);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7FFFu + lsb;
    return (uint16_t)(u >> 16);
}
static float bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }

uint32_t vso_row_bytes(int storage, uint32_t dim) {
    uint64_t bits = storage == S_F32 ? (uint64_t)dim * 32 : (storage == S_F16 || storage == S_BF16) ? (uint64_t)dim * 16
                    : storage == S_I8 ? (uint64_t)dim * 8 : dim;
    return (uint32_t)(((bits + 127) / 128) * 16);
}

static int elems_per_chunk(int storage) {
    return storage == S_F32 ? 4 : (storage == S_F16 || storage == S_BF16) ? 8 : storage == S_I8 ? 16 : 128;
}

/* Casts one f32 row into the padded storage row (zero tail). */
void vso_convert_row(int storage, const float* v, uint32_t dim, uint8_t* out) {
    uint32_t rb = vso_row_bytes(storage, dim);
    memset(out, 0, rb);
    for (uint32_t i = 0; i < dim; ++i) {
        switch (storage) {
            case S_F32: memcpy(out + 4 * i, &v[i], 4); break;
            case S_BF16: { uint16_t h = f32_to_bf16(v[i]); memcpy(out + 2 * i, &h, 2); } break;
            case S_F16: { _Float16 h = (_Float16)v[i]; memcpy(out + 2 * i, &h, 2); } break;
            case S_I8: {
                float c = v[i] < -1.0f ? -1.0f : (v[i] > 1.0f ? 1.0f : v[i]);
                int q = (int)roundf(c * 127.0f);
                out[i] = (uint8_t)(int8_t)q;
            } break;
            default:
                if (v[i] > 0.0f) out[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
    }
}

static float elem_f(int storage, const uint8_t* row, uint32_t e) {
    switch (storage) {
        case S_F32: { float f; memcpy(&f, row + 4 * e, 4); return f; }
        case S_BF16: { uint16_t h; memcpy(&h, row + 2 * e, 2); return bf16_to_f32(h); }
        default: { _Float16 h; memcpy(&h, row + 2 * e, 2); return (float)h; }
    }
}

static float butterfly(float* s) {
    for (int m = 16; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ m];
        memcpy(s, t, sizeof t);
    }
    return s[0];
}

/* canonical sum over chunks of op(a,b): mode 0 = a*b, mode 1 = (a-b)^2 */
static float canon_f(int storage, const uint8_t* a, const uint8_t* b, uint32_t row_bytes, int mode) {
    const int E = elems_per_chunk(storage);
    const uint32_t n_chunks = row_bytes / 16;
    float acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0f;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const int lane = (int)(c % 32);
        for (int e = 0; e < E; ++e) {
            float x = elem_f(storage, a, c * E + e), y = elem_f(storage, b, c * E + e);
            if (mode == 0) acc[lane] = fmaf(x, y, acc[lane]);
            else { float d = x - y; acc[lane] = fmaf(d, d, acc[lane]); }
        }
    }
    return butterfly(acc);
}

static int canon_i8(const uint8_t* a, const uint8_t* b, uint32_t row_bytes, int mode) {
    int s = 0;
    for (uint32_t i = 0; i < row_bytes; ++i) {
        int x = (int8_t)a[i], y = (int8_t)b[i];
        s += mode == 0 ? x * y : (x - y) * (x - y);
    }
    return s;
}

static int popcount_xor(const uint8_t* a, const uint8_t* b, uint32_t row_bytes) {
    int s = 0;
    for (uint32_t i = 0; i < row_bytes; ++i) s += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return s;
}

/* canonical sum of squares of a stored row (popcount for b1) */
float vso_sqnorm(int storage, const uint8_t* row, uint32_t row_bytes) {
    if (storage == S_B1) {
        int s = 0;
        for (uint32_t i = 0; i < row_bytes; ++i) s += __builtin_popcount(row[i]);
        return (float)s;
    }
    if (storage == S_I8) return (float)canon_i8(row, row, row_bytes, 0);
    return canon_f(storage, row, row, row_bytes, 0);
}

/* distance between two STORED rows; qn/xn = sqrtf(sqnorm) (cosine only) */
float vso_distance(int storage, int metric, const uint8_t* q, const uint8_t* x, uint32_t row_bytes, float qn, float xn) {
    if (storage == S_B1) return (float)popcount_xor(q, x, row_bytes);
    float dot;
    if (metric == M_L2SQ) {
        return storage == S_I8 ? (float)canon_i8(q, x, row_bytes, 1) : canon_f(storage, q, x, row_bytes, 1);
    }
    dot = storage == S_I8 ? (float)canon_i8(q, x, row_bytes, 0) : canon_f(storage, q, x, row_bytes, 0);
    if (metric == M_IP) {
        if (storage == S_I8) dot = dot / 16129.0f;
        return 1.0f - dot;
    }
    if (qn == 0.0f && xn == 0.0f) return 0.0f;
    if (qn == 0.0f || xn == 0.0f) return 1.0f;
    float d = 1.0f - dot / (qn * xn);
    if (d < 0.0f) d = 0.0f;
    if (d > 2.0f) d = 2.0f;
    return d;
}

typedef struct { float d; uint64_t key; uint32_t idx; } hit_t;
static int hit_cmp(const void* a, const void* b) {
    const hit_t* x = (const hit_t*)a; const hit_t* y = (const hit_t*)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return x->key < y->key ? -1 : (x->key > y->key ? 1 : 0);
}

/*
 * Exact top-k of f32 queries against f32 corpus rows, both cast to `storage` first.
 * alive[i]==0 rows are skipped (nullable).  Results ascending by (distance, key), padded with
 * key=UINT64_MAX / +inf.  out_idx (nullable) receives corpus row indices (UINT32_MAX padded).
 */
void vso_exact_topk(int storage, int metric, uint32_t dim, const float* corpus, const uint64_t* keys,
                    const uint8_t* alive, uint64_t n, const float* queries, uint64_t nq, uint32_t k,
                    uint64_t* out_keys, float* out_dists, uint32_t* out_counts, uint32_t* out_idx) {
    if (storage == S_B1) metric = M_HAMMING;
    const uint32_t rb = vso_row_bytes(storage, dim);
    uint8_t* X = (uint8_t*)malloc((size_t)(n ? n : 1) * rb);
    float* xn = (float*)malloc(sizeof(float) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        vso_convert_row(storage, corpus + (size_t)i * dim, dim, X + (size_t)i * rb);
        xn[i] = sqrtf(vso_sqnorm(storage, X + (size_t)i * rb, rb));
    }
#pragma omp parallel
    {
        uint8_t* Q = (uint8_t*)malloc(rb);
        hit_t* hits = (hit_t*)malloc(sizeof(hit_t) * (n ? n : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < (int64_t)nq; ++qi) {
            vso_convert_row(storage, queries + (size_t)qi * dim, dim, Q);
            const float qn = sqrtf(vso_sqnorm(storage, Q, rb));
            uint64_t m = 0;
            for (uint64_t i = 0; i < n; ++i) {
                if (alive && !alive[i]) continue;
                hits[m].d = vso_distance(storage, metric, Q, X + (size_t)i * rb, rb, qn, xn[i]);
                hits[m].key = keys ? keys[i] : i;
                hits[m].idx = (uint32_t)i;
                ++m;
            }
            qsort(hits, m, sizeof(hit_t), hit_cmp);
            uint32_t c = (uint32_t)(m < k ? m : k);
            for (uint32_t j = 0; j < k; ++j) {
                out_keys[(size_t)qi * k + j] = j < c ? hits[j].key : UINT64_MAX;
                out_dists[(size_t)qi * k + j] = j < c ? hits[j].d : INFINITY;
                if (out_idx) out_idx[(size_t)qi * k + j] = j < c ? hits[j].idx : UINT32_MAX;
            }
            if (out_counts) out_counts[qi] = c;
        }
        free(Q);
        free(hits);
    }
    free(X);
    free(xn);
}

/* all-pairs distances (for tiny cases): out[nq][n] */
void vso_distance_matrix(int storage, int metric, uint32_t dim, const float* corpus, uint64_t n,
                         const float* queries, uint64_t nq, float* out) {
    if (storage == S_B1) metric = M_HAMMING;
    const uint32_t rb = vso_row_bytes(storage, dim);
    uint8_t* X = (uint8_t*)malloc((size_t)(n ? n : 1) * rb);
    uint8_t* Q = (uint8_t*)malloc(rb);
    float* xn = (float*)malloc(sizeof(float) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) {
        vso_convert_row(storage, corpus + (size_t)i * dim, dim, X + (size_t)i * rb);
        xn[i] = sqrtf(vso_sqnorm(storage, X + (size_t)i * rb, rb));
    }
    for (uint64_t qi = 0; qi < nq; ++qi) {
        vso_convert_row(storage, queries + (size_t)qi * dim, dim, Q);
        const float qn = sqrtf(vso_sqnorm(storage, Q, rb));
        for (uint64_t i = 0; i < n; ++i)
            out[qi * n + i] = vso_distance(storage, metric, Q, X + (size_t)i * rb, rb, qn, xn[i]);
    }
    free(X); free(Q); free(xn);
}

/* A11: vs_index/usearch.rs:1179-1205 */
void vso_f32_to_b1x8(const float* v, uint64_t n, uint8_t* out) {
    uint64_t nb = (n + 7) / 8;
    for (uint64_t j = 0; j < nb; ++j) {
        uint8_t b = 0;
        for (uint64_t i = 0; i < 8 && 8 * j + i < n; ++i)
            if (v[8 * j + i] > 0.0f) b |= (uint8_t)(1u << i);
        out[j] = b;
    }
}
