"""tcgen05/TMA distance-tile kernel (exact_tc.cu): parity against the oracle through the C ABI.
16-bit storages multiply exactly on the tensor cores, so after the canonical re-rank the results
must be bit-identical to oracle/exact.c, exactly like the SIMT path."""
import os

import numpy as np
import pytest

import oracle as O
from conftest import embedding_like

pytestmark = pytest.mark.gpu


def V():
    import vector_store_b200 as v
    return v


def run_exact(x, keys, q, k, metric, storage, dead=None):
    v = V()
    idx = v.GpuIndex(x.shape[1], v.Metric(metric), v.Scalar(storage))
    idx.reserve(len(x))
    idx.add_batch(keys, x)
    if dead is not None:
        idx.remove_batch(keys[dead])
    tc0 = idx.stats()["tc_launches"]
    out = idx.search_batch(q, k, exact=True)
    tc_used = idx.stats()["tc_launches"] - tc0
    tc_storages = (O.BF16, O.F16) if os.environ.get("VSB_DISABLE_CERT") == "1" else (O.BF16, O.F16, O.F32)
    expect_tc = os.environ.get("VSB_DISABLE_TC") != "1" and storage in tc_storages and len(x) >= 8192
    assert (tc_used > 0) == expect_tc, f"tcgen05 launches: {tc_used}, expected the TC path: {expect_tc}"
    idx.close()
    return out


@pytest.mark.parametrize("storage,metric,n,dim,nq,k", [
    (O.BF16, O.COS, 20000, 128, 200, 10),
    (O.F16, O.L2SQ, 12001, 100, 77, 10),     # ragged K (208-byte rows: second slab is mostly TMA zero fill)
    (O.BF16, O.IP, 10000, 768, 64, 32),
    (O.BF16, O.L2SQ, 9000, 64, 300, 100),    # half-empty K slab, k' = 160
])
def test_tc_exact_bit_parity(storage, metric, n, dim, nq, k):
    rng = np.random.default_rng(n + dim)
    x = embedding_like(n, dim, n_clusters=32) * (1.0 + rng.random((n, 1)).astype(np.float32))
    q = embedding_like(nq, dim, seed=4321, n_clusters=32)
    keys = rng.permutation(n).astype(np.uint64)
    dead = rng.choice(n, n // 10, replace=False)
    alive = np.ones(n, np.uint8)
    alive[dead] = 0
    gk, gd, gc = run_exact(x, keys, q, k, metric, storage, dead)
    ok, od, oc, _ = O.exact_topk(x, q, k, metric, storage, keys=keys, alive=alive)
    assert np.array_equal(gc, oc)
    assert np.array_equal(gk, ok), f"{(gk != ok).any(axis=1).sum()} of {nq} queries differ"
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_tc_matches_simt_path():
    rng = np.random.default_rng(9)
    n, dim = 16384, 256
    x = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((130, dim)).astype(np.float32)
    keys = np.arange(n, dtype=np.uint64)
    a = run_exact(x, keys, q, 10, O.COS, O.BF16)
    os.environ["VSB_DISABLE_TC"] = "1"
    try:
        b = run_exact(x, keys, q, 10, O.COS, O.BF16)
    finally:
        del os.environ["VSB_DISABLE_TC"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_tf32_candidates_feed_the_graph_build():
    # f32 storage: the build's kNN lists come from kind::tf32 tiles + canonical re-rank
    n, dim, k = 20000, 96, 10
    x = embedding_like(n, dim, n_clusters=32)
    q = embedding_like(300, dim, seed=4321, n_clusters=32)
    v = V()
    idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.F32)
    idx.reserve(n)
    idx.add_batch(np.arange(n, dtype=np.uint64), x)
    idx.build()
    tk, _, _ = idx.search_batch(q, k, exact=True)
    gk, _, _ = idx.search_batch(q, k)
    r = O.recall_at_k(gk, tk)
    print(f"tf32-built graph recall@10 = {r:.4f}")
    assert r >= 0.95


# ---- exact search on f32 rows: TF32 tensor-core candidates + per-query certificate + SIMT fallback -------------

def _exact_f32(x, keys, q, k, metric, dead=None, env=None, allow=None):
    v = V()
    old = {}
    for name, val in (env or {}).items():
        old[name] = os.environ.get(name)
        os.environ[name] = val
    try:
        idx = v.GpuIndex(x.shape[1], v.Metric(metric), v.Scalar.F32)
    finally:
        for name, val in old.items():
            if val is None:
                del os.environ[name]
            else:
                os.environ[name] = val
    idx.reserve(len(x))
    idx.add_batch(keys, x)
    if dead is not None:
        idx.remove_batch(keys[dead])
    s0 = idx.stats()
    out = idx.search_batch(q, k, exact=True) if allow is None else idx.search_filtered(q, k, allow)
    s1 = idx.stats()
    idx.close()
    return out, {n: s1[n] - s0[n] for n in ("exact_certified", "exact_fallback", "exact_scanned", "tc_launches")}


@pytest.mark.parametrize("metric,n,dim,nq,k", [
    (O.COS, 20000, 128, 257, 10),
    (O.L2SQ, 12001, 100, 77, 10),
    (O.IP, 10000, 768, 64, 32),
    (O.COS, 9000, 64, 300, 100),
])
def test_certified_tf32_exact_is_bit_identical(metric, n, dim, nq, k):
    rng = np.random.default_rng(n + dim)
    x = embedding_like(n, dim, n_clusters=32) * (1.0 + rng.random((n, 1)).astype(np.float32))
    q = embedding_like(nq, dim, seed=4321, n_clusters=32)
    keys = rng.permutation(n).astype(np.uint64)
    dead = rng.choice(n, n // 10, replace=False)
    alive = np.ones(n, np.uint8)
    alive[dead] = 0
    (gk, gd, gc), st = _exact_f32(x, keys, q, k, metric, dead)
    print(f"metric {metric}: certified {st['exact_certified']}, fallback {st['exact_fallback']} of {nq}")
    assert st["tc_launches"] > 0 and st["exact_certified"] + st["exact_fallback"] == nq
    ok, od, oc, _ = O.exact_topk(x, q, k, metric, O.F32, keys=keys, alive=alive)
    assert np.array_equal(gc, oc)
    assert np.array_equal(gk, ok), f"{(gk != ok).any(axis=1).sum()} of {nq} queries differ"
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_certificate_rejects_near_ties_and_the_fallback_is_exact():
    # rows that differ from each other far below TF32 resolution: the candidate stage cannot order them, the
    # certificate must refuse, and the SIMT re-run must still produce the oracle's answer bit for bit
    rng = np.random.default_rng(77)
    n, dim, nq, k = 10000, 64, 96, 10
    base = rng.standard_normal((40, dim)).astype(np.float32)
    x = base[rng.integers(0, 40, n)] * (1.0 + 1e-6 * rng.standard_normal((n, 1)).astype(np.float32))
    x += (1e-6 * rng.standard_normal((n, dim))).astype(np.float32)
    x[:50] = x[50:100]                                   # exact duplicates: ties broken by key
    q = base[rng.integers(0, 40, nq)] + (1e-3 * rng.standard_normal((nq, dim))).astype(np.float32)
    keys = rng.permutation(n).astype(np.uint64)
    for metric in (O.L2SQ, O.COS, O.IP):
        (gk, gd, gc), st = _exact_f32(x, keys, q, k, metric)
        print(f"metric {metric}: certified {st['exact_certified']}, fallback {st['exact_fallback']}, "
              f"scanned {st['exact_scanned']} of {nq}")
        assert st["exact_fallback"] > 0 and st["exact_scanned"] > 0
        ok, od, oc, _ = O.exact_topk(x, q, k, metric, O.F32, keys=keys)
        assert np.array_equal(gc, oc) and np.array_equal(gk, ok)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_certified_path_equals_simt_path_and_small_lists_still_exact():
    rng = np.random.default_rng(5)
    n, dim, nq, k = 30000, 96, 500, 10
    x = rng.standard_normal((n, dim)).astype(np.float32)          # iid: the hardest case for the certificate
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    keys = np.arange(n, dtype=np.uint64)
    ref, st0 = _exact_f32(x, keys, q, k, O.COS, env={"VSB_DISABLE_CERT": "1"})
    assert st0["tc_launches"] == 0 and st0["exact_certified"] == 0
    for kp in ("32", "128", "256"):
        out, st = _exact_f32(x, keys, q, k, O.COS, env={"VSB_CERT_KP": kp})
        print(f"kp {kp}: certified {st['exact_certified']}, fallback {st['exact_fallback']}, scanned {st['exact_scanned']}")
        assert st["exact_certified"] + st["exact_fallback"] == nq
        assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1].view(np.uint32), ref[1].view(np.uint32))
        assert np.array_equal(out[2], ref[2])


def test_certified_filtered_search_on_f32_rows():
    rng = np.random.default_rng(6)
    n, dim, nq, k = 16000, 128, 100, 10
    x = embedding_like(n, dim, n_clusters=16)
    q = embedding_like(nq, dim, seed=99, n_clusters=16)
    keys = np.arange(n, dtype=np.uint64)
    allow = rng.random(n) < 0.02                                  # sparse filter: many lists are not full
    (gk, gd, gc), st = _exact_f32(x, keys, q, k, O.L2SQ, allow=allow)
    ok, od, oc, _ = O.exact_topk(x, q, k, O.L2SQ, O.F32, keys=keys, alive=allow.astype(np.uint8))
    assert st["exact_certified"] + st["exact_fallback"] == nq
    assert np.array_equal(gc, oc) and np.array_equal(gk, ok)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("storage", [O.BF16, O.F16, O.F32])
def test_near_ties_small_corpus_goes_through_the_simt_stage_and_the_scan(storage):
    # below tc_min_rows the first stage is the SIMT tiles; same certificate, same last resort
    rng = np.random.default_rng(78)
    n, dim, nq, k = 3000, 48, 40, 10
    base = rng.standard_normal((8, dim)).astype(np.float32)
    x = base[rng.integers(0, 8, n)] + (2e-3 * rng.standard_normal((n, dim))).astype(np.float32)
    q = base[rng.integers(0, 8, nq)]
    keys = rng.permutation(n).astype(np.uint64)
    v = V()
    idx = v.GpuIndex(dim, v.Metric.L2sq, v.Scalar(storage))
    idx.reserve(n)
    idx.add_batch(keys, x)
    gk, gd, gc = idx.search_batch(q, k, exact=True)
    st = idx.stats()
    idx.close()
    print(f"storage {storage}: certified {st['exact_certified']}, fallback {st['exact_fallback']}, scanned {st['exact_scanned']}")
    assert st["exact_certified"] + st["exact_fallback"] == nq and st["exact_scanned"] == st["exact_fallback"]
    ok, od, oc, _ = O.exact_topk(x, q, k, O.L2SQ, storage, keys=keys)
    assert np.array_equal(gc, oc) and np.array_equal(gk, ok)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
