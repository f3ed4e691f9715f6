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
    expect_tc = os.environ.get("VSB_DISABLE_TC") != "1" and storage in (O.BF16, O.F16) and len(x) >= 8192
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
