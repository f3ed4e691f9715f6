"""N4: parquet dataset readers (crates/benchmark/src/data/parquet.rs) — CPU only."""
import os

import numpy as np
import pytest

from conftest import embedding_like

pa = pytest.importorskip("pyarrow")


def DS():
    from importlib import import_module
    return import_module("vector_store_b200.host.datasets")


def _truth(rows, queries, k):
    d = ((queries[:, None, :] - rows[None, :, :]) ** 2).sum(-1)
    return np.argsort(d, axis=1, kind="stable")[:, :k]


@pytest.mark.parametrize("emb_type,files,rg", [("float32", 1, None), ("float64", 3, 70)])
def test_parquet_dataset_round_trip(tmp_path, emb_type, files, rg):
    ds = DS()
    rng = np.random.default_rng(3)
    n, dim, nq, k = 500, 24, 17, 10
    rows = embedding_like(n, dim, n_clusters=8)
    queries = embedding_like(nq, dim, seed=9, n_clusters=8)
    ids = rng.permutation(10_000)[:n].astype(np.int64)           # ids are arbitrary int64, not positions
    q_ids = np.arange(100, 100 + nq, dtype=np.int64)
    nb = ids[_truth(rows, queries, k)]
    d = str(tmp_path / "ds")
    ds.write_parquet_dataset(d, ids, rows, q_ids, queries, nb, train_files=files, row_group_rows=rg, emb_type=emb_type)
    open(os.path.join(d, "train.notparquet"), "w").write("x")     # wrong extension: ignored
    os.makedirs(os.path.join(d, "train_dir.parquet"))              # not a regular file: ignored
    assert len(ds.parquet_train_files(d)) == files
    assert ds.parquet_dimension(d) == dim
    got_ids, got_rows, batches = [], [], 0
    for bi, br in ds.parquet_vector_batches(d):
        assert br.dtype == np.float32 and br.flags["C_CONTIGUOUS"] and bi.dtype == np.int64
        got_ids.append(bi)
        got_rows.append(br)
        batches += 1
    assert batches >= files and (rg is None or batches > files)    # one batch per row group
    got_ids, got_rows = np.concatenate(got_ids), np.concatenate(got_rows)
    order = np.argsort(got_ids)
    ref = np.argsort(ids)
    assert np.array_equal(got_ids[order], ids[ref])
    assert np.array_equal(got_rows[order], rows[ref].astype(np.float32))   # f64 files narrow back to the same f32
    qs = ds.parquet_queries(d, limit=k)
    assert len(qs) == nq
    for (qv, truth), want_q, want_n in zip(qs, queries, nb):
        assert np.array_equal(qv, want_q) and truth == set(int(v) for v in want_n)


def test_parquet_queries_filter_limit_and_join(tmp_path):
    ds = DS()
    dim = 4
    rows = np.eye(dim, dtype=np.float32)
    ids = np.array([10, 11, 12, 13], np.int64)
    q_ids = np.array([1, 2, 3], np.int64)
    queries = np.stack([rows[0], rows[1], rows[2]])
    nb = np.array([[10, 11, 12], [11, 13, 10], [13, 13, 13]], np.int64)
    d = str(tmp_path / "ds")
    ds.write_parquet_dataset(d, ids, rows, q_ids, queries, nb)
    # id_ok drops 13 everywhere; limit keeps the first 2 survivors; query 3 loses all neighbours and disappears
    qs = ds.parquet_queries(d, id_ok=lambda i: i != 13, limit=2)
    assert [t for _, t in qs] == [{10, 11}, {11, 10}]
    assert np.array_equal(qs[0][0], rows[0]) and np.array_equal(qs[1][0], rows[1])
    # a query id without a ground-truth row is dropped by the join (parquet.rs:424-433)
    import pyarrow.parquet as pq
    t = pq.read_table(os.path.join(d, "neighbors.parquet")).slice(0, 1)
    pq.write_table(t, os.path.join(d, "neighbors.parquet"))
    assert len(ds.parquet_queries(d, limit=3)) == 1


def test_dataset_bench_tool_dry_run(tmp_path):
    """tools/dataset_bench.py writes a parquet dataset (ground truth from the oracle) and reads it back — no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = str(tmp_path / "ds")
    tool = os.path.join(root, "tools", "dataset_bench.py")
    subprocess.run([sys.executable, tool, "--make-synthetic", d, "--rows", "2000", "--dim", "16", "--queries", "12"], check=True,
                   capture_output=True)
    out = subprocess.run([sys.executable, tool, d, "--dry-run"], check=True, capture_output=True, text=True).stdout
    assert json.loads(out.strip().splitlines()[-1]) == {"rows": 2000, "dim": 16, "queries": 12, "truth_per_query": 10}
    # the written ground truth really is the brute-force top-10 of the written rows, addressed by the written ids
    ds = DS()
    ids, rows = map(np.concatenate, zip(*ds.parquet_vector_batches(d)))
    by_id = {int(i): r for i, r in zip(ids, rows)}
    for qv, truth in ds.parquet_queries(d, limit=10):
        dist = ((rows - qv) ** 2).sum(1)
        kth = np.sort(dist)[9]
        assert max(float(((by_id[t] - qv) ** 2).sum()) for t in truth) == float(kth)
