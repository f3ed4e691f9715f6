#!/usr/bin/env python
"""N4: run a dataset directory (parquet: VectorDBBench layout, or .fbin/.ibin) through the engine and report what the
reference's `vector-search-benchmark search-*` reports (crates/benchmark/src/main.rs:529-698): queries, QPS, latency
min / P1 / P10 / P25 / P50 / P75 / P90 / P99 / max, recall min / avg / max with the per-query recall of
db.rs:308 (|truth ∩ found| / |truth|, limit = |truth|).

  python tools/dataset_bench.py --make-synthetic /tmp/ds --rows 200000 --dim 128     # writes a parquet dataset
  python tools/dataset_bench.py /tmp/ds --metric l2sq --batch 1 --seconds 5
  python tools/dataset_bench.py /tmp/ds --dry-run                                   # readers only, no GPU
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_dataset(path, limit):
    """-> (batches iterator of (ids, rows), list of (query, truth set), dim)"""
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    if os.path.exists(os.path.join(path, "test.parquet")):
        return ds.parquet_vector_batches(path), ds.parquet_queries(path, limit=limit), ds.parquet_dimension(path)
    # fbin layout (data/fbin.rs): base.fbin, query.fbin, groundtruth.ibin
    base, query, gt = (os.path.join(path, n) for n in ("base.fbin", "query.fbin", "groundtruth.ibin"))
    n, dim = ds.read_bin_header(base)

    def batches():
        for s in range(0, n, 100_000):
            rows = ds.read_fbin(base, s, 100_000)
            yield np.arange(s, s + len(rows), dtype=np.int64), rows
    q = ds.read_fbin(query)
    g = ds.read_ibin(gt)
    lim = min(limit, g.shape[1])
    return batches(), [(q[i], set(int(v) for v in g[i, :lim])) for i in range(len(q))], dim


def report(lat_s, recalls, duration, label=""):
    lat = np.sort(np.asarray(lat_s))
    out = {"queries": len(recalls), "QPS": round(len(recalls) / duration, 1), "latency_min_ms": round(lat[0] * 1e3, 3)}
    for p in (1, 10, 25, 50, 75, 90, 99):
        out[f"latency_P{p}_ms"] = round(float(np.percentile(lat, p)) * 1e3, 3)
    out["latency_max_ms"] = round(lat[-1] * 1e3, 3)
    out.update({"recall_min": round(100 * min(recalls), 1), "recall_avg": round(100 * float(np.mean(recalls)), 1),
                "recall_max": round(100 * max(recalls), 1)})
    if label:
        out["label"] = label
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path", nargs="?")
    ap.add_argument("--make-synthetic", metavar="DIR")
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--kind", default="sift", choices=["sift", "embedding"],
                    help="--make-synthetic: SIFT-shaped (ground truth by L2) or embedding-shaped unit vectors (cosine)")
    ap.add_argument("--exact", action="store_true", help="brute-force search (recall must be 100: checks the id plumbing)")
    ap.add_argument("--metric", default="l2sq", choices=["l2sq", "cos", "ip"])
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1, help="queries per vsb_search call (1 = the reference's one query per request)")
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    if a.make_synthetic:
        gen = ds.sift_like if a.kind == "sift" else ds.embedding_like
        rows = gen(a.rows, a.dim)
        queries = gen(a.queries, a.dim, seed=4321)
        ids = np.arange(1, a.rows + 1, dtype=np.int64) * 7                       # ids are not positions
        import oracle as O                                                        # ground truth for a test dataset only
        tk, _, _, _ = O.exact_topk(rows, queries, a.limit, O.L2SQ if a.kind == "sift" else O.COS, O.F32,
                                   keys=ids.astype(np.uint64))
        ds.write_parquet_dataset(a.make_synthetic, ids, rows, np.arange(a.queries, dtype=np.int64), queries,
                                 tk.astype(np.int64), train_files=2, row_group_rows=50_000)
        print(json.dumps({"written": a.make_synthetic, "rows": a.rows, "dim": a.dim, "queries": a.queries}))
        return
    batches, queries, dim = load_dataset(a.path, a.limit)
    if a.dry_run:
        n = sum(len(i) for i, _ in batches)
        print(json.dumps({"rows": n, "dim": dim, "queries": len(queries), "truth_per_query": len(queries[0][1])}))
        return
    import vector_store_b200 as v
    metric = {"l2sq": v.Metric.L2sq, "cos": v.Metric.Cos, "ip": v.Metric.IP}[a.metric]
    idx = v.GpuIndex(dim, metric, v.Scalar.F32, bf16_traversal=True)
    t0 = time.perf_counter()
    n = 0
    for ids, rows in batches:
        idx.reserve(n + len(ids))
        idx.add_batch(ids.astype(np.uint64), rows)
        n += len(ids)
    idx.build()
    print(json.dumps({"rows": n, "dim": dim, "build_s": round(time.perf_counter() - t0, 3)}), flush=True)
    idx.set_search_params(expansion_search=a.ef)
    qm = np.stack([q for q, _ in queries])
    truth = [t for _, t in queries]
    k = max(len(t) for t in truth)
    rng = np.random.default_rng(0)
    lat, rec = [], []
    idx.search_batch(qm[:a.batch], k)
    start = time.perf_counter()
    while time.perf_counter() - start < a.seconds:
        pick = rng.integers(0, len(queries), a.batch)                            # main.rs:535 random(&queries)
        t1 = time.perf_counter()
        keys, _, counts = idx.search_batch(qm[pick], k, exact=a.exact)
        dt = time.perf_counter() - t1
        for j, qi in enumerate(pick):
            found = set(int(x) for x in keys[j, :counts[j]][:len(truth[qi])])
            rec.append(len(truth[qi] & found) / len(truth[qi]))
            lat.append(dt)
    print(json.dumps(report(lat, rec, time.perf_counter() - start, f"batch {a.batch}, ef {a.ef}")), flush=True)


if __name__ == "__main__":
    main()
