#!/usr/bin/env python
"""Diagnostic: which queries does the graph search miss completely, and why?"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module
import vector_store_b200 as v
ds = import_module("vector_store_b200.host.datasets")
n, dim, nq, k = 200_000, 256, 1000, 10
x = ds.embedding_like(n, dim)
q = ds.embedding_like(nq, dim, seed=4321)
for trav in (True,):
    idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.F32, bf16_traversal=trav)
    idx.reserve(n)
    idx.add_batch(np.arange(n, dtype=np.uint64), x)
    idx.build()
    print(json.dumps({"extra_seeds": idx.stats()["extra_seeds"], "n_seed_rows": idx.stats()["n_seed_rows"]}), flush=True)
    tk, td, _ = idx.search_batch(q, k, exact=True)
    for ef in (128,):
        for batch in (1000, 1):
            idx.set_search_params(expansion_search=ef)
            if batch == 1:
                gk = np.stack([idx.search_batch(q[i:i + 1], k)[0][0] for i in range(nq)])
                gd = None
            else:
                gk, gd, gc = idx.search_batch(q, k)
            rec = np.array([len(np.intersect1d(gk[i], tk[i])) for i in range(nq)]) / k
            bad = np.where(rec == 0)[0]
            print(json.dumps({"bf16_traversal": trav, "ef": ef, "batch": batch, "recall": round(float(rec.mean()), 4),
                              "zero_recall_queries": bad.tolist()[:12], "n_zero": int(len(bad)),
                              "n_below_half": int((rec < 0.5).sum())}), flush=True)
            if len(bad) and gd is not None:
                i = int(bad[0])
                print(json.dumps({"query": i, "true_d": [round(float(d), 4) for d in td[i][:5]],
                                  "ann_d": [round(float(d), 4) for d in gd[i][:5]],
                                  "true_keys": tk[i][:5].tolist(), "ann_keys": gk[i][:5].tolist()}), flush=True)
    idx.close()
