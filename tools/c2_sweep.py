#!/usr/bin/env python
"""BASELINE configs[1] in full: 1M x 768 f32 cosine, k in {10, 100}, query batches 1 .. 10 000.

For each k the smallest beam with recall@k >= 0.95 (vs exact GPU ground truth) is chosen on a 10 000-query
batch; then every batch size is timed END TO END through the host C ABI (pinned host buffers, H2D + search +
D2H inside the timed region), which is what a caller of the reference's index layer would see."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def recall(found, truth):
    hit = 0
    for f, t in zip(found, truth):
        hit += len(np.intersect1d(f, t, assume_unique=False))
    return hit / truth.size


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--ks", default="10,100")
    ap.add_argument("--batches", default="1,16,64,256,1024,10000")
    ap.add_argument("--target", type=float, default=0.95)
    ap.add_argument("--seconds", type=float, default=1.0, help="timed seconds per point")
    a = ap.parse_args()
    import torch
    from importlib import import_module
    import vector_store_b200 as v
    ds = import_module("vector_store_b200.host.datasets")
    idx = v.GpuIndex(a.dim, v.Metric.Cos, v.Scalar.F32, device=0, bf16_traversal=True)
    idx.reserve(a.rows)
    CH = 100_000
    for c0 in range(0, a.rows, CH):
        xc = ds.embedding_like(min(CH, a.rows - c0), a.dim, seed=1234 + c0 // CH)
        idx.add_batch(np.arange(c0, c0 + len(xc), dtype=np.uint64), xc)
    t0 = time.perf_counter()
    idx.build()
    print(json.dumps({"build_s": round(time.perf_counter() - t0, 3)}), flush=True)
    QN = 10_000
    q = torch.from_numpy(ds.embedding_like(QN, a.dim, seed=4321)).pin_memory()
    for k in [int(x) for x in a.ks.split(",")]:
        hk = torch.empty((QN, k), dtype=torch.int64).pin_memory()
        hd = torch.empty((QN, k), dtype=torch.float32).pin_memory()
        hc = torch.empty((QN,), dtype=torch.int32).pin_memory()
        idx.search_raw(q.data_ptr(), QN, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr(), exact=True)
        truth = hk.numpy().copy()
        chosen = None
        for ef in (32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 512, 768, 1024):
            if ef < k:
                continue
            idx.set_search_params(expansion_search=ef, search_width=2)
            idx.search_raw(q.data_ptr(), QN, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            r = recall(hk.numpy(), truth)
            print(json.dumps({"k": k, "ef": ef, f"recall_at_{k}": round(r, 4)}), flush=True)
            if r >= a.target:
                chosen = (ef, r)
                break
        if chosen is None:
            print(json.dumps({"k": k, "error": "target recall not reached"}), flush=True)
            continue
        for B in [int(x) for x in a.batches.split(",")]:
            # distinct query windows per call so that small batches do not re-walk a cached neighbourhood
            n_win = QN // B
            lat = []
            for w in range(min(n_win, 20)):  # warm-up
                idx.search_raw(q.data_ptr() + w * B * a.dim * 4, B, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            t_end = time.perf_counter() + a.seconds
            calls = 0
            t0 = time.perf_counter()
            while time.perf_counter() < t_end or calls < 5:
                w = calls % n_win
                t1 = time.perf_counter()
                idx.search_raw(q.data_ptr() + w * B * a.dim * 4, B, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
                lat.append(time.perf_counter() - t1)
                calls += 1
            el = time.perf_counter() - t0
            lat = np.array(lat) * 1e3
            print(json.dumps({"k": k, "ef": chosen[0], f"recall_at_{k}": round(chosen[1], 4), "batch": B,
                              "e2e_qps": round(calls * B / el, 1), "p50_ms": round(float(np.percentile(lat, 50)), 3),
                              "p99_ms": round(float(np.percentile(lat, 99)), 3), "calls": calls}), flush=True)


if __name__ == "__main__":
    main()
