#!/usr/bin/env python
"""Where a small-batch search spends its time: per-phase CUDA-event times (vsb_set_kernel_timing) next to the
host-observed call latency, for batch sizes 1..256 and k in {10, 100} on the C2 index."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from importlib import import_module
    import vector_store_b200 as v
    ds = import_module("vector_store_b200.host.datasets")
    n, dim = int(os.environ.get("ROWS", 1_000_000)), 768
    idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.F32, device=0, bf16_traversal=True)
    idx.reserve(n)
    for c0 in range(0, n, 100_000):
        xc = ds.embedding_like(min(100_000, n - c0), dim, seed=1234 + c0 // 100_000)
        idx.add_batch(np.arange(c0, c0 + len(xc), dtype=np.uint64), xc)
    idx.build()
    q = torch.from_numpy(ds.embedding_like(10_000, dim, seed=4321)).pin_memory()
    # small-batch operating points: recall over 2 560 queries issued as 10 calls of 256 (the CTA-per-query path)
    k = 10
    hk = torch.empty((10_000, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((10_000, k), dtype=torch.float32).pin_memory()
    hc = torch.empty((10_000,), dtype=torch.int32).pin_memory()
    idx.search_raw(q.data_ptr(), 2560, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr(), exact=True)
    truth = hk.numpy()[:2560].copy()
    for width in [int(x) for x in os.environ.get("WIDTHS", "2").split(",")]:
        for ef in [int(x) for x in os.environ.get("EFS", "128,160,192").split(",")]:
            idx.set_search_params(expansion_search=ef, search_width=width)
            found = np.empty((2560, k), np.int64)
            for c in range(10):
                idx.search_raw(q.data_ptr() + c * 256 * dim * 4, 256, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
                found[c * 256:(c + 1) * 256] = hk.numpy()[:256]
            rec = np.mean([len(np.intersect1d(f, t)) for f, t in zip(found, truth)]) / k
            for w in range(20):
                idx.search_raw(q.data_ptr() + w * dim * 4, 1, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            lat = []
            for w in range(300):
                t1 = time.perf_counter()
                idx.search_raw(q.data_ptr() + (w % 100) * dim * 4, 1, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
                lat.append(time.perf_counter() - t1)
            lat = np.array(lat) * 1e3
            print(json.dumps({"width": width, "ef": ef, "recall_at_10": round(float(rec), 4),
                              "batch1_p50_ms": round(float(np.percentile(lat, 50)), 4),
                              "batch1_p99_ms": round(float(np.percentile(lat, 99)), 4)}), flush=True)
    for k, ef in ((10, 160), (100, 128)):
        hk = torch.empty((10_000, k), dtype=torch.int64).pin_memory()
        hd = torch.empty((10_000, k), dtype=torch.float32).pin_memory()
        hc = torch.empty((10_000,), dtype=torch.int32).pin_memory()
        idx.set_search_params(expansion_search=ef, search_width=2)
        for B in (1, 16, 256):
            for w in range(20):
                idx.search_raw(q.data_ptr() + w * B * dim * 4, B, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            calls = 200
            t0 = time.perf_counter()
            for w in range(calls):
                idx.search_raw(q.data_ptr() + (w % 30) * B * dim * 4, B, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            host_ms = (time.perf_counter() - t0) / calls * 1e3
            idx.set_kernel_timing(True)
            for w in range(calls):
                idx.search_raw(q.data_ptr() + (w % 30) * B * dim * 4, B, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
            st = idx.stats()
            idx.set_kernel_timing(False)
            ph = {p: round(st[p + "_ns"] / max(st[p + "_launches"], 1) / 1e6, 4)
                  for p in ("convert", "seed", "graph_search", "exact", "merge")}
            print(json.dumps({"k": k, "ef": ef, "batch": B, "host_call_ms": round(host_ms, 4), "phase_ms": ph,
                              "sum_phases_ms": round(sum(ph.values()), 4)}), flush=True)


if __name__ == "__main__":
    main()
