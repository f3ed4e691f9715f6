#!/usr/bin/env python
"""Profiling driver for the dense stages (run under ncu): builds a 131072-row bf16 index (the all-pairs stage of every
large build: 8 K1-TC launches of 16384 x 131072 x 768) and runs one exact search of 10 000 queries over 1 M rows."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vector_store_b200 as v  # noqa: E402
from importlib import import_module  # noqa: E402

ds = import_module("vector_store_b200.host.datasets")
mode = sys.argv[1] if len(sys.argv) > 1 else "allpairs"
dim = 768
dev = torch.device("cuda", 0)
n = 131072 if mode == "allpairs" else 1_000_000
buf = torch.empty((n, dim), dtype=torch.float32, device=dev)
ds.embedding_mix_dev(buf.data_ptr(), n, dim, row0=0, seed=1234, n_clusters=2560 if mode == "allpairs" else 256)
torch.cuda.synchronize()
idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.BF16, device=0)
idx.reserve(n)
idx.add_dev(np.arange(n, dtype=np.uint64), buf.data_ptr(), n)
del buf
if mode == "allpairs":
    for rep in range(2):
        t0 = time.perf_counter()
        idx.build()
        bs = idx.build_stats()
        print(f"build {rep}: {time.perf_counter() - t0:.3f} s, all-pairs {bs['allpairs_ns'] / 1e6:.1f} ms = "
              f"{bs['allpairs_flops'] / bs['allpairs_ns'] / 1e3:.0f} TFLOP/s, prune {bs['prune_ns'] / 1e6:.1f} ms")
else:
    qb = torch.empty((10_000, dim), dtype=torch.float32, device=dev)
    ds.embedding_mix_dev(qb.data_ptr(), 10_000, dim, row0=0, seed=4321, n_clusters=256)
    q = qb.cpu().numpy()
    idx.set_kernel_timing(True)
    for rep in range(4):
        s0 = idx.stats()
        t0 = time.perf_counter()
        idx.search_batch(q, 10, exact=True)
        dt = time.perf_counter() - t0
        s1 = idx.stats()
        dev_ms = (s1["exact_ns"] - s0["exact_ns"]) / 1e6
        print(f"exact search {rep}: {dt * 1e3:.1f} ms host incl. copies; K1-TC + K3 on the device {dev_ms:.2f} ms = "
              f"{2 * 10_000 * n * dim / (dev_ms * 1e-3) / 1e12:.0f} TFLOP/s")
idx.close()
