import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from importlib import import_module
import vector_store_b200 as v
ds = import_module("vector_store_b200.host.datasets")
n, dim = 1_000_000, 768
idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.F32, device=0, bf16_traversal=True)
idx.reserve(n)
for c0 in range(0, n, 100_000):
    xc = ds.embedding_like(100_000, dim, seed=1234 + c0 // 100_000)
    idx.add_batch(np.arange(c0, c0 + len(xc), dtype=np.uint64), xc)
idx.build()
q = torch.from_numpy(ds.embedding_like(1000, dim, seed=4321)).pin_memory()
k = 10
hk = torch.empty((1000, k), dtype=torch.int64).pin_memory()
hd = torch.empty((1000, k), dtype=torch.float32).pin_memory()
hc = torch.empty((1000,), dtype=torch.int32).pin_memory()
idx.set_search_params(expansion_search=160, search_width=2)
for w in range(40):
    idx.search_raw(q.data_ptr() + w * dim * 4, 1, k, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
