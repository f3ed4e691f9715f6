#!/usr/bin/env python
"""Exploration helper (not part of the bench contract): build one index, then sweep the graph-search
operating point (expansion_search x search_width) and print recall@10, QPS, distance evaluations and the
K4 kernel time / achieved HBM bandwidth for each."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=10_000)
    ap.add_argument("--storage", default="f32")
    ap.add_argument("--efs", default="64,128,192,256")
    ap.add_argument("--widths", default="1,2,4")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bf16-traversal", action="store_true")
    ap.add_argument("--i8-traversal", action="store_true")
    ap.add_argument("--exact-only", action="store_true", help="skip the graph build and the ANN sweep")
    a = ap.parse_args()
    import torch
    from importlib import import_module
    import vector_store_b200 as v
    ds = import_module("vector_store_b200.host.datasets")
    scalar = {"f32": v.Scalar.F32, "bf16": v.Scalar.BF16, "f16": v.Scalar.F16}[a.storage]
    idx = v.GpuIndex(a.dim, v.Metric.Cos, scalar, device=0, bf16_traversal=a.bf16_traversal, i8_traversal=a.i8_traversal)
    idx.reserve(a.n)
    CH = 100_000
    for c0 in range(0, a.n, CH):
        xc = ds.embedding_like(min(CH, a.n - c0), a.dim, seed=1234 + c0 // CH)
        idx.add_batch(np.arange(c0, c0 + len(xc), dtype=np.uint64), xc)
    t0 = time.perf_counter()
    if not a.exact_only:
        idx.build()
    print(json.dumps({"build_s": time.perf_counter() - t0, "stats": idx.stats()}), flush=True)
    dev = torch.device("cuda", 0)
    B, k = a.batch, 10
    q = torch.from_numpy(ds.embedding_like(B, a.dim, seed=4321)).to(dev)
    ok = torch.empty((B, k), dtype=torch.int64, device=dev)
    od = torch.empty((B, k), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    idx.search_dev(q.data_ptr(), B, k, ok.data_ptr(), od.data_ptr(), 0, stream, True)
    torch.cuda.synchronize()
    gt = ok.cpu().numpy().copy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st0 = idx.stats()
    tc0 = st0["tc_launches"]
    e0.record()
    for _ in range(3):
        idx.search_dev(q.data_ptr(), B, k, ok.data_ptr(), od.data_ptr(), 0, stream, True)
    e1.record()
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1) / 3
    print(json.dumps({"exact_search_ms": round(ems, 2), "tflops": round(2.0 * B * a.n * a.dim / ems / 1e9, 1),
                      "tensor_core_path": idx.stats()["tc_launches"] > tc0,
                      "certified": (idx.stats()["exact_certified"] - st0["exact_certified"]) // 3,
                      "fallback": (idx.stats()["exact_fallback"] - st0["exact_fallback"]) // 3, "same_as_first": bool((ok.cpu().numpy() == gt).all())}),
          flush=True)
    peak = 6541.1
    if a.exact_only:
        return
    for w in [int(x) for x in a.widths.split(",")]:
        for ef in [int(x) for x in a.efs.split(",")]:
            idx.set_search_params(expansion_search=ef, search_width=w)
            idx.set_instrumented(True)
            idx.search_dev(q.data_ptr(), B, k, ok.data_ptr(), od.data_ptr(), 0, stream, False)
            torch.cuda.synchronize()
            st = idx.stats()
            idx.set_instrumented(False)
            got = ok.cpu().numpy()
            rec = np.mean([len(np.intersect1d(got[i], gt[i])) for i in range(0, B, 5)]) / k
            E = st["distance_evals"] / B
            P = st["parent_expansions"] / B
            idx.set_kernel_timing(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                idx.search_dev(q.data_ptr(), B, k, ok.data_ptr(), od.data_ptr(), 0, stream, False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            st = idx.stats()
            idx.set_kernel_timing(False)
            k4 = st["graph_search_ns"] / 1e6 / max(st["graph_search_launches"], 1)
            seed = st["seed_ns"] / 1e6 / max(st["seed_launches"], 1)
            rb = st["row_bytes"] // 4 if a.i8_traversal else (st["row_bytes"] // 2 if a.bf16_traversal else st["row_bytes"])
            exact_ms = st["exact_ns"] / 1e6 / max(st["exact_launches"], 1)
            bpq = E * (rb + 4) + P * st["graph_degree"] * 4
            gbs = B * bpq / (k4 * 1e-3) / 1e9
            print(json.dumps({"width": w, "ef": ef, "recall": round(float(rec), 4), "qps": round(B / ms * 1e3),
                              "E": round(E, 1), "P": round(P, 1), "k4_ms": round(k4, 3), "seed_ms": round(seed, 3), "rerank_ms": round(exact_ms, 3),
                              "GBps": round(gbs), "frac": round(gbs / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
