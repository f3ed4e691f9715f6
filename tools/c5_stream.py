#!/usr/bin/env python
"""BASELINE config #5: streaming CDC-style mutations interleaved with queries.

1M x 768 cosine corpus; a mutation stream of `--rate` ops/s (70 % insert new key, 20 % delete,
10 % update = RemoveBeforeAdd with a bumped epoch, exactly the message sequences of SURVEY Appendix A) is applied
in ticks of 100 ms through the actor mirror's batched calls, interleaved with query batches of 1 and 1000;
recall@10 is measured against exact ground truth on the live set at checkpoints."""
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
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--rate", type=int, default=10_000)
    ap.add_argument("--ef", type=int, default=160)
    a = ap.parse_args()
    from importlib import import_module
    import vector_store_b200 as v
    ds = import_module("vector_store_b200.host.datasets")
    rng = np.random.default_rng(5)
    idx = v.GpuIndex(a.dim, v.Metric.Cos, v.Scalar.F32, device=0, bf16_traversal=True)
    extra = int(a.rate * a.seconds * 0.8) + 10_000
    idx.reserve(a.n + extra)
    EPOCH = np.uint64(1) << np.uint64(48)
    CH = 100_000
    for c0 in range(0, a.n, CH):
        xc = ds.embedding_like(min(CH, a.n - c0), a.dim, seed=1234 + c0 // CH)
        idx.add_batch(np.arange(c0, c0 + len(xc), dtype=np.uint64), xc)
    t0 = time.perf_counter()
    idx.build()
    print(json.dumps({"phase": "initial build", "s": round(time.perf_counter() - t0, 2)}), flush=True)
    idx.set_search_params(expansion_search=a.ef, search_width=2, stream_threshold=4096)
    pool = ds.embedding_like(extra, a.dim, seed=777)        # vectors for inserts / updates
    queries = ds.embedding_like(20_000, a.dim, seed=4321)
    cap_rows = a.n + extra
    alive = np.zeros(cap_rows, dtype=bool)                  # table row id -> is it live
    alive[:a.n] = True
    cur_key = np.arange(cap_rows, dtype=np.uint64)          # row id -> current key (epoch in the high 16 bits)
    next_row, pool_pos = a.n, 0
    tick_ops = max(1, a.rate // 10)
    lat1, q1000, applied = [], [], 0
    checkpoints = []
    t_start = time.perf_counter()
    tick = 0
    while True:
        now = time.perf_counter() - t_start
        if now >= a.seconds:
            break
        # ---- one 100 ms tick of mutations ----
        n_ins = int(tick_ops * 0.7)
        n_del = int(tick_ops * 0.2)
        n_upd = tick_ops - n_ins - n_del
        cand = np.unique(rng.integers(0, next_row, size=4 * (n_del + n_upd)))
        cand = rng.permutation(cand[alive[cand]])[:n_del + n_upd]
        del_rows, upd_rows = cand[:n_del], cand[n_del:]
        t1 = time.perf_counter()
        idx.remove_batch(cur_key[cand])                                      # RemoveValue / RemoveBeforeAddValue
        alive[del_rows] = False
        cur_key[upd_rows] += EPOCH                                           # same row id, bumped epoch
        ins_rows = np.arange(next_row, next_row + n_ins)
        alive[ins_rows] = True
        next_row += n_ins
        new_keys = np.concatenate([cur_key[upd_rows], cur_key[ins_rows]])
        vecs = pool[pool_pos:pool_pos + len(new_keys)]
        pool_pos += len(new_keys)
        idx.add_batch(new_keys, vecs)                                        # AddVector (K7 links them in batches)
        applied += len(cand) + n_ins
        mut_s = time.perf_counter() - t1
        # ---- interleaved queries ----
        for _ in range(5):
            i = rng.integers(0, len(queries))
            t2 = time.perf_counter()
            idx.search_batch(queries[i:i + 1], 10)
            lat1.append(time.perf_counter() - t2)
        j = rng.integers(0, len(queries) - 1000)
        t2 = time.perf_counter()
        idx.search_batch(queries[j:j + 1000], 10)
        q1000.append(time.perf_counter() - t2)
        tick += 1
        # pace to the requested rate
        target = tick * 0.1
        now = time.perf_counter() - t_start
        if now < target:
            time.sleep(target - now)
        if tick % 100 == 0 or (time.perf_counter() - t_start) >= a.seconds:
            qs = queries[:500]
            tk, _, _ = idx.search_batch(qs, 10, exact=True)
            gk, _, _ = idx.search_batch(qs, 10)
            rec = np.mean([len(np.intersect1d(gk[i], tk[i])) for i in range(len(qs))]) / 10
            st = idx.stats()
            checkpoints.append({"t": round(time.perf_counter() - t_start, 1), "recall_at_10": round(float(rec), 4),
                                "live": idx.size(), "n_slots": st["n_slots"], "n_graphed": st["n_graphed"],
                                "last_tick_mutation_ms": round(mut_s * 1e3, 2)})
            print(json.dumps(checkpoints[-1]), flush=True)
    wall = time.perf_counter() - t_start
    print(json.dumps({"config": f"C5 {a.n}x{a.dim} f32 cosine, {a.rate} mutations/s requested for {a.seconds}s",
                      "mutations_applied": applied, "mutations_per_s": round(applied / wall),
                      "p50_batch1_ms": round(float(np.percentile(lat1, 50) * 1e3), 3),
                      "p99_batch1_ms": round(float(np.percentile(lat1, 99) * 1e3), 3),
                      "batch1000_qps": round(1000 / float(np.mean(q1000))),
                      "checkpoints": checkpoints}), flush=True)


if __name__ == "__main__":
    main()
