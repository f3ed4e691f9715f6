#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): CDC-style mutations at a fixed rate CONCURRENT with queries.

  1 M x 768 f32 cosine (bf16 traversal), mutation stream 10 000 ops/s for 60 s:
  70 % insert of a new key, 20 % delete, 10 % update (= delete + insert with a bumped epoch, Appendix A of the survey)
  — issued by a MUTATOR thread in batches of `--mut-batch` ops at wall-clock pace;
  a QUERY thread issues batch-1 searches back to back (and a 1000-query batch every `--big-every` s) the whole time;
  recall@10 vs exact ground truth of the live set at 3 checkpoints (GT recomputed by vsb_search_exact, which is
  bit-equal to the oracle — tests/test_gpu_round2.py).

All mutation batches are generated BEFORE the clock starts, so the two Python threads spend their time inside
ctypes calls (GIL released) and the latencies are the library's, not the interpreter's.  The reference serialises
this mix through its reader/writer gate (usearch.rs:515-624: Remove is exclusive and drains every search);
here searches run on the published view while the mutator prepares the next one.
Prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--rate", type=int, default=10_000, help="mutations per second")
    ap.add_argument("--mut-batch", type=int, default=500)
    ap.add_argument("--big-every", type=float, default=1.0)
    ap.add_argument("--clusters", type=int, default=256)
    ap.add_argument("--target-recall", type=float, default=0.96, help="operating point is tuned to this on the fresh build")
    a = ap.parse_args()

    import torch
    import vector_store_b200 as v
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    dev = torch.device("cuda", 0)
    n0, dim, k = a.rows, a.dim, 10
    total_ops = int(a.rate * a.seconds)
    n_batches = total_ops // a.mut_batch
    n_ins_b, n_del_b, n_upd_b = int(a.mut_batch * 0.7), int(a.mut_batch * 0.2), a.mut_batch - int(a.mut_batch * 0.7) - int(a.mut_batch * 0.2)
    n_new = n_batches * n_ins_b

    # ---- corpus (generated in HBM) + the rows the stream will insert (host, the CDC path delivers host vectors) ----
    idx = v.GpuIndex(dim, v.Metric.Cos, v.Scalar.F32, device=0, bf16_traversal=True)
    idx.reserve(n0 + n_new + 4096)
    buf = torch.empty((min(n0, 500_000), dim), dtype=torch.float32, device=dev)
    for c0 in range(0, n0, len(buf)):
        nb = min(len(buf), n0 - c0)
        ds.embedding_mix_dev(buf.data_ptr(), nb, dim, row0=c0, seed=1234, n_clusters=a.clusters)
        torch.cuda.synchronize()
        idx.add_dev(np.arange(c0, c0 + nb, dtype=np.uint64), buf.data_ptr(), nb)
    t0 = time.perf_counter()
    idx.build()
    t_build = time.perf_counter() - t0
    nbuf = torch.empty((n_new, dim), dtype=torch.float32, device=dev)
    ds.embedding_mix_dev(nbuf.data_ptr(), n_new, dim, row0=n0, seed=1234, n_clusters=a.clusters)
    new_rows = nbuf.cpu().numpy()
    qbuf = torch.empty((4000, dim), dtype=torch.float32, device=dev)
    ds.embedding_mix_dev(qbuf.data_ptr(), 4000, dim, row0=0, seed=4321, n_clusters=a.clusters)
    q = qbuf.cpu().numpy()
    del buf, nbuf, qbuf

    # ---- operating point on the fresh build ----
    tk, _, _ = idx.search_batch(q[:2000], k, exact=True)

    def recall_now(truth=None):
        t = truth if truth is not None else idx.search_batch(q[:2000], k, exact=True)[0]
        g, _, _ = idx.search_batch(q[:2000], k)
        hits = sum(len(np.intersect1d(g[i], t[i])) for i in range(len(g)))
        return hits / (len(g) * k)

    ef_used, r0 = None, 0.0
    for ef in (96, 128, 160, 192, 224, 256, 320, 384):
        idx.set_search_params(expansion_search=ef, search_width=2)
        r0 = recall_now(tk)
        ef_used = ef
        if r0 >= a.target_recall:
            break

    # ---- the mutation stream, precomputed: rows already updated or deleted are not picked again ----
    rng = np.random.default_rng(11)
    pool = rng.permutation(n0)[: n_batches * (n_del_b + n_upd_b)]
    batches = []
    for b in range(n_batches):
        sel = pool[b * (n_del_b + n_upd_b):(b + 1) * (n_del_b + n_upd_b)]
        dele, upd = sel[:n_del_b], sel[n_del_b:]
        ins = np.arange(n0 + b * n_ins_b, n0 + (b + 1) * n_ins_b)
        rm_keys = np.concatenate([dele, upd]).astype(np.uint64)                       # epoch 0 keys
        add_keys = np.concatenate([ins.astype(np.uint64), upd.astype(np.uint64) | np.uint64(1 << 48)])
        # an update re-inserts the row's vector under the bumped epoch (new vector = a fresh draw of the stream)
        add_rows = np.ascontiguousarray(np.concatenate([new_rows[b * n_ins_b:(b + 1) * n_ins_b],
                                                        new_rows[(b * n_upd_b) % (n_new - n_upd_b):][:n_upd_b] * np.float32(1.0)]))
        batches.append((rm_keys, add_keys, add_rows))

    stop = threading.Event()
    lat1, lat1_t, lat_big, errors = [], [], [], []
    q1 = [np.ascontiguousarray(q[i:i + 1]) for i in range(2000, 4000)]
    qbig = np.ascontiguousarray(q[:1000])

    def query_thread():
        try:
            i, next_big = 0, time.perf_counter() + a.big_every
            while not stop.is_set():
                t1 = time.perf_counter()
                idx.search_batch(q1[i % len(q1)], k)
                t2 = time.perf_counter()
                lat1.append(t2 - t1)
                lat1_t.append(t2)
                i += 1
                if t2 >= next_big:
                    t1 = time.perf_counter()
                    idx.search_batch(qbig, k)
                    lat_big.append(time.perf_counter() - t1)
                    next_big += a.big_every
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    checkpoints = []
    mut_calls = []  # (start, end) of every mutator call
    behind = 0.0

    def mutator_thread():
        nonlocal behind
        try:
            t_start = time.perf_counter()
            for b, (rm_keys, add_keys, add_rows) in enumerate(batches):
                due = t_start + b * a.mut_batch / a.rate
                now = time.perf_counter()
                if now < due:
                    time.sleep(due - now)
                else:
                    behind = max(behind, now - due)
                t1 = time.perf_counter()
                idx.remove_batch(rm_keys)
                idx.add_batch(add_keys, add_rows)
                mut_calls.append((t1, time.perf_counter()))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    qt = threading.Thread(target=query_thread)
    mt = threading.Thread(target=mutator_thread)
    t_begin = time.perf_counter()
    qt.start()
    mt.start()
    # recall checkpoints at 1/3, 2/3 and the end (main thread; exact GT of the live set at that moment)
    ck_windows = []
    for frac in (1 / 3, 2 / 3):
        time.sleep(max(0.0, t_begin + frac * a.seconds - time.perf_counter()))
        st = idx.stats()
        ck0 = round(time.perf_counter() - t_begin, 3)
        checkpoints.append({"t_s": round(time.perf_counter() - t_begin, 1), "recall_at_10": round(recall_now(), 4),
                            "n_slots": st["n_slots"], "n_graphed": st["n_graphed"], "size": idx.size()})
        ck_windows.append([ck0, round(time.perf_counter() - t_begin, 3)])
    mt.join()
    t_mut = time.perf_counter() - t_begin
    stop.set()
    qt.join()
    idx.insert_pending()
    st = idx.stats()
    checkpoints.append({"t_s": round(time.perf_counter() - t_begin, 1), "recall_at_10": round(recall_now(), 4),
                        "n_slots": st["n_slots"], "n_graphed": st["n_graphed"], "size": idx.size()})
    bs = idx.build_stats()
    # batch-1 latency inside the slowest mutator calls (K7 streaming insert + refinement pass run there)
    slow = sorted(mut_calls, key=lambda w: w[0] - w[1])[:max(1, len(mut_calls) // 100)]
    lt = np.array(lat1_t)
    la = np.array(lat1)
    inside = np.concatenate([la[(lt > s) & (lt < e)] for s, e in slow]) if len(la) else np.array([])
    # timeline of the slow batch-1 queries: when they ended, how long they took, and the mutator call (if any) they ended in
    slow_q = []
    for i in np.argsort(-la)[:40]:
        if la[i] < 0.003:
            break
        inside_call = next(((s, e) for s, e in mut_calls if s < lt[i] < e + la[i]), None)
        slow_q.append({"end_s": round(float(lt[i] - t_begin), 4), "ms": round(float(la[i] * 1e3), 2),
                       "mut_call_ms": round((inside_call[1] - inside_call[0]) * 1e3, 1) if inside_call else None,
                       "offset_in_call_ms": round((lt[i] - la[i] - inside_call[0]) * 1e3, 1) if inside_call else None})
    slow_q.sort(key=lambda d: d["end_s"])
    out = {
        "config": f"C5: {n0}x{dim} f32 cosine (bf16 traversal), {a.rate} mutations/s (70/20/10 insert/delete/update) for "
                  f"{a.seconds:.0f} s in batches of {a.mut_batch}, concurrent batch-1 queries + a 1000-query batch every {a.big_every} s",
        "expansion_search": ef_used, "recall_at_10_fresh_build": round(r0, 4), "build_s": round(t_build, 2),
        "mutations": n_batches * a.mut_batch, "mutation_rate_achieved": n_batches * a.mut_batch / t_mut,
        "mutator_max_lag_s": round(behind, 3),
        "mutator_call_ms": {"p50": float(np.percentile([e - s for s, e in mut_calls], 50) * 1e3),
                            "p99": float(np.percentile([e - s for s, e in mut_calls], 99) * 1e3),
                            "max": float(max(e - s for s, e in mut_calls) * 1e3)},
        "batch1_queries": len(la),
        "batch1_ms": {"p50": float(np.percentile(la, 50) * 1e3), "p99": float(np.percentile(la, 99) * 1e3),
                      "max": float(la.max() * 1e3)},
        "batch1_ms_inside_slowest_mutator_calls": ({"n": int(len(inside)), "p50": float(np.percentile(inside, 50) * 1e3),
                                                    "p99": float(np.percentile(inside, 99) * 1e3)} if len(inside) else None),
        "batch1000_ms": {"n": len(lat_big), "p50": float(np.percentile(lat_big, 50) * 1e3) if lat_big else None,
                         "p99": float(np.percentile(lat_big, 99) * 1e3) if lat_big else None},
        "checkpoints": checkpoints, "checkpoint_windows_s": ck_windows, "slow_batch1_queries": slow_q,
        "slow_mutator_calls": [{"start_s": round(s - t_begin, 3), "ms": round((e - s) * 1e3, 1)} for s, e in sorted(slow)],
        "refine_rows": bs["refine_rows"], "stream_rows": bs["stream_rows"], "compact_ms": bs["compact_ns"] / 1e6,
        "errors": errors,
    }
    print(json.dumps(out), flush=True)
    idx.close()
    return 1 if errors else 0


if __name__ == "__main__":
    sys.exit(main())
