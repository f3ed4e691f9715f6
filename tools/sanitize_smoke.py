#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer (SURVEY §5 "race detection"):

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

add (K0) -> exact search (K1-TC / K1-SIMT / K3, certified f32 path) -> build (all-pairs K1-TC, K6) -> ANN batch (seed
tiles, K4, K3 re-rank) -> batch-1 (seed scan, K4b) -> filtered ANN (K4 with a bitmap) and filtered exact -> streaming
insert (K7) -> remove + compaction -> snapshot round trip.  Sizes are tiny: the sanitizer slows kernels 10-100x."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vector_store_b200 as v  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    n, dim, k = 12_000, 96, 10
    centers = rng.standard_normal((24, dim)).astype(np.float32)
    x = centers[rng.integers(0, 24, n)] + 0.3 * rng.standard_normal((n, dim)).astype(np.float32)
    q = centers[rng.integers(0, 24, 300)] + 0.3 * rng.standard_normal((300, dim)).astype(np.float32)
    keys = np.arange(n, dtype=np.uint64)
    for storage, flags in ((v.Scalar.F32, dict(bf16_traversal=True)), (v.Scalar.BF16, {}), (v.Scalar.I8, {})):
        idx = v.GpuIndex(dim, v.Metric.Cos, storage, device=0, **flags)
        idx.reserve(n + 6000)
        idx.add_batch(keys[:9000], x[:9000])
        idx.search_batch(q, k, exact=True)
        idx.build()
        idx.search_batch(q, k)
        idx.search_batch(q[:1], k)
        mask = np.zeros(n + 6000, dtype=bool)
        mask[::3] = True
        idx.search_filtered(q[:64], k, mask)
        mask[:] = False
        mask[:50] = True
        idx.search_filtered(q[:8], k, mask)
        idx.add_batch(keys[9000:], x[9000:])          # tail + K7 once stream_threshold rows are pending
        idx.insert_pending()
        idx.search_batch(q, k)
        idx.remove_batch(keys[:4000])
        idx.add_batch(keys[:3000] | np.uint64(1 << 48), x[:3000])   # forces the slot-reclaiming compaction
        idx.search_batch(q, k)
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "snap.vsb")
            idx.save(p)
            idx2 = v.GpuIndex.load(p, device=0)
            idx2.search_batch(q[:32], k)
            idx2.close()
        st = idx.stats()
        print(f"storage {storage.name}: launches {st['kernel_launches']}, tc launches {st['tc_launches']}, size {idx.size()}")
        idx.close()
    print("sanitize_smoke done")


if __name__ == "__main__":
    main()
