#!/usr/bin/env python
"""What can a beam search at best read?  Random whole-row gathers over a 15 GB buffer (the footprint of the 10 M x 768
bf16 corpus) with nothing but loads (tools/synth/synth.cu::row_gather_kernel), for the row sizes K4 meets: 256 B (C4),
768 B (int8 traversal copy), 1536 B (bf16), 3072 B (f32).  Prints GB/s and rows/s per (row size, rows in flight per
warp); the best line per row size is the ceiling K4's `roofline.frac` should be read against (the streaming-copy peak
in MEASURED_PEAKS.json is only reachable by long contiguous reads)."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module  # noqa: E402

ds = import_module("vector_store_b200.host.datasets")


def main():
    lib = ds.synth_lib()
    lib.vsbsynth_row_gather.restype = C.c_int
    lib.vsbsynth_row_gather.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                        C.c_void_p, C.c_void_p]
    dev = torch.device("cuda", 0)
    total = 15_360_000_000
    buf = torch.empty(total // 4, dtype=torch.int32, device=dev)
    buf.random_(0, 1 << 30)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = []
    for row_bytes in (256, 768, 1536, 3072):
        n_rows = total // row_bytes
        for rif in (2, 4, 8):
            if row_bytes * rif > 3072 * 4:
                continue
            for warps_per_sm in (16, 32, 64):
                warps = 148 * warps_per_sm
                steps = max(8, int(40e9 / (warps * rif * row_bytes)))  # ~40 GB per launch
                best = None
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = lib.vsbsynth_row_gather(buf.data_ptr(), n_rows, row_bytes, rif, warps, steps, 77 + rep, sink.data_ptr(),
                                                 stream or None)
                    assert rc == 0, rc
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    best = ms if best is None else min(best, ms)
                rows = warps * steps * rif
                out.append({"row_bytes": row_bytes, "rows_in_flight_per_warp": rif, "warps_per_sm": warps_per_sm,
                            "gb_per_s": round(rows * row_bytes / best / 1e6, 1), "g_rows_per_s": round(rows / best / 1e6, 3)})
                print(json.dumps(out[-1]), flush=True)
    best = {}
    for o in out:
        if o["row_bytes"] not in best or o["gb_per_s"] > best[o["row_bytes"]]["gb_per_s"]:
            best[o["row_bytes"]] = o
    print(json.dumps({"random_row_gather_ceiling": list(best.values())}))


if __name__ == "__main__":
    main()
