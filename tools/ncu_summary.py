#!/usr/bin/env python
"""Turns an .ncu-rep (captured on the GPU box, read here without a GPU) into a small text summary for
profiles/: the headline metrics of each captured launch plus the hottest SASS lines by stall samples."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none; cold-cache, serialised)")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')][:110]}")
        for w in WANT:
            if w in hdr:
                print(f"{w:95s} {r[hdr.index(w)]:>20s} {units[hdr.index(w)]}")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        si, so, ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
        data = [(int(r[si] or 0), r[so].strip(), int(r[ex] or 0)) for r in src[2:]
                if len(r) > max(si, so, ex) and (r[si] or '0').isdigit()]
        tot = sum(d[0] for d in data) or 1
        print(f"\n## hottest SASS lines of the first captured launch (of {tot} stall samples)")
        for s, t, e in sorted(data, key=lambda d: -d[0])[:top]:
            print(f"{s / tot * 100:5.1f}%  executed={e:12d}  {t[:100]}")


if __name__ == "__main__":
    main()
