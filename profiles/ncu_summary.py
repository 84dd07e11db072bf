#!/usr/bin/env python
"""Summarise an ncu report: `python profiles/ncu_summary.py <file.ncu-rep> [regex]` prints the
metrics the roofline discussion needs (duration, DRAM bytes/throughput, tensor pipe, occupancy,
issue activity, top warp-stall reasons) for each captured launch."""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
            "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        print(f"== {name[:100]}  (ID {r[0]})")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 3 and h.split(".")[1] in ("TriageCompute",) else h
            if h in keys or short in keys:
                print(f"   {h} [{units[i]}] = {r[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m and r[i]:
                try:
                    stalls.append((float(r[i].replace(",", "")), m.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stall reasons (warps per issue-active):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
