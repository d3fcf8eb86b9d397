#!/usr/bin/env python3
"""Print the handful of ncu raw-page metrics used in profiles/ for every kernel of an .ncu-rep."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:60])
    for k in KEYS:
        if k in h: print("   %-70s %s" % (k, r[h.index(k)]))
    for i, k in enumerate(h):
        if "issue_stalled" in k and k.endswith("_per_warp_active.pct") and float(r[i] or 0) > 4: print("   %-70s %s" % (k, r[i]))
