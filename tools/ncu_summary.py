#!/usr/bin/env python3
"""Summarise .ncu-rep files into the CSV/JSON committed under profiles/ (read here, no GPU needed)."""
import csv, json, subprocess, sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size"]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    res = []
    for r in rows[2:]:
        res.append({hdr[i]: r[i] for i in idx})
    return res, {hdr[i]: units[i] for i in idx}


if __name__ == "__main__":
    out_csv = sys.argv[1]
    allrows, units = [], {}
    for rep in sys.argv[2:]:
        r, units = rows_of(rep)
        allrows += r
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        keys = [k for k in WANT if k in units]
        w.writerow(keys); w.writerow([units[k] for k in keys])
        for r in allrows:
            w.writerow([r[k] for k in keys])
    print("wrote", out_csv, len(allrows), "kernels")
