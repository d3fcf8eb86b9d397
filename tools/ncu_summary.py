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


if __name__ == "__main__" and sys.argv[1] != "--traffic":
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


def traffic_json(rep, out_json, batch, note="", whole=False):
    """profiles/rNN_traffic.json: per stage of the LAST pass in `rep` (tools/profile_run.py runs the pass twice) the kernel time,
    DRAM bytes, executed warp instructions and issue / pipe utilisation — what bench.py's roofline.traffic and issue figures read."""
    import json
    rows, _ = rows_of(rep)
    stage_of = {"k_resize": "pyramid", "k_fast": "fast", "k_octree": "quadtree", "k_blur": "blur", "k_describe": "describe"}
    seq = [(next((v for k, v in stage_of.items() if k in r["Kernel Name"]), None), r) for r in rows]
    seq = [(s, r) for s, r in seq if s]
    n_desc = sum(1 for s, _ in seq if s == "describe")
    # keep the kernels after the second-to-last describe launch (= the last pass)
    # whole: the capture already holds exactly one pass (e.g. --launch-skip 22 -c 22: the two half passes of a split batch)
    if n_desc > 1 and not whole:
        seen = 0
        for i, (s, _) in enumerate(seq):
            if s == "describe":
                seen += 1
                if seen == n_desc - 1:
                    seq = seq[i + 1:]
                    break
    stages = {}
    for s, r in seq:
        d = stages.setdefault(s, {"duration_ns": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "warp_instructions": 0.0, "launches": 0,
                                  "issue_active_pct": 0.0, "alu_pipe_pct": 0.0})
        t = float(r["gpu__time_duration.sum"])
        d["duration_ns"] += t; d["dram_read_bytes"] += float(r["dram__bytes_read.sum"]); d["dram_write_bytes"] += float(r["dram__bytes_write.sum"])
        d["warp_instructions"] += float(r["smsp__inst_executed.sum"]); d["launches"] += 1
        d["issue_active_pct"] += t * float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"])
        d["alu_pipe_pct"] += t * float(r["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"])
    for d in stages.values():
        d["issue_active_pct"] /= d["duration_ns"]; d["alu_pipe_pct"] /= d["duration_ns"]
    json.dump({"source": "ncu --set full --clock-control none --import-source on, tools/profile_run.py %d 2 (C1 frames), last pass; %s" % (batch, note),
               "batch": batch, "stages": stages}, open(out_json, "w"), indent=1)
    print("wrote", out_json)


if __name__ == "__main__" and sys.argv[1] == "--traffic":
    rest = sys.argv[5:]
    traffic_json(sys.argv[2], sys.argv[3], int(sys.argv[4]), " ".join(a for a in rest if a != "--whole"), whole="--whole" in rest)
