#!/usr/bin/env python3
"""Per-source-line and per-opcode breakdown of one kernel from an .ncu-rep (source page CSV): where the instructions and stall samples go."""
import csv, subprocess, sys, collections, re

rep, kern = sys.argv[1], sys.argv[2]
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern] + (["--print-source", "cuda,sass"] if view == "cuda" else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) > max(i_ex, i_s) and r[i_ex].isdigit() and r[i_s].isdigit()]
if view == "cuda":      # keep the per-source-line aggregate rows only (first column = line number)
    data = [r for r in data if r[0].isdigit()]
tot = sum(int(r[i_ex]) for r in data); ts = sum(int(r[i_s]) for r in data)
print("warp instructions", tot, "samples", ts, "lines", len(data))
if view == "cuda":
    for r in data:
        if int(r[i_ex]) > 0.004 * tot or int(r[i_s]) > 0.004 * ts:
            print("%6.2f%% inst %6.2f%% smp | %4s %s" % (100 * int(r[i_ex]) / tot, 100 * int(r[i_s]) / ts, r[0], r[i_src].strip()[:130]))
else:
    h = collections.Counter(); hs = collections.Counter()
    for r in data:
        t = r[i_src].split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.split(".")[0]
        h[op] += int(r[i_ex]); hs[op] += int(r[i_s])
    for op, c in h.most_common(30):
        print("%-10s %6.2f%% inst %6.2f%% smp" % (op, 100 * c / tot, 100 * hs[op] / ts))
