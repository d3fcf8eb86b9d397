#!/usr/bin/env python3
"""Streaming e2e probe: frames/s of orbfe_extract_batch_submit / _wait with two batches in flight, for chunk schedules given as
arguments ("" = the library's default, otherwise an ORBFE_SCHED string such as 128,128,128,128)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, synth
from monoorbslam3_b200.extractor import KP_DTYPE

H, W, NF, B, K = 480, 752, 1000, 512, 20
base = synth.frames(16, H, W, 1000, "dense")
host = torch.from_numpy(np.concatenate([base] * (B // 16))).pin_memory()
cap = NF + 64
keep = []
def outset():
    n = torch.zeros(B, dtype=torch.int32).pin_memory(); k = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
    d = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory(); keep.append((n, k, d))
    return (n.numpy(), k.numpy().view(KP_DTYPE).reshape(B, cap), d.numpy())
outs = (outset(), outset()); fr = host.numpy()
for sched in sys.argv[1:] or [""]:
    if sched: os.environ["ORBFE_SCHED"] = sched
    else: os.environ.pop("ORBFE_SCHED", None)
    ex = ORBExtractor(NF, 1.2, 8, 20, 7, max_batch=B)
    def run():
        prev = None
        for k in range(K):
            t = ex.extract_batch_submit(fr, outs[k & 1], cap=cap)
            if prev is not None: ex.extract_batch_wait(prev)
            prev = t
        ex.extract_batch_wait(prev)
    run()
    t0 = time.perf_counter(); run(); dt = (time.perf_counter() - t0) / K
    print("sched %-28s %.3f ms/step  %.0f frames/s" % (sched or "default", dt * 1e3, B / dt))
    ex.close()
