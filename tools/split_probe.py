#!/usr/bin/env python3
"""Does splitting a resident batch over two handles / streams (so that the latency-bound tail of one half overlaps with the
issue-bound head of the other) beat one pass over the whole batch?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, synth

H, W, NF, B = 480, 752, 1000, 512
dev = torch.device("cuda", 0)
base = synth.frames(16, H, W, 1000, "dense")
fr = torch.from_numpy(np.concatenate([base] * (B // 16))).to(dev)
cap = NF + 64
kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(B, dtype=torch.int32, device=dev)

def bench(parts, K=20):
    exs = [ORBExtractor(NF, 1.2, 8, 20, 7, max_batch=B // parts) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = B // parts
    def run():
        for i, (ex, st) in enumerate(zip(exs, streams)):
            s = slice(i * step, (i + 1) * step)
            ex.extract_batch_device(fr[s], step, H, W, kps[s], desc[s], cap, n[s], stream=st.cuda_stream, sync=False)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams: st.wait_event(e0)
    for _ in range(K): run()
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%d part(s): %.3f ms/step  %.0f frames/s  (n sum %d)" % (parts, ms, B / ms * 1e3, int(n.sum())))
    for ex in exs: ex.close()

for p in (1, 2, 4):
    bench(p)
