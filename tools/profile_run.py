#!/usr/bin/env python3
"""Short extractor run for ncu: `passes` device-resident passes over `batch` C1 frames (same kernels as bench.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, synth

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W, NF = 480, 752, 1000
dev = torch.device("cuda", 0)
base = synth.frames(min(batch, 16), H, W, 1000, "dense")
fr = torch.from_numpy(np.concatenate([base] * ((batch + len(base) - 1) // len(base)))[:batch]).to(dev)
ex = ORBExtractor(NF, 1.2, 8, 20, 7, max_batch=batch)
cap = NF + 64
kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(passes):
    ex.extract_batch_device(fr, batch, H, W, kps, desc, cap, n, stream=s.cuda_stream, sync=True)
print("key points / frame:", float(n.float().mean()))
if "--match" in sys.argv:
    nq = 8192
    d = torch.randint(0, 256, (nq, 32), dtype=torch.uint8).to(dev)
    bi = torch.zeros(nq, dtype=torch.int32, device=dev); bd = torch.zeros_like(bi); sd = torch.zeros_like(bi)
    ORBMatcher(handle=ex._h).hamming_allpairs_device(d, nq, d, nq, bi, bd, sd, stream=s.cuda_stream, sync=True)
    print("match ok", int(bd.sum()))
