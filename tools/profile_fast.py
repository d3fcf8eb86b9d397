#!/usr/bin/env python3
"""One device-resident extractor pass for ncu captures of the FAST kernel: profile_fast.py <w> <h> <n_features> <batch> [profile]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, synth

w, h, nf, batch = (int(v) for v in sys.argv[1:5])
profile = sys.argv[5] if len(sys.argv) > 5 else "dense"
dev = torch.device("cuda", 0)
base = synth.frames(min(batch, 8), h, w, 1000, profile)
fr = torch.from_numpy(np.concatenate([base] * ((batch + len(base) - 1) // len(base)))[:batch]).to(dev)
ex = ORBExtractor(nf, 1.2, 8, 20, 7, max_batch=batch)
cap = nf + 128
kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(2): ex.extract_batch_device(fr, batch, h, w, kps, desc, cap, n, stream=s.cuda_stream, sync=True)
print("key points / frame:", float(n.float().mean()))
