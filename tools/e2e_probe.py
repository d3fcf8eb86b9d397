#!/usr/bin/env python3
"""e2e pipeline probe: pinned H2D/D2H copy rates and orbfe_extract_batch frames/s for several ORBFE_CHUNK values."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, synth

H, W, NF, B = 480, 752, 1000, 512
dev = torch.device("cuda", 0)
base = synth.frames(16, H, W, 1000, "dense")
host = torch.from_numpy(np.concatenate([base] * (B // 16))).pin_memory()
d = torch.empty_like(host, device=dev)
for _ in range(3): d.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("H2D %d MB: %.3f ms  %.1f GB/s" % (host.numel() >> 20, dt * 1e3, host.numel() / dt / 1e9))
back = torch.empty((B, 1064, 60), dtype=torch.uint8).pin_memory(); dsrc = torch.empty_like(back, device=dev)
for _ in range(3): back.copy_(dsrc, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): back.copy_(dsrc, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("D2H %d MB: %.3f ms  %.1f GB/s" % (back.numel() >> 20, dt * 1e3, back.numel() / dt / 1e9))
from monoorbslam3_b200.extractor import KP_DTYPE
cap = NF + 64
h_n = torch.zeros(B, dtype=torch.int32).pin_memory(); h_kps = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
h_desc = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
out = (h_n.numpy(), h_kps.numpy().view(KP_DTYPE).reshape(B, cap), h_desc.numpy()); fr = host.numpy()
for chunk in sys.argv[1:] or ["128"]:
    os.environ["ORBFE_CHUNK"] = chunk
    ex = ORBExtractor(NF, 1.2, 8, 20, 7, max_batch=B)
    for _ in range(3): ex.extract_batch(fr, cap=cap, out=out)
    t0 = time.perf_counter()
    for _ in range(10): ex.extract_batch(fr, cap=cap, out=out)
    dt = (time.perf_counter() - t0) / 10
    print("chunk %s: %.3f ms/step  %.0f frames/s" % (chunk, dt * 1e3, B / dt))
    ex.close()
