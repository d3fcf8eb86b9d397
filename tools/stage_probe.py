#!/usr/bin/env python3
"""Per-stage device times for any BASELINE config: stage_probe.py <w> <h> <n_features> <batch> [dense|natural|photo]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, synth

w, h, nf, batch = (int(v) for v in sys.argv[1:5])
profile = sys.argv[5] if len(sys.argv) > 5 else "dense"
dev = torch.device("cuda", 0)
if profile == "photo":      # the recorded 752x480 photograph of tests/golden/photos_ref.npz and seven flipped / shifted variants of it
    img = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "photos_ref.npz"))["img_china"]
    assert img.shape == (h, w), img.shape
    base = np.stack([img, img[::-1], img[:, ::-1], img[::-1, ::-1], np.roll(img, 37, 1), np.roll(img, 53, 0), np.roll(img[::-1], 91, 1), np.roll(img[:, ::-1], 17, 0)])[:min(batch, 8)]
    base = np.ascontiguousarray(base)
else:
    base = synth.frames(min(batch, 8), h, w, 1000, profile)
fr = torch.from_numpy(np.concatenate([base] * ((batch + len(base) - 1) // len(base)))[:batch]).to(dev)
ex = ORBExtractor(nf, 1.2, 8, 20, 7, max_batch=batch)
cap = nf + 128
kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(3): ex.extract_batch_device(fr, batch, h, w, kps, desc, cap, n, stream=s.cuda_stream, sync=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ex.extract_batch_device(fr, batch, h, w, kps, desc, cap, n, stream=s.cuda_stream, sync=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ex.profile(True); ex.profile_read(reset=True)
for _ in range(5): ex.extract_batch_device(fr, batch, h, w, kps, desc, cap, n, stream=s.cuda_stream, sync=True)
st, passes = ex.profile_read(reset=True)
print("%dx%d nf=%d batch=%d %s: %.3f ms/pass = %.0f frames/s, %.1f kps/frame; stages (ms): %s" %
      (w, h, nf, batch, profile, ms, batch / ms * 1e3, float(n.float().mean()), {k: round(v / passes, 3) for k, v in st.items()}))
