#!/usr/bin/env python3
"""One device-resident pass over 64 variants of the recorded photograph (tests/golden/photos_ref.npz), for ncu captures."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from monoorbslam3_b200 import ORBExtractor

img = np.load(os.path.join(ROOT, "tests", "golden", "photos_ref.npz"))["img_china"]
h, w = img.shape
base = np.ascontiguousarray(np.stack([img, img[::-1], img[:, ::-1], img[::-1, ::-1], np.roll(img, 37, 1), np.roll(img, 53, 0), np.roll(img[::-1], 91, 1), np.roll(img[:, ::-1], 17, 0)]))
batch = 32
dev = torch.device("cuda", 0)
fr = torch.from_numpy(np.concatenate([base] * (batch // 8))).to(dev)
ex = ORBExtractor(1000, 1.2, 8, 20, 7, max_batch=batch)
cap = 1128
kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(2): ex.extract_batch_device(fr, batch, h, w, kps, desc, cap, n, stream=s.cuda_stream, sync=True)
print("key points / frame:", float(n.float().mean()))
