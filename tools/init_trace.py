#!/usr/bin/env python3
"""SearchForInitialization latency split: Python wrapper, C-ABI call, device work (ORBFE_TRACE=1 prints the C side)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, synth
H, W = 480, 752
fa, fb = synth.shifted_pair(H, W, 1000)
ex = ORBExtractor(2000, 1.2, 8, 20, 7)
ka, da = ex(fa); kb, db = ex(fb)
f1, f2 = FrameView(ka, da, W, H), FrameView(kb, db, W, H)
mi = ORBMatcher(0.9, True, handle=ex._h)
pre0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
for _ in range(3): mi.SearchForInitialization(f1, f2, pre0.copy(), 100)
t0 = time.perf_counter()
for _ in range(20): n, _ = mi.SearchForInitialization(f1, f2, pre0.copy(), 100)
print("python call: %.3f ms, matches %d" % ((time.perf_counter() - t0) / 20 * 1e3, n))
