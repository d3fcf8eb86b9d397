#!/usr/bin/env python3
"""Latency of the tracking matchers per call on the bench's frame pair (752x480, 2000-feature initial extractor): host arrays (every call
uploads key points + descriptors and rebuilds the grid) against device-resident frames (orbfe_frame), with ORBFE_TRACE timelines."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, DeviceFrame, synth

H, W = 480, 752
a, b = synth.shifted_pair(H, W, 1000)
ex = ORBExtractor(2000, 1.2, 8, 20, 7)
ka, da = ex(a); kb, db = ex(b)
rng = np.random.default_rng(0)
n = len(ka)
q_u = (ka["x"] - 7 + rng.normal(0, 1.0, n)).astype(np.float32); q_v = (ka["y"] - 3 + rng.normal(0, 1.0, n)).astype(np.float32)
q_l = ka["octave"].astype(np.int32); q_a = ka["angle"].astype(np.float32); q_valid = (rng.random(n) < 0.9).astype(np.uint8)
occ = np.zeros(len(kb), np.uint8)
hf1, hf2 = FrameView(ka, da, W, H), FrameView(kb, db, W, H)
df1, df2 = DeviceFrame.upload(ka, da, W, H, handle=ex._h), DeviceFrame.upload(kb, db, W, H, handle=ex._h)
pre0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)


def bench(f, reps=200):
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    return (time.perf_counter() - t0) / reps * 1e3, r


m = ORBMatcher(0.9, True, handle=ex._h)
for name, f1, f2 in (("host arrays", hf1, hf2), ("device frames", df1, df2)):
    g, (gn, _) = bench(lambda: m.SearchForInitialization(f1, f2, pre0.copy(), 100), 50)
    print("%-14s SearchForInitialization: %.3f ms per call (%d matches, %d level-0 queries)" % (name, g, gn, int((ka["octave"] == 0).sum())))
    q_r = (np.float32(15) * ka["size"]).astype(np.float32)
    g, (gn, _) = bench(lambda: m.SearchByProjection(q_u, q_v, q_r, q_l, q_a, da, q_valid, f2, occ))
    print("%-14s SearchByProjection th 15: %.3f ms per call (%d matches, %d queries)" % (name, g, gn, n))
    sf = np.array([ex.getScaleFactor(int(l)) for l in q_l], np.float32)
    q_r2 = (np.float32(2) * np.float32(4.0) * sf).astype(np.float32)
    m2 = ORBMatcher(0.8, True, handle=ex._h)
    g, (gn, _) = bench(lambda: m2.SearchLocalPoints(q_u, q_v, q_r2, q_l, da, q_valid, f2, occ))
    print("%-14s SearchLocalPoints th 2:   %.3f ms per call (%d matches)" % (name, g, gn))
