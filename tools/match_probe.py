#!/usr/bin/env python3
"""Latency of the tracking-loop matcher calls (one call = what Tracking issues per frame): GPU C-ABI call vs the CPU oracle port."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, synth
from oracle import orb_oracle as orc

H, W = 480, 752
a, b = synth.shifted_pair(H, W, 1000)
ex = ORBExtractor(1500, 1.2, 8, 20, 7)
ka, da = ex(a); kb, db = ex(b)
rng = np.random.default_rng(0)
n = len(ka)
q_u = (ka["x"] - 7 + rng.normal(0, 1.0, n)).astype(np.float32); q_v = (ka["y"] - 3 + rng.normal(0, 1.0, n)).astype(np.float32)
q_l = ka["octave"].astype(np.int32); q_a = ka["angle"].astype(np.float32); q_valid = (rng.random(n) < 0.9).astype(np.uint8)
occ = np.zeros(len(kb), np.uint8)
cur = FrameView(kb, db, W, H)


def bench(f, reps=30):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    return (time.perf_counter() - t0) / reps * 1e3, r


for th in (15, 30):
    q_r = (np.float32(th) * ka["size"]).astype(np.float32)
    m = ORBMatcher(0.9, True, handle=ex._h)
    g, (gn, _) = bench(lambda: m.SearchByProjection(q_u, q_v, q_r, q_l, q_a, da, q_valid, cur, occ))
    c, (cn, _) = bench(lambda: orc.search_by_projection(q_u, q_v, q_r, q_l, q_a, da, q_valid, kb, db, W, H, occ, True), 5)
    print("SearchByProjection th=%d: %d queries, GPU %.3f ms (%d matches), CPU port %.3f ms (%d)" % (th, n, g, gn, c, cn))
sf = np.array([ex.getScaleFactor(int(l)) for l in q_l], np.float32)
for th in (1, 2):
    q_r = (np.float32(th) * np.float32(4.0) * sf).astype(np.float32)
    m = ORBMatcher(0.8, True, handle=ex._h)
    g, (gn, _) = bench(lambda: m.SearchLocalPoints(q_u, q_v, q_r, q_l, da, q_valid, cur, occ))
    c, (cn, _) = bench(lambda: orc.search_local_points(q_u, q_v, q_r, q_l, da, q_valid, kb, db, W, H, occ, 0.8), 5)
    print("SearchLocalPoints th=%d: GPU %.3f ms (%d matches), CPU port %.3f ms (%d)" % (th, g, gn, c, cn))
