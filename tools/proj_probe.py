import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, DeviceFrame, synth
H, W = 480, 752
a, b = synth.shifted_pair(H, W, 1000)
ex = ORBExtractor(2000, 1.2, 8, 20, 7)
ka, da = ex(a); kb, db = ex(b)
rng = np.random.default_rng(0); n = len(ka)
q_u = (ka["x"] - 7 + rng.normal(0, 1.0, n)).astype(np.float32); q_v = (ka["y"] - 3 + rng.normal(0, 1.0, n)).astype(np.float32)
q_valid = (rng.random(n) < 0.9).astype(np.uint8); occ = np.zeros(len(kb), np.uint8)
df2 = DeviceFrame.upload(kb, db, W, H, handle=ex._h)
m = ORBMatcher(0.9, True, handle=ex._h)
q_r = (np.float32(15) * ka["size"]).astype(np.float32)
for _ in range(12):
    m.SearchByProjection(q_u, q_v, q_r, ka["octave"], ka["angle"], da, q_valid, df2, occ)
import time
for name, fn in (("SearchByProjection th 15", lambda: m.SearchByProjection(q_u, q_v, q_r, ka["octave"], ka["angle"], da, q_valid, df2, occ)),
                 ("SearchLocalPoints th 2", lambda: ORBMatcher(0.8, True, handle=ex._h).SearchLocalPoints(q_u, q_v, (np.float32(2 * 2.5) * ka["size"] / np.float32(31.0) * np.float32(1.0)).astype(np.float32) if False else (np.float32(2.0) * np.float32(2.5) * np.float32(1.2) ** ka["octave"]).astype(np.float32), ka["octave"], da, q_valid, df2, occ))):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(300): r = fn()
    print("%s: %.4f ms per call on a device-resident frame (%d queries, %d matches)" % (name, (time.perf_counter() - t0) / 300 * 1e3, n, r[0]))
