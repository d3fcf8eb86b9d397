import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, synth
H, W = 480, 752
a, b = synth.shifted_pair(H, W, 1000)
ex = ORBExtractor(2000, 1.2, 8, 20, 7)
ka, da = ex(a); kb, db = ex(b)
f1, f2 = FrameView(ka, da, W, H), FrameView(kb, db, W, H)
m = ORBMatcher(0.9, True, handle=ex._h)
pre0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
for _ in range(3): m.SearchForInitialization(f1, f2, pre0.copy(), 100)
t0 = time.perf_counter()
for _ in range(50): n, _ = m.SearchForInitialization(f1, f2, pre0.copy(), 100)
print("init search: %.3f ms per call, %d matches" % ((time.perf_counter() - t0) * 1e3 / 50, n))
