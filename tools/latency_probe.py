#!/usr/bin/env python3
"""Single-frame latency: wall time of ORBExtractor.__call__ and the per-stage device times of a 1-frame pass."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, synth

for (w, h, nf) in ((752, 480, 1000), (1241, 376, 2000), (1920, 1080, 4000)):
    img = synth.frame(h, w, 1000, "dense")
    ex = ORBExtractor(nf, 1.2, 8, 20, 7)
    for _ in range(20): ex(img)
    t0 = time.perf_counter()
    for _ in range(200): kps, desc = ex(img)
    dt = (time.perf_counter() - t0) / 200
    ex.profile(True); ex.profile_read(reset=True)
    for _ in range(50): ex(img)
    st, passes = ex.profile_read(reset=True)
    ex.profile(False)
    print("%dx%d nf=%d: %.3f ms per call (%d kps); device stages per pass (ms): %s  sum %.3f" %
          (w, h, nf, dt * 1e3, len(kps), {k: round(v / max(passes, 1), 4) for k, v in st.items()}, sum(st.values()) / max(passes, 1)))
    ex.close()
