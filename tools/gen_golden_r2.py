#!/usr/bin/env python3
"""Round-2 golden fixtures (run in the build container: needs cv2, scikit-learn's sample photographs and /root/reference):
  photos_ref.npz     — two real photographs (scikit-learn's china.jpg / flower.jpg, gray, resized to the EuRoC and KITTI frame shapes)
                       and, for each, the output of the reference's own ORBExtractor.cpp compiled verbatim (canonical tie-break)
  matcher_c3_ref.npz — BASELINE config 3: SearchForInitialization(window 100, ratio 0.9) of the reference's own ORBMatcher.cpp
                       (oracle/_ref/libref_matcher.so) between two 1920x1080 frames extracted with the 8000-feature initial
                       extractor (Tracking.cpp:24, 606; test/ORB/initializeSearchTest.cpp).  Only the outputs and a digest of the
                       inputs are stored: the inputs are re-created by the oracle extractor from synth.shifted_pair(1080, 1920, 3003).
"""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from sklearn.datasets import load_sample_image
from monoorbslam3_b200 import synth
from oracle import orb_oracle as orc, ref_matcher as ref

OUT = os.path.join(ROOT, "tests", "golden")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.array(h.hexdigest())


def photos():
    d = {}
    for name, src, (w, h), nf in (("china", "china.jpg", (752, 480), 1000), ("flower", "flower.jpg", (1241, 376), 2000)):
        rgb = load_sample_image(src)
        gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        img = np.ascontiguousarray(cv2.resize(gray, (w, h), interpolation=cv2.INTER_AREA if w < gray.shape[1] else cv2.INTER_CUBIC))
        kps, desc = orc.ReferenceExtractor(nf, 1.2, 8, 20, 7, canonical=True)(img)
        d["img_" + name] = img; d["nf_" + name] = np.array(nf); d["kps_" + name] = kps; d["desc_" + name] = desc
        print(name, img.shape, len(kps))
    np.savez_compressed(os.path.join(OUT, "photos_ref.npz"), **d)


def matcher_c3():
    orc.build()
    a, b = synth.shifted_pair(1080, 1920, 3003)
    ex = orc.Extractor(8000, 1.2, 8, 20, 7)
    ka, da = ex(a); kb, db = ex(b)
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    n, m12, pre2 = ref.search_for_initialization(ka, da, kb, db, 1920, 1080, pre, 100, 0.9, True)
    print("C3 init search:", len(ka), len(kb), "key points,", n, "matches")
    np.savez_compressed(os.path.join(OUT, "matcher_c3_ref.npz"), seed=3003, n=n, m12=m12, pre=pre2, inputs=digest(ka, da, kb, db),
                        n1=len(ka), n2=len(kb))


if __name__ == "__main__":
    photos()
    matcher_c3()
