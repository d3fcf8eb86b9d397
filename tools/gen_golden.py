#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ (run in the build container, where cv2 4.13.0 and
/root/reference are available).  The fixtures pin the oracle:
  cv2_primitives.npz  — inputs and cv2 outputs of resize / GaussianBlur / FAST / fastAtan2
  extract_ref.npz     — small frames and the output of the reference's own ORBExtractor.cpp compiled verbatim
                        (oracle/_ref/libref_orb_canon.so: canonical tie-break at ORBExtractor.cpp:757)
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from monoorbslam3_b200 import synth
from oracle import orb_oracle as orc

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def primitives():
    rng = np.random.default_rng(42)
    d = {"cv2_version": np.array(cv2.__version__)}
    img = synth.frame(120, 160, 5, "dense")
    d["img"] = img
    # resize: the level chain of a 160x120 frame (sizes from the reference's float32 rule) + an odd pair
    sizes = [(133, 100), (111, 83), (93, 69)]
    src = img
    for i, (w, h) in enumerate(sizes):
        dst = cv2.resize(src, (w, h), interpolation=cv2.INTER_LINEAR)
        d["resize_%d" % i] = dst
        src = dst
    d["resize_sizes"] = np.array(sizes)
    noise = rng.integers(0, 256, (67, 91), dtype=np.uint8)
    d["noise"] = noise
    d["resize_noise"] = cv2.resize(noise, (76, 56), interpolation=cv2.INTER_LINEAR)
    # blur
    d["blur_img"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    d["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # FAST (ordered key points with responses), whole image and a 36x36 cell, thresholds 20 and 7
    for name, im in (("img", img), ("cell", np.ascontiguousarray(img[20:56, 40:76])), ("noise", noise)):
        for t in (20, 7):
            kps = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(im)
            d["fast_%s_%d" % (name, t)] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.int32).reshape(-1, 3)
    # fastAtan2
    yx = np.concatenate([rng.integers(-200000, 200000, (2000, 2)).astype(np.float32),
                         np.array([[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [5, -5]], np.float32)])
    d["atan_yx"] = yx
    d["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), **d)
    print("cv2_primitives.npz", {k: getattr(v, "shape", None) for k, v in d.items()})


def extractor():
    d = {}
    cases = [("a", 240, 320, 300, "dense", 11), ("b", 200, 376, 500, "natural", 12)]
    for name, h, w, nf, prof, seed in cases:
        img = synth.frame(h, w, seed, prof)
        ref = orc.ReferenceExtractor(nf, 1.2, 8, 20, 7, canonical=True)
        kps, desc = ref(img)
        d["img_" + name] = img; d["nf_" + name] = np.array(nf)
        d["kps_" + name] = kps; d["desc_" + name] = desc
        print(name, img.shape, len(kps))
    np.savez_compressed(os.path.join(OUT, "extract_ref.npz"), **d)


if __name__ == "__main__":
    primitives()
    extractor()
