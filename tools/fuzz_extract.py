#!/usr/bin/env python3
"""Randomised parity run: random image sizes / constructor arguments / image statistics through the CUDA extractor (TMA and
vector-load staging, single-frame and batch entry points) against the CPU oracle.  usage: fuzz_extract.py [n_cases] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, OrbfeError, synth
from oracle import orb_oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def image(h, w, kind, seed):
    r = np.random.default_rng(seed)
    if kind == "dense": return synth.frame(h, w, seed, "dense")
    if kind == "natural": return synth.frame(h, w, seed, "natural")
    if kind == "noise": return r.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "binary": return np.where(synth.frame(h, w, seed, "dense") > 128, 255, 0).astype(np.uint8)
    if kind == "blocks":
        img = np.kron(r.integers(0, 256, ((h + 15) // 16, (w + 15) // 16)), np.ones((16, 16)))[:h, :w]
        return np.clip(img + r.normal(0, 2, (h, w)), 0, 255).astype(np.uint8)
    img = np.full((h, w), 100, np.uint8); img[h // 3:h // 3 + 40, w // 4:w // 4 + 60] = synth.frame(40, 60, seed, "dense")   # sparse
    return img


bad = 0
t0 = time.time()
for case in range(n_cases):
    w = int(rng.integers(150, 1300)); h = int(rng.integers(140, 800))
    nf = int(rng.choice([50, 300, 1000, 2500])); sf = float(rng.choice([1.1, 1.2, 1.2, 1.3, 1.5, 2.0])); nl = int(rng.integers(1, 10))
    ini = int(rng.choice([20, 20, 12, 40, 130])); mn = int(rng.choice([7, 7, 5, 20]))
    kind = str(rng.choice(["dense", "natural", "noise", "binary", "blocks", "sparse"]))
    tma = bool(rng.random() < 0.7)
    img = image(h, w, kind, 5000 + case)
    tag = "case %d: %dx%d nf=%d sf=%.1f levels=%d th=%d/%d %s tma=%s" % (case, w, h, nf, sf, nl, ini, mn, kind, tma)
    try:
        ex = ORBExtractor(nf, sf, nl, ini, mn, use_tma=tma, max_batch=3)
    except OrbfeError as e:
        print(tag, "-> create rejected:", e); continue
    try:
        kps, desc = ex(img)
    except OrbfeError as e:
        print(tag, "-> rejected:", str(e)[:100]); ex.close(); continue
    okps, odesc = orc.Extractor(nf, sf, nl, ini, mn)(img)
    ok = len(kps) == len(okps) and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    if ok:   # batch entry point (3 frames: this one twice + a shifted crop of it) must reproduce the single-frame result
        other = np.ascontiguousarray(np.roll(img, 5, axis=1))
        n, bk, bd = ex.extract_batch(np.stack([img, other, img]))
        ok = n[0] == len(kps) and n[2] == len(kps) and bk[0, :n[0]].tobytes() == kps.tobytes() and bk[2, :n[2]].tobytes() == kps.tobytes() \
            and np.array_equal(bd[0, :n[0]], desc) and np.array_equal(bd[2, :n[2]], desc)
        k2, d2 = ex(other)
        ok = ok and n[1] == len(k2) and bk[1, :n[1]].tobytes() == k2.tobytes() and np.array_equal(bd[1, :n[1]], d2)
    print(tag, "->", "ok (%d kps)" % len(kps) if ok else "MISMATCH (%d vs %d kps)" % (len(kps), len(okps)))
    bad += 0 if ok else 1
    ex.close()
print("%d cases, %d mismatches, %.0f s" % (n_cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)
