#!/usr/bin/env python3
"""Golden vectors of the reference's own modules/ORB/ORBMatcher.cpp (compiled verbatim into oracle/_ref/libref_matcher.so, see
oracle/Makefile and oracle/matcher_harness.cpp): inputs and outputs of every matcher entry point on one synthetic frame pair.
Writes tests/golden/matcher_ref.npz.  Run in the build container (the library needs /root/reference to build)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orb_oracle as orc, ref_matcher as ref
from monoorbslam3_b200 import synth

W, H = 752, 480
orc.build()
a, b = synth.shifted_pair(H, W, 2024)
ex = orc.Extractor(1200, 1.2, 8, 20, 7)
ka, da = ex(a); kb, db = ex(b)
rng = np.random.default_rng(5)
db = db.copy(); dup = rng.integers(0, len(db), 50); db[dup] = db[(dup + 1) % len(db)]
sf = np.array([ex.scale(l) for l in range(8)], np.float32)
out = dict(ka=ka, da=da, kb=kb, db=db, sf=sf, w=W, h=H)

def fv(desc, bits):
    node = desc[:, 0].astype(np.int32) >> (8 - bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)

pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
for tag, (window, ratio, orient) in dict(init_a=(100, 0.9, 1), init_b=(40, 0.7, 0)).items():
    n, m12, pre2 = ref.search_for_initialization(ka, da, kb, db, W, H, pre, window, ratio, bool(orient))
    out.update({tag + "/args": np.array([window, ratio, orient], np.float64), tag + "/n": n, tag + "/m12": m12, tag + "/pre": pre2})
    print(tag, n)

nq = len(ka)
q_u = (ka["x"] + 7 + rng.normal(0, 1.0, nq)).astype(np.float32); q_v = (ka["y"] + 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
q_valid = (rng.random(nq) < 0.8).astype(np.uint8); occ = (rng.random(len(kb)) < 0.1).astype(np.uint8)
q_l = ka["octave"].astype(np.int32); q_a = ka["angle"].astype(np.float32)
out.update(q_u=q_u, q_v=q_v, q_valid=q_valid, occ=occ)
for tag, (th, orient, from_kf) in dict(proj_frame=(15, 1, 0), proj_kf=(30, 1, 1), proj_noorient=(15, 0, 0)).items():
    q_r = (np.float32(th) * ka["size"]).astype(np.float32)
    n, asg = ref.search_by_projection(q_u, q_v, q_r, q_l, q_a, da, q_valid, kb, db, W, H, occ, bool(orient), bool(from_kf))
    out.update({tag + "/args": np.array([th, orient, from_kf]), tag + "/n": n, tag + "/assigned": asg})
    print(tag, n)
for tag, (th, ratio) in dict(local_a=(2, 0.8), local_b=(15, 0.6)).items():
    vc = np.where(rng.random(nq) < 0.5, np.float32(0.9995), np.float32(0.9)).astype(np.float32)
    n, asg = ref.search_local_points(q_u, q_v, vc, q_l, da, q_valid, th, kb, db, W, H, occ, ratio)
    out.update({tag + "/args": np.array([th, ratio], np.float64), tag + "/view_cos": vc, tag + "/n": n, tag + "/assigned": asg})
    print(tag, n)
has1 = (rng.random(len(da)) < 0.3).astype(np.uint8); has2 = (rng.random(len(db)) < 0.3).astype(np.uint8)
valid1 = (rng.random(len(da)) < 0.7).astype(np.uint8); occ2 = (rng.random(len(db)) < 0.2).astype(np.uint8)
out.update(has1=has1, has2=has2, valid1=valid1, occ2=occ2)
for tag, (bits, orient) in dict(tri_a=(4, 0), tri_b=(3, 1)).items():
    n, m12 = ref.search_for_triangulation(da, ka["angle"], has1, fv(da, bits), db, kb["angle"], has2, fv(db, bits), bool(orient))
    out.update({tag + "/args": np.array([bits, orient]), tag + "/n": n, tag + "/m12": m12})
    print(tag, n)
for tag, (bits, ratio, orient) in dict(bow_a=(3, 0.7, 1), bow_b=(5, 0.9, 0)).items():
    n, asg = ref.search_by_bow(da, ka["angle"], valid1, fv(da, bits), db, kb["angle"], occ2, fv(db, bits), ratio, bool(orient))
    out.update({tag + "/args": np.array([bits, ratio, orient], np.float64), tag + "/n": n, tag + "/assigned": asg})
    print(tag, n)
nf = 1200
src = rng.integers(0, len(kb), nf)
f_u = (kb["x"][src] + 7 + rng.normal(0, 1.0, nf)).astype(np.float32); f_v = (kb["y"][src] + 3 + rng.normal(0, 1.0, nf)).astype(np.float32)
f_l = np.clip(kb["octave"][src] + rng.integers(-1, 2, nf), 0, 7).astype(np.int32); f_valid = (rng.random(nf) < 0.9).astype(np.uint8)
out.update(f_u=f_u, f_v=f_v, f_l=f_l, f_valid=f_valid, f_src=src.astype(np.int32))
for tag, th in dict(fuse_a=3.0, fuse_b=5.0).items():
    n, bi = ref.search_fuse(f_u, f_v, f_l, db[src], f_valid, th, ka, da, W, H)
    out.update({tag + "/args": np.array([th]), tag + "/n": n, tag + "/best_idx": bi})
    print(tag, n)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "matcher_ref.npz"), **out)
print("wrote tests/golden/matcher_ref.npz")
