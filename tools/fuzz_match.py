#!/usr/bin/env python3
"""Randomised parity run of the matchers (window searches with random radii / occupancy / validity, node searches with random
vocabulary groupings, fuse search, computeDescriptor) against the oracle's sequential restatements.  usage: fuzz_match.py [n] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, synth
from oracle import orb_oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
H, W = 480, 752
ex = ORBExtractor(1500, 1.2, 8, 20, 7)
pairs = []
for seed in (1000, 1001):
    a, b = synth.shifted_pair(H, W, seed)
    pairs.append((ex(a), ex(b)))


def fv_random(desc, n_nodes, r):
    node = r.integers(0, n_nodes, len(desc)) * 7 + 3                  # arbitrary ascending ids with gaps
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


bad = 0
t0 = time.time()
for case in range(n_cases):
    (ka, da), (kb, db) = pairs[case % 2]
    sub = rng.random(len(ka)) < rng.uniform(0.2, 1.0)                  # random subset of queries
    kq, dq = ka[sub], da[sub]
    nq = len(kq)
    jitter = rng.uniform(0.5, 12.0)
    q_u = (kq["x"] - 7 + rng.normal(0, jitter, nq)).astype(np.float32); q_v = (kq["y"] - 3 + rng.normal(0, jitter, nq)).astype(np.float32)
    q_l = np.clip(kq["octave"] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32); q_a = kq["angle"].astype(np.float32)
    q_valid = (rng.random(nq) < rng.uniform(0.3, 1.0)).astype(np.uint8); occ = (rng.random(len(kb)) < rng.uniform(0, 0.5)).astype(np.uint8)
    th = float(rng.choice([1, 2, 7, 15, 30, 60])); ratio = float(rng.choice([0.6, 0.8, 0.9, 1.0])); orient = bool(rng.random() < 0.7)
    q_r = (np.float32(th) * kq["size"] * np.float32(rng.choice([1.0, 2.5, 4.0]))).astype(np.float32)
    cur = FrameView(kb, db, W, H)
    m = ORBMatcher(ratio, orient, handle=ex._h)
    res = []
    n, asg = m.SearchByProjection(q_u, q_v, q_r, q_l, q_a, dq, q_valid, cur, occ)
    on, oasg = orc.search_by_projection(q_u, q_v, q_r, q_l, q_a, dq, q_valid, kb, db, W, H, occ, orient)
    res.append(("proj", n == on and np.array_equal(asg, oasg)))
    n, asg = m.SearchLocalPoints(q_u, q_v, q_r, q_l, dq, q_valid, cur, occ)
    on, oasg = orc.search_local_points(q_u, q_v, q_r, q_l, dq, q_valid, kb, db, W, H, occ, ratio)
    res.append(("local", n == on and np.array_equal(asg, oasg)))
    nn = int(rng.choice([1, 3, 9, 40, 200]))
    fv1, fv2 = fv_random(dq, nn, rng), fv_random(db, nn, rng)
    flag1 = (rng.random(nq) < 0.6).astype(np.uint8)
    n, asg = m.SearchByBow(dq, kq["angle"], flag1, fv1, db, kb["angle"], occ, fv2)
    on, oasg = orc.search_by_bow(dq, kq["angle"], flag1, fv1, db, kb["angle"], occ, fv2, ratio, orient)
    res.append(("bow", n == on and np.array_equal(asg, oasg)))
    n, m12 = m.SearchForTriangulation(dq, kq["angle"], flag1, fv1, db, kb["angle"], occ, fv2)
    on, om12 = orc.search_for_triangulation(dq, kq["angle"], flag1, fv1, db, kb["angle"], occ, fv2, orient)
    res.append(("tri", n == on and np.array_equal(m12, om12)))
    pre = np.stack([q_u, q_v], 1).astype(np.float32); opre = pre.copy()
    win = int(rng.choice([30, 100, 200]))
    n, m12 = m.SearchForInitialization(FrameView(kq, dq, W, H), cur, pre, win)
    on, om12, opre = orc.search_for_initialization(kq, dq, kb, db, W, H, opre, win, ratio, orient)
    res.append(("init", n == on and np.array_equal(m12, om12) and np.array_equal(pre, opre)))
    fails = [k for k, ok in res if not ok]
    print("case %d: nq %d th %.0f ratio %.1f orient %s nodes %d ->" % (case, nq, th, ratio, orient, nn), "ok" if not fails else "MISMATCH " + ",".join(fails))
    bad += len(fails)
print("%d cases, %d mismatches, %.0f s" % (n_cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)
