"""Replays tests/golden/matcher_ref.npz (outputs of the reference's own ORBMatcher.cpp compiled verbatim, tools/gen_golden_matcher.py)
through an implementation given as a dict of callables with the flat signatures of oracle/orb_oracle.py; used for the restatement
(CPU suite) and for the CUDA product through the C-ABI (GPU suite)."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(__file__), "golden", "matcher_ref.npz")


def feature_vector(desc, bits):
    node = desc[:, 0].astype(np.int32) >> (8 - bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


def replay(impl):
    """impl: search_for_initialization, search_by_projection, search_local_points, search_for_triangulation, search_by_bow,
    search_fuse with the argument lists of oracle/orb_oracle.py.  Returns the number of cases checked."""
    g = np.load(PATH)
    ka, da, kb, db, sf = g["ka"], g["da"], g["kb"], g["db"], g["sf"]
    W, H = int(g["w"]), int(g["h"])
    checked = 0
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    for tag in ("init_a", "init_b"):
        window, ratio, orient = g[tag + "/args"]
        n, m12, pre2 = impl["search_for_initialization"](ka, da, kb, db, W, H, pre.copy(), int(window), float(ratio), bool(orient))
        assert n == int(g[tag + "/n"]) and np.array_equal(m12, g[tag + "/m12"]) and np.array_equal(pre2, g[tag + "/pre"]), tag
        checked += 1
    q_u, q_v, q_valid, occ = g["q_u"], g["q_v"], g["q_valid"], g["occ"]
    q_l = ka["octave"].astype(np.int32); q_a = ka["angle"].astype(np.float32)
    for tag in ("proj_frame", "proj_kf", "proj_noorient"):
        th, orient, _ = g[tag + "/args"]
        q_r = (np.float32(th) * ka["size"]).astype(np.float32)
        n, asg = impl["search_by_projection"](q_u, q_v, q_r, q_l, q_a, da, q_valid, kb, db, W, H, occ, bool(orient))
        assert n == int(g[tag + "/n"]) and np.array_equal(asg, g[tag + "/assigned"]), tag
        checked += 1
    for tag in ("local_a", "local_b"):
        th, ratio = g[tag + "/args"]
        vc = g[tag + "/view_cos"]
        q_r = ((np.float32(th) * np.where(vc > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)).astype(np.float32) * sf[q_l]).astype(np.float32)
        n, asg = impl["search_local_points"](q_u, q_v, q_r, q_l, da, q_valid, kb, db, W, H, occ, float(ratio))
        assert n == int(g[tag + "/n"]) and np.array_equal(asg, g[tag + "/assigned"]), tag
        checked += 1
    for tag in ("tri_a", "tri_b"):
        bits, orient = g[tag + "/args"]
        n, m12 = impl["search_for_triangulation"](da, ka["angle"], g["has1"], feature_vector(da, int(bits)), db, kb["angle"], g["has2"],
                                                  feature_vector(db, int(bits)), bool(orient))
        assert n == int(g[tag + "/n"]) and np.array_equal(m12, g[tag + "/m12"]), tag
        checked += 1
    for tag in ("bow_a", "bow_b"):
        bits, ratio, orient = g[tag + "/args"]
        n, asg = impl["search_by_bow"](da, ka["angle"], g["valid1"], feature_vector(da, int(bits)), db, kb["angle"], g["occ2"],
                                       feature_vector(db, int(bits)), float(ratio), bool(orient))
        assert n == int(g[tag + "/n"]) and np.array_equal(asg, g[tag + "/assigned"]), tag
        checked += 1
    f_l = g["f_l"]
    for tag in ("fuse_a", "fuse_b"):
        th = float(g[tag + "/args"][0])
        radius = (np.float32(th) * sf[f_l]).astype(np.float32)
        n, bi = impl["search_fuse"](g["f_u"], g["f_v"], radius, f_l, db[g["f_src"]], g["f_valid"], ka, da, W, H, sf * sf)[:2]
        assert n == int(g[tag + "/n"]) and np.array_equal(bi, g[tag + "/best_idx"]), tag
        checked += 1
    return checked
