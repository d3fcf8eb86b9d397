"""The C++ matcher adapter with the reference's own seven signatures (host/ORBMatcher.h, ORBFE_REFERENCE_TYPES) on Frame / KeyFrame /
MapPoint objects, against the reference's own ORBMatcher.cpp compiled verbatim (oracle/_ref/libref_matcher.so) on the same inputs.
tests/cpp/matcher_ref_test.cpp builds the objects (stand-in types of oracle/matchshim — the headers the verbatim build uses), calls
SearchForInitialization, SearchByProjection x3, SearchForTriangulation, SearchByBow and the static fuse SearchByProjection, and
reports what they did to the objects.  Map points are placed at (u, v, 1) in front of an identity pose and a unit pinhole."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
W, H = 752, 480


def fv(desc, bits):
    node = desc[:, 0].astype(np.int32) >> (8 - bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


class Writer:
    def __init__(self):
        self.parts = []

    def i(self, *v):
        self.parts.append(np.array(v, np.int32).tobytes())

    def f(self, *v):
        self.parts.append(np.array(v, np.float32).tobytes())

    def a(self, arr, dt):
        self.parts.append(np.ascontiguousarray(arr, dtype=dt).tobytes())

    def frame(self, k, d):
        self.i(len(k)); self.parts.append(np.ascontiguousarray(k).tobytes()); self.a(d, np.uint8)

    def fvec(self, f):
        self.i(len(f[0])); self.a(f[0], np.int32); self.a(f[1], np.int32); self.a(f[2], np.int32)


def test_reference_signatures_equal_the_verbatim_matcher(tmp_path, oracle):
    from monoorbslam3_b200 import ORBExtractor, KP_DTYPE, synth, build
    from oracle import ref_matcher as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref_matcher.so is not built")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))
    import build_matcher_ref_test
    exe = build_matcher_ref_test.build()
    ex = ORBExtractor(1200, 1.2, 8, 20, 7)
    a, b = synth.shifted_pair(H, W, 2025)
    ka, da = ex(a); kb, db = ex(b)
    rng = np.random.default_rng(17)
    db = db.copy(); dup = rng.integers(0, len(db), 40); db[dup] = db[(dup + 1) % len(db)]
    n1, n2 = len(ka), len(kb)
    wr = Writer(); exp = {}
    wr.i(W, H)
    # 1. initialization
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    wr.frame(ka, da); wr.frame(kb, db); wr.i(100); wr.f(0.9); wr.i(1); wr.a(pre, np.float32)
    exp["init"] = ref.search_for_initialization(ka, da, kb, db, W, H, pre, 100, 0.9, True)
    # 2. / 3. projection from the last frame / key frame
    for from_kf, th, orient in ((0, 15.0, 1), (1, 30.0, 0)):
        q_u = (ka["x"] + 7 + rng.normal(0, 1.0, n1)).astype(np.float32); q_v = (ka["y"] + 3 + rng.normal(0, 1.0, n1)).astype(np.float32)
        q_u[:15] -= 800; q_v[15:25] += 600                                              # projections outside the image
        state = rng.choice([0, 1, 2, 3], n1, p=[0.15, 0.7, 0.08, 0.07]).astype(np.uint8)
        occ = (rng.random(n2) < 0.1).astype(np.uint8)
        wr.frame(kb, db); wr.i(n1); wr.f(th); wr.i(orient); wr.parts.append(np.ascontiguousarray(ka).tobytes())
        wr.a(state, np.uint8); wr.a(q_u, np.float32); wr.a(q_v, np.float32); wr.a(da, np.uint8); wr.a(occ, np.uint8)
        inside = (q_u >= 0) & (q_u < W) & (q_v >= 0) & (q_v < H)
        q_r = (np.float32(th) * ka["size"]).astype(np.float32)
        exp["proj%d" % from_kf] = ref.search_by_projection(q_u, q_v, q_r, ka["octave"].astype(np.int32), ka["angle"], da, ((state == 1) & inside).astype(np.uint8),
                                                          kb, db, W, H, occ, bool(orient), bool(from_kf))
    # 4. local map points
    th, ratio = 2.0, 0.8
    q_u = (ka["x"] + 7 + rng.normal(0, 1.0, n1)).astype(np.float32); q_v = (ka["y"] + 3 + rng.normal(0, 1.0, n1)).astype(np.float32)
    state = rng.choice([1, 2], n1, p=[0.9, 0.1]).astype(np.uint8); in_view = (rng.random(n1) < 0.85).astype(np.uint8)
    vc = np.where(rng.random(n1) < 0.5, np.float32(0.9995), np.float32(0.9)).astype(np.float32)
    lvl = ka["octave"].astype(np.int32); occ = (rng.random(n2) < 0.1).astype(np.uint8)
    wr.frame(kb, db); wr.i(n1); wr.f(th, ratio); wr.a(state, np.uint8); wr.a(q_u, np.float32); wr.a(q_v, np.float32); wr.a(da, np.uint8)
    wr.a(in_view, np.uint8); wr.a(vc, np.float32); wr.a(lvl, np.int32); wr.a(q_u, np.float32); wr.a(q_v, np.float32); wr.a(occ, np.uint8)
    exp["local"] = ref.search_local_points(q_u, q_v, vc, lvl, da, ((state == 1) & (in_view == 1)).astype(np.uint8), th, kb, db, W, H, occ, ratio)
    # 5. triangulation
    has1 = (rng.random(n1) < 0.3).astype(np.uint8); has2 = (rng.random(n2) < 0.3).astype(np.uint8)
    f1, f2 = fv(da, 4), fv(db, 4)
    wr.frame(ka, da); wr.frame(kb, db); wr.fvec(f1); wr.fvec(f2); wr.a(has1, np.uint8); wr.a(has2, np.uint8); wr.i(1)
    exp["tri"] = ref.search_for_triangulation(da, ka["angle"], has1, f1, db, kb["angle"], has2, f2, True)
    # 6. bag of words
    st1 = rng.choice([0, 1, 2], n1, p=[0.25, 0.65, 0.1]).astype(np.uint8); occ2 = (rng.random(n2) < 0.2).astype(np.uint8)
    f1, f2 = fv(da, 3), fv(db, 3)
    wr.frame(ka, da); wr.frame(kb, db); wr.fvec(f1); wr.fvec(f2); wr.a(st1, np.uint8); wr.a(occ2, np.uint8); wr.f(0.7); wr.i(1)
    exp["bow"] = ref.search_by_bow(da, ka["angle"], (st1 == 1).astype(np.uint8), f1, db, kb["angle"], occ2, f2, 0.7, True)
    # 7. fuse
    nf, th = 1000, 3.0
    src = rng.permutation(n2)[:nf]                                                      # every map point once
    f_u = (kb["x"][src] + 7 + rng.normal(0, 1.0, nf)).astype(np.float32); f_v = (kb["y"][src] + 3 + rng.normal(0, 1.0, nf)).astype(np.float32)
    f_u[:12] -= 900
    f_l = np.clip(kb["octave"][src] + rng.integers(-1, 2, nf), 0, 7).astype(np.int32)
    state = rng.choice([0, 1, 2, 3], nf, p=[0.05, 0.85, 0.05, 0.05]).astype(np.uint8)
    nobs = rng.integers(1, 6, nf).astype(np.int32)
    slot_obs = np.where(rng.random(n1) < 0.3, rng.integers(1, 6, n1), -1).astype(np.int32)
    wr.frame(ka, da); wr.i(nf); wr.f(th); wr.a(state, np.uint8); wr.a(f_u, np.float32); wr.a(f_v, np.float32); wr.a(db[src], np.uint8)
    wr.a(f_l, np.int32); wr.a(nobs, np.int32); wr.a(slot_obs, np.int32)
    inside = (f_u >= 0) & (f_u < W) & (f_v >= 0) & (f_v < H)
    exp["fuse"] = ref.search_fuse(f_u, f_v, f_l, db[src], ((state == 1) & inside).astype(np.uint8), th, ka, da, W, H)

    pin, pout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    open(pin, "wb").write(b"".join(wr.parts))
    r = subprocess.run([exe, pin, pout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(pout, "rb").read(); pos = [0]

    def take(dt, n):
        arr = np.frombuffer(buf, dt, n, pos[0]).copy(); pos[0] += arr.nbytes
        return arr

    n = int(take(np.int32, 1)[0]); m12 = take(np.int32, n1); pre_out = take(np.float32, 2 * n1).reshape(-1, 2)
    en, em12, epre = exp["init"]
    assert n == en and np.array_equal(m12, em12) and np.array_equal(pre_out, epre) and n > 100
    for key in ("proj0", "proj1", "local"):
        n = int(take(np.int32, 1)[0]); own = take(np.int32, n2)
        en, eown = exp[key]
        assert n == en and np.array_equal(own, eown) and n > 100, key
    n = int(take(np.int32, 1)[0]); m12 = take(np.int32, n1)
    assert n == exp["tri"][0] and np.array_equal(m12, exp["tri"][1]) and n > 20
    n = int(take(np.int32, 1)[0]); own = take(np.int32, n2)
    assert n == exp["bow"][0] and np.array_equal(own, exp["bow"][1]) and n > 20
    n = int(take(np.int32, 1)[0]); fused = take(np.int32, nf); replaced = take(np.int32, nf)
    en, ebest = exp["fuse"]                                                             # the reference's search result per map point (empty key-frame slots)
    assert n == en and n > 100
    hit = ebest >= 0
    empty_slot = np.zeros(nf, bool); empty_slot[hit] = slot_obs[ebest[hit]] < 0
    assert np.array_equal(fused, np.where(hit & empty_slot, ebest, -1))                 # :575-577 addObservation on an empty slot
    more = np.zeros(nf, bool); more[hit] = slot_obs[ebest[hit]] > nobs[hit]
    assert np.array_equal(replaced, np.where(hit & ~empty_slot, np.where(more, 1, 2), 0))   # :578-584 who replaces whom
    queried = take(np.int32, int(hit.sum()))
    assert np.array_equal(queried, ebest[hit]) and int(take(np.int32, 1)[0]) == -12345  # getMapPoint calls in list order
