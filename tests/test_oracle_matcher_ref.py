"""Pins the matcher restatement (oracle/orb_oracle.c: orc_search_*, orc_descriptor_distance, orc_compute_three_maxima) to the
reference's own modules/ORB/ORBMatcher.cpp compiled verbatim (oracle/_ref/libref_matcher.so, see oracle/matcher_harness.cpp): the same
flat inputs go through both, and match counts, assignments, matches12 and the updated vecPreMatched must be identical.  Inputs are real
extractor output (CPU oracle) of a synthetic frame pair plus adversarial edits (duplicate descriptors, occupied slots, invalid
queries).  The library needs /root/reference to build, so these tests run in the build container and wherever oracle/_ref travelled."""
import numpy as np
import pytest

from oracle import orb_oracle as orc
from oracle import ref_matcher as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_matcher.so not built (needs the reference sources)")

W, H = 752, 480


@pytest.fixture(scope="module")
def pair():
    from monoorbslam3_b200 import synth
    orc.build()
    a, b = synth.shifted_pair(H, W, 1000)
    ex = orc.Extractor(1500, 1.2, 8, 20, 7)
    ka, da = ex(a); kb, db = ex(b)
    db = db.copy()
    rng = np.random.default_rng(1)
    dup = rng.integers(0, len(db), 60)
    db[dup] = db[(dup + 1) % len(db)]                      # exact duplicate descriptors: best == second, ties between candidates
    sf = np.array([ex.scale(l) for l in range(8)], np.float32)
    return dict(ka=ka, da=da, kb=kb, db=db, sf=sf)


def _fv(desc, n_bits):
    node = desc[:, 0].astype(np.int32) >> (8 - n_bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


def test_descriptor_distance_and_three_maxima(pair):
    rng = np.random.default_rng(0)
    for _ in range(300):
        i, j = rng.integers(0, len(pair["da"])), rng.integers(0, len(pair["db"]))
        assert orc.descriptor_distance(pair["da"][i], pair["db"][j]) == ref.descriptor_distance(pair["da"][i], pair["db"][j])
    for _ in range(300):
        counts = rng.integers(0, rng.integers(1, 40), 30).astype(np.int32)
        if rng.random() < 0.3:
            counts[rng.integers(0, 30, 3)] = counts.max()           # ties between the maxima
        if rng.random() < 0.1:
            counts[:] = 0
        assert orc.compute_three_maxima(counts) == ref.compute_three_maxima(counts)


@pytest.mark.parametrize("window,ratio,orient", [(100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (200, 1.0, True), (15, 0.6, True)])
def test_search_for_initialization(pair, window, ratio, orient):
    ka, da, kb, db = pair["ka"], pair["da"], pair["kb"], pair["db"]
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    o = orc.search_for_initialization(ka, da, kb, db, W, H, pre, window, ratio, orient)
    r = ref.search_for_initialization(ka, da, kb, db, W, H, pre, window, ratio, orient)
    assert o[0] == r[0] and o[0] > 20 and np.array_equal(o[1], r[1]) and np.array_equal(o[2], r[2])
    o2 = orc.search_for_initialization(ka, da, kb, db, W, H, o[2], window, ratio, orient)      # second call on the updated vecPreMatched
    r2 = ref.search_for_initialization(ka, da, kb, db, W, H, r[2], window, ratio, orient)
    assert o2[0] == r2[0] and np.array_equal(o2[1], r2[1]) and np.array_equal(o2[2], r2[2])


def _queries(pair, rng, th):
    ka = pair["ka"]; nq = len(ka)
    q_u = (ka["x"] + 7 + rng.normal(0, 1.0, nq)).astype(np.float32); q_v = (ka["y"] + 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
    q_r = (np.float32(th) * ka["size"]).astype(np.float32)
    q_valid = (rng.random(nq) < 0.8).astype(np.uint8)
    occupied = (rng.random(len(pair["kb"])) < 0.1).astype(np.uint8)
    return q_u, q_v, q_r, ka["octave"].astype(np.int32), ka["angle"].astype(np.float32), pair["da"], q_valid, occupied


@pytest.mark.parametrize("th,orient,from_kf", [(15, True, False), (30, True, False), (15, False, False), (15, True, True), (7, True, True)])
def test_search_by_projection_frame_and_keyframe(pair, th, orient, from_kf):
    rng = np.random.default_rng(th)
    q_u, q_v, q_r, q_l, q_a, q_d, q_valid, occ = _queries(pair, rng, th)
    o = orc.search_by_projection(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, pair["kb"], pair["db"], W, H, occ, orient)
    r = ref.search_by_projection(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, pair["kb"], pair["db"], W, H, occ, orient, from_keyframe=from_kf)
    assert o[0] == r[0] and o[0] > 50 and np.array_equal(o[1], r[1])


@pytest.mark.parametrize("th,ratio", [(1, 0.8), (2, 0.8), (15, 0.6), (4, 1.0)])
def test_search_local_points(pair, th, ratio):
    rng = np.random.default_rng(100 + th)
    q_u, q_v, _, q_l, _, q_d, q_valid, occ = _queries(pair, rng, th)
    view_cos = np.where(rng.random(len(q_l)) < 0.5, np.float32(0.9995), np.float32(0.9)).astype(np.float32)
    # ORBMatcher.cpp:361-364: radius = th; radius *= (cos > 0.998 ? 2.5f : 4.f); radius *= scaleFactor[level]  (float, in this order)
    q_r = ((np.float32(th) * np.where(view_cos > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)).astype(np.float32) * pair["sf"][q_l]).astype(np.float32)
    o = orc.search_local_points(q_u, q_v, q_r, q_l, q_d, q_valid, pair["kb"], pair["db"], W, H, occ, ratio)
    r = ref.search_local_points(q_u, q_v, view_cos, q_l, q_d, q_valid, th, pair["kb"], pair["db"], W, H, occ, ratio)
    assert o[0] == r[0] and o[0] > 10 and np.array_equal(o[1], r[1])


@pytest.mark.parametrize("bits,orient", [(3, False), (5, False), (4, True), (1, True)])
def test_search_for_triangulation(pair, bits, orient):
    rng = np.random.default_rng(bits)
    da, db = pair["da"], pair["db"]
    has1 = (rng.random(len(da)) < 0.3).astype(np.uint8); has2 = (rng.random(len(db)) < 0.3).astype(np.uint8)
    fv1, fv2 = _fv(da, bits), _fv(db, bits)
    o = orc.search_for_triangulation(da, pair["ka"]["angle"], has1, fv1, db, pair["kb"]["angle"], has2, fv2, orient)
    r = ref.search_for_triangulation(da, pair["ka"]["angle"], has1, fv1, db, pair["kb"]["angle"], has2, fv2, orient)
    assert o[0] == r[0] and o[0] > 10 and np.array_equal(o[1], r[1])
    assert not (r[1] == 0).any()                          # ORBMatcher.cpp:484: the reference itself never accepts index 0


@pytest.mark.parametrize("bits,ratio,orient", [(3, 0.7, True), (5, 0.7, True), (4, 0.9, False), (2, 0.6, True)])
def test_search_by_bow(pair, bits, ratio, orient):
    rng = np.random.default_rng(100 + bits)
    da, db = pair["da"], pair["db"]
    valid1 = (rng.random(len(da)) < 0.7).astype(np.uint8); occ2 = (rng.random(len(db)) < 0.2).astype(np.uint8)
    fv1, fv2 = _fv(da, bits), _fv(db, bits)
    if bits == 5:                                           # node sets that differ: exercises the lower_bound skips
        fv1 = tuple(x.copy() for x in fv1); keep = np.ones(len(fv1[0]), bool); keep[::3] = False
        idx = np.concatenate([fv1[2][fv1[1][i]:fv1[1][i + 1]] for i in np.nonzero(keep)[0]])
        off = np.concatenate([[0], np.cumsum([fv1[1][i + 1] - fv1[1][i] for i in np.nonzero(keep)[0]])]).astype(np.int32)
        fv1 = (fv1[0][keep], off, idx.astype(np.int32))
    o = orc.search_by_bow(da, pair["ka"]["angle"], valid1, fv1, db, pair["kb"]["angle"], occ2, fv2, ratio, orient)
    r = ref.search_by_bow(da, pair["ka"]["angle"], valid1, fv1, db, pair["kb"]["angle"], occ2, fv2, ratio, orient)
    assert o[0] == r[0] and o[0] > 10 and np.array_equal(o[1], r[1])


@pytest.mark.parametrize("th", [3.0, 5.0, 1.5])
def test_search_fuse(pair, th):
    rng = np.random.default_rng(int(th * 10))
    ka, da, kb, db = pair["ka"], pair["da"], pair["kb"], pair["db"]
    nq = 1200
    src = rng.integers(0, len(kb), nq)
    u = (kb["x"][src] + 7 + rng.normal(0, 1.0, nq)).astype(np.float32); v = (kb["y"][src] + 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
    level = np.clip(kb["octave"][src] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    radius = (np.float32(th) * pair["sf"][level]).astype(np.float32)          # ORBMatcher.cpp:551
    valid = (rng.random(nq) < 0.9).astype(np.uint8)
    qd = db[src].copy()
    o = orc.search_fuse(u, v, radius, level, qd, valid, ka, da, W, H, pair["sf"] * pair["sf"])
    r = ref.search_fuse(u, v, level, qd, valid, th, ka, da, W, H)
    assert o[0] == r[0] and o[0] > 50 and np.array_equal(o[1], r[1])
