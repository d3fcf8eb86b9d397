"""GPU parity on the BASELINE.json configurations round 1 left untested, all through the C-ABI:
  config 3  1920x1080 frame pair, 8000-feature initial extractor, SearchForInitialization(window 100, ratio 0.9)
            (Tracking.cpp:24, 606; test/ORB/initializeSearchTest.cpp) against the verbatim-reference fixture and the oracle
  config 4  a window of 20 key frames x 2000 key points: SearchForTriangulation / SearchByProjection per key-frame pair
            (LocalMapping.cpp:148-168), and the 40 000 x 40 000 all-pairs search against the oracle
  recorded frames: two real photographs against the verbatim-reference fixture and the oracle
  quadtree: inputs that need more node slots than frames do in practice (the shared-memory pool moves to the global one)
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def mods():
    import monoorbslam3_b200 as m
    from monoorbslam3_b200 import synth
    return m, synth


def same_keypoints(kps, desc, okps, odesc):
    assert len(kps) == len(okps), (len(kps), len(okps))
    assert kps.tobytes() == okps.tobytes()                   # every field bit for bit, angles included
    assert np.array_equal(desc, odesc)


# ---------------------------------------------------------------- config 3
def test_config3_init_search_1080p_8000(mods, oracle):
    from test_oracle_golden_r2 import c3_inputs
    m, synth = mods
    g, a, b, ka, da, kb, db = c3_inputs(oracle)
    ex = m.ORBExtractor(8000, 1.2, 8, 20, 7)
    ga, gda = ex(a); gb, gdb = ex(b)
    same_keypoints(ga, gda, ka, da); same_keypoints(gb, gdb, kb, db)
    mt = m.ORBMatcher(0.9, True)
    f1 = m.FrameView(ga, gda, 1920, 1080); f2 = m.FrameView(gb, gdb, 1920, 1080)
    pre = np.stack([ga["x"], ga["y"]], 1).astype(np.float32)
    n, m12 = mt.SearchForInitialization(f1, f2, pre, 100)
    assert n == int(g["n"]) and np.array_equal(m12, g["m12"]) and np.array_equal(pre, g["pre"])       # reference's own ORBMatcher.cpp
    assert n > 1000
    # a second call on the updated vecPreMatched (Tracking re-enters with it), other window / ratio / no orientation check
    for window, ratio, orient in ((100, 0.9, True), (30, 0.7, False), (200, 0.95, True)):
        mt2 = m.ORBMatcher(ratio, orient)
        p1 = pre.copy()
        n1, m1 = mt2.SearchForInitialization(f1, f2, p1, window)
        on, om, op = oracle.search_for_initialization(ka, da, kb, db, 1920, 1080, pre.copy(), window, ratio, orient)
        assert n1 == on and np.array_equal(m1, om) and np.array_equal(p1, op), (window, ratio, orient)
    ex.close()


# ---------------------------------------------------------------- config 4
def feature_vector(desc, bits):
    node = desc[:, 0].astype(np.int32) >> (8 - bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


@pytest.fixture(scope="module")
def window20(mods):
    """20 key frames of one scene (a camera sliding right / down over a large synthetic image, fresh sensor noise per frame),
    KITTI shape, 2000 features each: about 40 k descriptors."""
    m, synth = mods
    W, H, N = 1241, 376, 20
    big = synth.frame(H + 4 * N + 8, W + 9 * N + 8, 4242, "dense").astype(np.int16)
    rng = np.random.default_rng(99)
    ex = m.ORBExtractor(2000, 1.2, 8, 20, 7, max_batch=N)
    frames = np.stack([np.clip(big[4 + 3 * k:4 + 3 * k + H, 4 + 7 * k:4 + 7 * k + W] + np.rint(rng.normal(0, 1.5, (H, W))).astype(np.int16), 0, 255).astype(np.uint8)
                       for k in range(N)])
    n, kps, desc = ex.extract_batch(np.ascontiguousarray(frames))
    kfs = [(kps[k, :n[k]].copy(), desc[k, :n[k]].copy()) for k in range(N)]
    ex.close()
    return W, H, kfs


def test_config4_window_searches_over_20_keyframes(mods, oracle, window20):
    m, _ = mods
    W, H, kfs = window20
    assert sum(len(k) for k, _ in kfs) > 39000
    rng = np.random.default_rng(7)
    pairs = [(k, k + 1) for k in range(19)] + [(0, 5), (3, 12), (19, 0), (10, 10)]
    tri = m.ORBMatcher(0.6, False); proj = m.ORBMatcher(0.9, True)
    n_tri = n_proj = 0
    for i, j in pairs:
        (k1, d1), (k2, d2) = kfs[i], kfs[j]
        has1 = (rng.random(len(k1)) < 0.4).astype(np.uint8); has2 = (rng.random(len(k2)) < 0.4).astype(np.uint8)
        bits = 4 + (i % 3)
        fv1, fv2 = feature_vector(d1, bits), feature_vector(d2, bits)
        n, m12 = tri.SearchForTriangulation(d1, k1["angle"], has1, fv1, d2, k2["angle"], has2, fv2)
        on, om12 = oracle.search_for_triangulation(d1, k1["angle"], has1, fv1, d2, k2["angle"], has2, fv2, False)
        assert n == on and np.array_equal(m12, om12), ("triangulation", i, j)
        n_tri += n
        # SearchByProjection(KeyFrame -> Frame): key frame i's points projected into frame j (known shift + projection noise)
        dx, dy = 7 * (i - j), 3 * (i - j)
        q_u = (k1["x"] + dx + rng.normal(0, 1.0, len(k1))).astype(np.float32); q_v = (k1["y"] + dy + rng.normal(0, 1.0, len(k1))).astype(np.float32)
        q_valid = (rng.random(len(k1)) < 0.8).astype(np.uint8); occ = (rng.random(len(k2)) < 0.1).astype(np.uint8)
        q_r = (np.float32(15 if i % 2 else 30) * k1["size"]).astype(np.float32)
        f2 = m.FrameView(k2, d2, W, H)
        n, asg = proj.SearchByProjection(q_u, q_v, q_r, k1["octave"], k1["angle"], d1, q_valid, f2, occ)
        on, oasg = oracle.search_by_projection(q_u, q_v, q_r, k1["octave"], k1["angle"], d1, q_valid, k2, d2, W, H, occ, True)
        assert n == on and np.array_equal(asg, oasg), ("projection", i, j)
        n_proj += n
    assert n_tri > 2000 and n_proj > 10000


def oracle_allpairs_threaded(oracle, q, t, threads=None):
    threads = threads or min(32, os.cpu_count() or 1)
    blocks = np.array_split(np.arange(len(q)), threads * 4)
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda r: oracle.hamming_allpairs(q[r[0]:r[-1] + 1], t) if len(r) else (np.zeros(0, np.int32),) * 3, blocks))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))


def test_config4_allpairs_40k_against_the_oracle(mods, oracle, window20):
    """40 000 x 40 000: the window's own descriptors (about 40 k, real extractor output, every row finds itself at distance 0 and the
    second distance carries the information) topped up with random rows, and a shuffled copy as the train side."""
    m, _ = mods
    _, _, kfs = window20
    rng = np.random.default_rng(3)
    table = np.concatenate([d for _, d in kfs])
    table = np.concatenate([table, rng.integers(0, 256, (max(0, 40000 - len(table)), 32), dtype=np.uint8)])[:40000]
    train = table[rng.permutation(len(table))].copy()
    train[rng.integers(0, 40000, 300)] = train[rng.integers(0, 40000, 300)]                 # exact duplicates: first index wins
    mt = m.ORBMatcher()
    got = mt.hamming_allpairs(table, train)
    exp = oracle_allpairs_threaded(oracle, table, train)
    for a, b, what in zip(got, exp, ("index", "best", "second")):
        assert np.array_equal(a, b), what


# ---------------------------------------------------------------- recorded frames
@pytest.mark.parametrize("name", ["china", "flower"])
def test_real_photographs(mods, oracle, name):
    m, _ = mods
    g = np.load(os.path.join(GOLD, "photos_ref.npz"))
    img, nf = g["img_" + name], int(g["nf_" + name])
    for use_tma in (True, False):
        ex = m.ORBExtractor(nf, 1.2, 8, 20, 7, use_tma=use_tma, keep_stages=True)
        kps, desc = ex(img)
        same_keypoints(kps, desc, g["kps_" + name], g["desc_" + name])                         # reference's own ORBExtractor.cpp
        oc = oracle.Extractor(nf, 1.2, 8, 20, 7)
        oc(img)
        for l in range(8):
            c = oc.level_candidates(l)
            assert np.array_equal(ex.level_candidates(l), np.stack([c["x"], c["y"], c["score"]], 1).reshape(-1, 3)), (name, l)
        ex.close()
    # the same photograph inside a batch and through the device-resident entry point's host mirror
    ex = m.ORBExtractor(nf, 1.2, 8, 20, 7, max_batch=3)
    n, kps, desc = ex.extract_batch(np.ascontiguousarray(np.stack([img, img[::-1], img])))
    same_keypoints(kps[0, :n[0]], desc[0, :n[0]], g["kps_" + name], g["desc_" + name])
    same_keypoints(kps[2, :n[2]], desc[2, :n[2]], g["kps_" + name], g["desc_" + name])
    ex.close()


# ---------------------------------------------------------------- quadtree node pool
def dots_image(h, w, pts):
    """White single-pixel dots on black: each dot is exactly one FAST corner (all 16 ring pixels darker) and nothing else is."""
    img = np.zeros((h, w), np.uint8)
    img[pts[:, 1] + 19, pts[:, 0] + 19] = 255
    return img


@pytest.mark.parametrize("gx,gy,cluster", [(26, 16, 8), (20, 20, 8)])
def test_quadtree_inputs_beyond_the_typical_pool(mods, oracle, gx, gy, cluster):
    """The sparse tight-pair inputs of tests/test_octree_model.py as an image: level 0 needs more node slots than the round-1 pool
    formula allowed (the call failed with ORBFE_E_INTERNAL); the pool now grows into the proven-size global pool."""
    from test_octree_model import tight_pairs
    m, _ = mods
    pts = tight_pairs(1882, 1042, gx, gy, cluster)
    img = dots_image(1080, 1920, pts)
    ex = m.ORBExtractor(2687, 1.2, 8, 20, 7, keep_stages=True)                # level-0 quota 868
    assert ex.getFeaturesPerLevel(0) == 868
    kps, desc = ex(img)
    okps, odesc = oracle.Extractor(2687, 1.2, 8, 20, 7)(img)
    cand = ex.level_candidates(0)
    assert np.array_equal(cand[:, :2], pts)                                  # every dot, and nothing else, is a level-0 candidate
    from test_octree_model import gpu_formulation
    _, slots = gpu_formulation(cand[:, 0].astype(int), cand[:, 1].astype(int), cand[:, 2].astype(int), 1882, 1042, 1882 + 19, ex.getFeaturesPerLevel(0))
    assert slots > 8 * (ex.getFeaturesPerLevel(0) + 4) + 5 * 2 + 64          # more than round 1's pool held
    same_keypoints(kps, desc, okps, odesc)
    ex.close()


def test_quadtree_pool_migration_is_exact(mods, oracle, monkeypatch):
    """ORBFE_OCT_SMEM_NODES shrinks the shared-memory node pool so that ordinary frames move to the global pool in the middle of
    the subdivision (at different passes on different levels): the result must not change."""
    m, synth = mods
    monkeypatch.setenv("ORBFE_OCT_SMEM_NODES", "96")
    frames = np.ascontiguousarray(np.stack([synth.frame(480, 752, 31, "dense"), synth.frame(480, 752, 32, "natural")]))
    ex = m.ORBExtractor(1000, 1.2, 8, 20, 7, max_batch=2)
    n, kps, desc = ex.extract_batch(frames)
    oc = oracle.Extractor(1000, 1.2, 8, 20, 7)
    for b in range(2):
        same_keypoints(kps[b, :n[b]], desc[b, :n[b]], *oc(frames[b]))
        k1, d1 = ex(frames[b])
        same_keypoints(k1, d1, *oc(frames[b]))
    ex.close()
