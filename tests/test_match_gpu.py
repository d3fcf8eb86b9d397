"""T3 (GPU): the Hamming matchers, called through the C-ABI, against the flat-array oracle restatements of ORBMatcher.cpp.
Integer work: everything is compared bit-exactly (indices, distances, match counts, updated vecPreMatched)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, synth
    ex = ORBExtractor(2000, 1.2, 8, 20, 7)
    a, b = synth.shifted_pair(480, 752, 1000)
    ka, da = ex(a); kb, db = ex(b)
    return dict(ORBMatcher=ORBMatcher, FrameView=FrameView, synth=synth, ex=ex, ka=ka, da=da, kb=kb, db=db, w=752, h=480)


def test_descriptor_distance(ctx, oracle):
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(0)
    ia = rng.integers(0, len(ctx["da"]), 500); ib = rng.integers(0, len(ctx["db"]), 500)
    got = m.descriptor_distances(ctx["da"], ctx["db"], ia, ib)
    exp = [oracle.descriptor_distance(ctx["da"][i], ctx["db"][j]) for i, j in zip(ia, ib)]
    assert got.tolist() == exp
    assert m.DescriptorDistance(ctx["da"][0], ctx["da"][0]) == 0
    assert m.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


# problems of at least 256 x 256 run on the tensor cores (k_allpairs_imma), smaller ones on the popc kernel: shapes on both sides of
# the switch, ragged last tiles (nt % 64, nq % 128), one train split and several
@pytest.mark.parametrize("nq,nt", [(1, 1), (7, 513), (64, 512), (65, 1025), (2012, 2009), (300, 5000), (256, 256), (257, 319), (1153, 4097), (129, 300),
                                   (5000, 257)])
def test_allpairs(ctx, oracle, nq, nt):
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[rng.integers(0, nt, max(nt // 10, 1))] = q[rng.integers(0, nq, max(nt // 10, 1))]      # exact duplicates: distance-0 ties, first index must win
    bi, bd, sd = m.hamming_allpairs(q, t)
    obi, obd, osd = oracle.hamming_allpairs(q, t)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)


def test_allpairs_tensor_core_and_popc_kernels_agree(ctx, monkeypatch):
    """The +-1 int8 GEMM formulation and the popc kernel give identical (index, best, second) on real descriptors and on
    adversarial ones (all-equal rows: every distance 0, the first index must win; complementary rows: distance 256)."""
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(11)
    q = np.concatenate([ctx["da"], ctx["db"]])[:3000].copy(); t = np.concatenate([ctx["db"], ctx["da"]])[:2900].copy()
    t[100:400] = t[100]                                   # 300 identical train rows
    q[5] = t[100]; q[6] = ~t[100]                         # distance 0 against all of them / distance 256
    t[rng.integers(0, len(t), 200)] = q[rng.integers(0, len(q), 200)]
    got = m.hamming_allpairs(q, t)
    monkeypatch.setenv("ORBFE_ALLPAIRS_POPC", "1")
    ref = m.hamming_allpairs(q, t)
    monkeypatch.delenv("ORBFE_ALLPAIRS_POPC")
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    monkeypatch.setenv("ORBFE_ALLPAIRS", "imma")
    assert all(np.array_equal(a, b) for a, b in zip(m.hamming_allpairs(q, t), ref))
    monkeypatch.delenv("ORBFE_ALLPAIRS")
    assert got[1][5] == 0 and got[0][5] <= 100 and got[2][5] == 0


def numpy_allpairs(q, t, excl=None):
    """Brute force in numpy (small shapes): first minimum wins; 257 where there is no (second) candidate."""
    d = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2).astype(np.int32)
    if excl is not None:
        cols = np.arange(len(t))[None, :]
        d[(cols >= excl[:, :1]) & (cols < excl[:, 1:2])] = 1000
    bi = d.argmin(1).astype(np.int32); bd = d.min(1)
    d2 = d.copy(); d2[np.arange(len(q)), bi] = 1000
    sd = d2.min(1)
    bi[bd >= 1000] = -1
    return bi, np.where(bd >= 1000, 257, bd).astype(np.int32), np.where(sd >= 1000, 257, sd).astype(np.int32)


@pytest.mark.parametrize("kernel", ["tc", "imma", "popc"])
@pytest.mark.parametrize("nq,nt", [(256, 512), (257, 513), (700, 900), (1025, 3000), (2600, 641), (511, 12800)])
def test_allpairs_every_kernel_against_numpy(ctx, monkeypatch, kernel, nq, nt):
    """tcgen05 (k_allpairs_tc), mma.sync (k_allpairs_imma) and popc (k_hamming_allpairs) kernels on shapes with ragged query / train
    tiles, one and several train splits, exact duplicates (ties: the first index wins) and complementary rows (distance 256)."""
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(nq * 13 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[rng.integers(0, nt, nt // 8)] = q[rng.integers(0, nq, nt // 8)]
    t[nt // 2:nt // 2 + 40] = t[nt // 2]; q[3] = t[nt // 2]; q[4] = ~t[nt // 2]
    monkeypatch.setenv("ORBFE_ALLPAIRS", kernel)
    got = m.hamming_allpairs(q, t)
    exp = numpy_allpairs(q, t) if nq * nt <= 4_000_000 else None
    if exp is None:
        monkeypatch.setenv("ORBFE_ALLPAIRS", "popc")
        exp = m.hamming_allpairs(q, t)
    for a, b, what in zip(got, exp, ("index", "best", "second")):
        assert np.array_equal(a, b), (kernel, what)


@pytest.mark.parametrize("kernel", ["tc", "popc"])
def test_allpairs_with_keyframe_block_exclusion(ctx, oracle, monkeypatch, kernel):
    """A key-frame window matched against itself: every descriptor skips the block of its own key frame (orbfe_hamming_allpairs_excl).
    Blocks of uneven size that straddle tile boundaries, an empty range, a range covering everything (no candidate: -1 / 257)."""
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(5)
    sizes = [130, 257, 1, 400, 128, 383, 64, 300, 900]                                 # 900: several train tiles lie inside the own block of whole warps
    n = sum(sizes)
    table = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    table[rng.integers(0, n, 200)] = table[rng.integers(0, n, 200)]                    # duplicates inside and across blocks
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    blk = np.repeat(np.arange(len(sizes)), sizes)
    excl = np.stack([starts[blk], starts[blk + 1]], 1).astype(np.int32)
    excl[5] = (7, 7); excl[6] = (0, n)
    monkeypatch.setenv("ORBFE_ALLPAIRS", kernel)
    got = m.hamming_allpairs(table, table, excl)
    exp = numpy_allpairs(table, table, excl)
    for a, b, what in zip(got, exp, ("index", "best", "second")):
        assert np.array_equal(a, b), (kernel, what)
    assert got[0][6] == -1 and got[1][6] == 257 and got[1][5] == 0
    own = (got[0] >= excl[:, 0]) & (got[0] < excl[:, 1])
    own[5] = False
    assert not own.any()                                                               # nobody matched inside its own block
    # without exclusion every row finds itself
    bi, bd, _ = m.hamming_allpairs(table, table)
    assert (bd == 0).all()


def test_allpairs_real_descriptors_and_empty(ctx, oracle):
    m = ctx["ORBMatcher"]()
    bi, bd, sd = m.hamming_allpairs(ctx["da"], ctx["db"])
    obi, obd, osd = oracle.hamming_allpairs(ctx["da"], ctx["db"])
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)
    bi, bd, sd = m.hamming_allpairs(ctx["da"][:5], np.zeros((0, 32), np.uint8))
    assert bi.tolist() == [-1] * 5 and bd.tolist() == [257] * 5 and sd.tolist() == [257] * 5


@pytest.mark.parametrize("nq,nt,max_len", [(1, 1, 1), (9, 40, 3), (500, 700, 40), (257, 3000, 300)])
def test_hamming_window(ctx, oracle, nq, nt, max_len):
    """orbfe_hamming_window: ragged candidate lists (empty, single, long, repeated candidates, exact duplicates)."""
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(nq * 31 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[rng.integers(0, nt, max(nt // 8, 1))] = q[rng.integers(0, nq, max(nt // 8, 1))]
    lens = rng.integers(0, max_len + 1, nq); lens[0] = max_len; lens[-1] = 0 if nq > 1 else lens[-1]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    idx = rng.integers(0, nt, off[-1]).astype(np.int32)
    bi, bd, sd = m.hamming_window(q, t, off, idx)
    obi, obd, osd = oracle.hamming_window(q, t, off, idx)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)
    # the full list 0..nt-1 for every query is the all-pairs search
    full_off = (np.arange(nq + 1) * nt).astype(np.int32); full_idx = np.tile(np.arange(nt, dtype=np.int32), nq)
    assert all(np.array_equal(a, b) for a, b in zip(m.hamming_window(q, t, full_off, full_idx), m.hamming_allpairs(q, t)))
    with pytest.raises(Exception):
        m.hamming_window(q, t, off, np.where(np.arange(len(idx)) == 0, nt, idx).astype(np.int32))     # candidate out of range


@pytest.mark.parametrize("window,ratio,orient", [(100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (200, 1.0, True)])
def test_search_for_initialization(ctx, oracle, window, ratio, orient):
    m = ctx["ORBMatcher"](ratio, orient)
    FV = ctx["FrameView"]
    f1 = FV(ctx["ka"], ctx["da"], ctx["w"], ctx["h"]); f2 = FV(ctx["kb"], ctx["db"], ctx["w"], ctx["h"])
    pre = np.stack([ctx["ka"]["x"], ctx["ka"]["y"]], 1).astype(np.float32)
    opre = pre.copy()
    n, m12 = m.SearchForInitialization(f1, f2, pre, window)
    on, om12, opre = oracle.search_for_initialization(ctx["ka"], ctx["da"], ctx["kb"], ctx["db"], ctx["w"], ctx["h"], opre, window, ratio, orient)
    assert n == on and n > 50
    assert np.array_equal(m12, om12) and np.array_equal(pre, opre)
    # second call with the updated vecPreMatched (what Tracking::Initialization does frame after frame)
    n2, m12b = m.SearchForInitialization(f1, f2, pre, window)
    on2, om12b, _ = oracle.search_for_initialization(ctx["ka"], ctx["da"], ctx["kb"], ctx["db"], ctx["w"], ctx["h"], opre, window, ratio, orient)
    assert n2 == on2 and np.array_equal(m12b, om12b)


def _projection_queries(ctx, rng, th):
    """Stand-in for the adapter's projection step: last-frame key points re-projected with a small motion + noise."""
    ka = ctx["ka"]
    nq = len(ka)
    q_u = (ka["x"] + 7 + rng.normal(0, 1.0, nq)).astype(np.float32); q_v = (ka["y"] + 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
    q_r = (th * ka["size"]).astype(np.float32)
    q_valid = (rng.random(nq) < 0.8).astype(np.uint8)
    occupied = (rng.random(len(ctx["kb"])) < 0.1).astype(np.uint8)
    return q_u, q_v, q_r, ka["octave"].astype(np.int32), ka["angle"].astype(np.float32), ctx["da"], q_valid, occupied


@pytest.mark.parametrize("th,orient", [(15, True), (30, True), (15, False)])
def test_search_by_projection(ctx, oracle, th, orient):
    rng = np.random.default_rng(th)
    q_u, q_v, q_r, q_l, q_a, q_d, q_valid, occ = _projection_queries(ctx, rng, th)
    m = ctx["ORBMatcher"](0.9, orient)
    cur = ctx["FrameView"](ctx["kb"], ctx["db"], ctx["w"], ctx["h"])
    n, assigned = m.SearchByProjection(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, cur, occ)
    on, oassigned = oracle.search_by_projection(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, ctx["kb"], ctx["db"], ctx["w"], ctx["h"], occ, orient)
    assert n == on and n > 50 and np.array_equal(assigned, oassigned)


@pytest.mark.parametrize("th,ratio", [(1, 0.8), (2, 0.8), (15, 0.6)])
def test_search_local_points(ctx, oracle, th, ratio):
    rng = np.random.default_rng(100 + th)
    q_u, q_v, _, q_l, _, q_d, q_valid, occ = _projection_queries(ctx, rng, th)
    scale = np.array([ctx["ex"].getScaleFactor(int(l)) for l in q_l], np.float32)
    q_r = (np.float32(th) * np.where(rng.random(len(q_l)) < 0.5, np.float32(2.5), np.float32(4.0)).astype(np.float32) * scale).astype(np.float32)
    m = ctx["ORBMatcher"](ratio, True)
    fr = ctx["FrameView"](ctx["kb"], ctx["db"], ctx["w"], ctx["h"])
    n, assigned = m.SearchLocalPoints(q_u, q_v, q_r, q_l, q_d, q_valid, fr, occ)
    on, oassigned = oracle.search_local_points(q_u, q_v, q_r, q_l, q_d, q_valid, ctx["kb"], ctx["db"], ctx["w"], ctx["h"], occ, ratio)
    assert n == on and n > 20 and np.array_equal(assigned, oassigned)


def _feature_vector(desc, n_bits):
    """Synthetic DBoW2 FeatureVector (no vocabulary file in the reference repo): node id = leading descriptor bits."""
    node = (desc[:, 0].astype(np.int32) >> (8 - n_bits))
    ids = np.unique(node)
    off = [0]; idx = []
    for i in ids:
        members = np.nonzero(node == i)[0]
        idx.extend(members.tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


@pytest.mark.parametrize("bits,orient", [(3, False), (5, False), (4, True)])
def test_search_for_triangulation(ctx, oracle, bits, orient):
    rng = np.random.default_rng(bits)
    da, db = ctx["da"], ctx["db"]
    has1 = (rng.random(len(da)) < 0.3).astype(np.uint8); has2 = (rng.random(len(db)) < 0.3).astype(np.uint8)
    fv1, fv2 = _feature_vector(da, bits), _feature_vector(db, bits)
    m = ctx["ORBMatcher"](0.6, orient)
    n, m12 = m.SearchForTriangulation(da, ctx["ka"]["angle"], has1, fv1, db, ctx["kb"]["angle"], has2, fv2)
    on, om12 = oracle.search_for_triangulation(da, ctx["ka"]["angle"], has1, fv1, db, ctx["kb"]["angle"], has2, fv2, orient)
    assert n == on and n > 20 and np.array_equal(m12, om12)
    assert not (m12 == 0).any()                       # ORBMatcher.cpp:484: index 0 is never accepted (sic)


@pytest.mark.parametrize("bits,ratio,orient", [(3, 0.7, True), (5, 0.7, True), (4, 0.9, False), (2, 0.6, True)])
def test_search_by_bow(ctx, oracle, bits, ratio, orient):
    """ORBMatcher::SearchByBow (ORBMatcher.cpp:118-201): key-frame key points with a good map point against the frame's free
    slots of the same vocabulary node, float ratio test, rotation histogram."""
    rng = np.random.default_rng(100 + bits)
    da, db = ctx["da"], ctx["db"]
    valid1 = (rng.random(len(da)) < 0.7).astype(np.uint8); occ2 = (rng.random(len(db)) < 0.2).astype(np.uint8)
    fv1, fv2 = _feature_vector(da, bits), _feature_vector(db, bits)
    m = ctx["ORBMatcher"](ratio, orient)
    n, asg = m.SearchByBow(da, ctx["ka"]["angle"], valid1, fv1, db, ctx["kb"]["angle"], occ2, fv2)
    on, oasg = oracle.search_by_bow(da, ctx["ka"]["angle"], valid1, fv1, db, ctx["kb"]["angle"], occ2, fv2, ratio, orient)
    assert n == on and n > 20 and np.array_equal(asg, oasg)
    assert not (asg[occ2 != 0] >= 0).any() and valid1[asg[asg >= 0]].all()
    # identical descriptors on both sides: distance-0 ties (second == best == 0 fails the ratio test, :164)
    n2, asg2 = m.SearchByBow(da, ctx["ka"]["angle"], valid1, fv1, da, ctx["ka"]["angle"], np.zeros(len(da), np.uint8), fv1)
    on2, oasg2 = oracle.search_by_bow(da, ctx["ka"]["angle"], valid1, fv1, da, ctx["ka"]["angle"], np.zeros(len(da), np.uint8), fv1, ratio, orient)
    assert n2 == on2 and np.array_equal(asg2, oasg2)


@pytest.mark.parametrize("th", [3.0, 6.0])
def test_search_fuse(ctx, oracle, th):
    """Search half of the fuse SearchByProjection(KeyFrame, mapPoints) (ORBMatcher.cpp:524-571): strict window, chi-square gate, first minimum."""
    rng = np.random.default_rng(int(th))
    ka, da, kb, db = ctx["ka"], ctx["da"], ctx["kb"], ctx["db"]
    nq = 1500
    src = rng.integers(0, len(kb), nq)                            # map points observed around frame-b key points, projected into frame a
    u = (kb["x"][src] + 7 + rng.normal(0, 1.0, nq)).astype(np.float32); v = (kb["y"][src] + 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
    level = np.clip(kb["octave"][src] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    sf = np.array([ctx["ex"].getScaleFactor(l) for l in range(8)], np.float32)
    radius = (np.float32(th) * sf[level]).astype(np.float32)
    valid = (rng.random(nq) < 0.9).astype(np.uint8)
    qd = db[src].copy()
    m = ctx["ORBMatcher"](handle=ctx["ex"]._h)
    n, bi, bd = m.SearchFuse(ctx["FrameView"](ka, da, ctx["w"], ctx["h"]), u, v, radius, level, qd, valid)
    on, obi, obd = oracle.search_fuse(u, v, radius, level, qd, valid, ka, da, ctx["w"], ctx["h"], sf * sf)
    assert n == on and n > 100 and np.array_equal(bi, obi) and np.array_equal(bd, obd)
    assert (bi[valid == 0] == -1).all()


def test_compute_descriptors(ctx, oracle):
    """MapPoint::computeDescriptor (MapPoint.cpp:103-152): least median distance, first minimum wins; group sizes 0..70."""
    rng = np.random.default_rng(9)
    sizes = [1, 2, 3, 0, 7, 32, 33, 70, 5, 2, 2, 40] + rng.integers(1, 25, 300).tolist()
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = ctx["da"][rng.integers(0, len(ctx["da"]), len(sizes))]
    desc = np.repeat(base, sizes, axis=0).copy()
    noise = (rng.random((len(desc), 32, 8)) < 0.06)
    desc ^= np.packbits(noise, axis=2).reshape(len(desc), 32)
    desc[off[4]:off[4] + 3] = desc[off[4]]                        # exact duplicates: median ties, the first row must win
    m = ctx["ORBMatcher"](handle=ctx["ex"]._h)
    got = m.compute_descriptors(desc, off)
    assert np.array_equal(got, oracle.compute_descriptors(desc, off))
    assert got[3] == -1


def test_tracking_and_local_mapping_threads_run_concurrently(ctx, oracle):
    """SURVEY section 8b, threading: the reference runs ORBMatcher on the tracking thread and on the local-mapping thread at the same
    time (System.cpp:55) while the tracking thread also extracts.  One handle per host thread, no shared mutable state: every
    result of the concurrent runs equals the single-threaded one."""
    import threading
    from monoorbslam3_b200 import ORBExtractor, matcher as matcher_mod
    FV = ctx["FrameView"]
    a, b = ctx["synth"].shifted_pair(480, 752, 1000)
    da, db, ka, kb = ctx["da"], ctx["db"], ctx["ka"], ctx["kb"]
    rng = np.random.default_rng(3)
    has1 = (rng.random(len(da)) < 0.3).astype(np.uint8); has2 = (rng.random(len(db)) < 0.3).astype(np.uint8)
    fv1, fv2 = _feature_vector(da, 4), _feature_vector(db, 4)
    exp_tri = oracle.search_for_triangulation(da, ka["angle"], has1, fv1, db, kb["angle"], has2, fv2, True)
    pre0 = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    exp_init = oracle.search_for_initialization(ka, da, kb, db, 752, 480, pre0.copy(), 100, 0.9, True)
    exp_ap = oracle.hamming_allpairs(da, db)
    exp_kps, exp_desc = ctx["ex"](a)
    errors, handles = [], {}

    def tracking():
        try:
            ex = ORBExtractor(2000, 1.2, 8, 20, 7)
            m = ctx["ORBMatcher"](0.9, True)
            handles["tracking"] = m._h
            for _ in range(25):
                k, d = ex(a)
                assert k.tobytes() == exp_kps.tobytes() and np.array_equal(d, exp_desc)
                pre = pre0.copy()
                n, m12 = m.SearchForInitialization(FV(ka, da, 752, 480), FV(kb, db, 752, 480), pre, 100)
                assert n == exp_init[0] and np.array_equal(m12, exp_init[1]) and np.array_equal(pre, exp_init[2])
            ex.close()
        except BaseException as e:       # noqa: BLE001 - reported by the main thread
            errors.append(("tracking", repr(e)))

    def local_mapping():
        try:
            m = ctx["ORBMatcher"](0.6, True)
            handles["mapping"] = m._h
            for _ in range(25):
                n, m12 = m.SearchForTriangulation(da, ka["angle"], has1, fv1, db, kb["angle"], has2, fv2)
                assert n == exp_tri[0] and np.array_equal(m12, exp_tri[1])
                got = m.hamming_allpairs(da, db)
                assert all(np.array_equal(x, y) for x, y in zip(got, exp_ap))
        except BaseException as e:       # noqa: BLE001
            errors.append(("local mapping", repr(e)))

    threads = [threading.Thread(target=tracking), threading.Thread(target=local_mapping)]
    for t in threads: t.start()
    for t in threads: t.join(timeout=300)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads)
    main = matcher_mod._handle()
    addr = lambda h: h.value if hasattr(h, "value") else int(h)
    assert len({addr(handles["tracking"]), addr(handles["mapping"]), addr(main)}) == 3       # every thread got its own handle


def test_matchers_reproduce_the_reference_matcher_golden(ctx):
    """Every matcher entry point of the C-ABI against tests/golden/matcher_ref.npz: outputs of the reference's own
    modules/ORB/ORBMatcher.cpp compiled verbatim (tools/gen_golden_matcher.py) — match counts, assignments, matches12 and the
    updated vecPreMatched bit for bit."""
    from golden_matcher import replay
    M, FV = ctx["ORBMatcher"], ctx["FrameView"]

    def init(ka, da, kb, db, w, h, pre, window, ratio, orient):
        n, m12 = M(ratio, orient).SearchForInitialization(FV(ka, da, w, h), FV(kb, db, w, h), pre, window)
        return n, m12, pre

    def proj(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, kb, db, w, h, occ, orient):
        return M(0.6, orient).SearchByProjection(q_u, q_v, q_r, q_l, q_a, q_d, q_valid, FV(kb, db, w, h), occ)

    def local(q_u, q_v, q_r, q_l, q_d, q_valid, kb, db, w, h, occ, ratio):
        return M(ratio, True).SearchLocalPoints(q_u, q_v, q_r, q_l, q_d, q_valid, FV(kb, db, w, h), occ)

    def tri(d1, a1, has1, fv1, d2, a2, has2, fv2, orient):
        return M(0.6, orient).SearchForTriangulation(d1, a1, has1, fv1, d2, a2, has2, fv2)

    def bow(d1, a1, valid1, fv1, d2, a2, occ2, fv2, ratio, orient):
        return M(ratio, orient).SearchByBow(d1, a1, valid1, fv1, d2, a2, occ2, fv2)

    def fuse(u, v, radius, level, qd, valid, ka, da, w, h, square_sigmas):
        return M().SearchFuse(FV(ka, da, w, h), u, v, radius, level, qd, valid)

    assert replay(dict(search_for_initialization=init, search_by_projection=proj, search_local_points=local, search_for_triangulation=tri,
                       search_by_bow=bow, search_fuse=fuse)) == 13


def test_device_resident_frames_give_the_host_results(ctx, oracle):
    """orbfe_frame (uploaded once, or wrapped around the device outputs of orbfe_extract_batch_device + orbfe_frame_postprocess_device):
    SearchForInitialization, SearchByProjection (twice on the same pair, as Tracking.cpp:284-296 does) and the local-map search equal
    the host-array entry points and the oracle."""
    import torch
    from monoorbslam3_b200 import DeviceFrame, Camera, frame_postprocess_device, KP_DTYPE
    M = ctx["ORBMatcher"]; ka, da, kb, db, W, H = ctx["ka"], ctx["da"], ctx["kb"], ctx["db"], ctx["w"], ctx["h"]
    f1h, f2h = ctx["FrameView"](ka, da, W, H), ctx["FrameView"](kb, db, W, H)
    f1, f2 = DeviceFrame.upload(ka, da, W, H), DeviceFrame.upload(kb, db, W, H)
    assert f1.num_kps == len(ka)
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    for window, ratio, orient in ((100, 0.9, True), (40, 0.7, False)):
        m = M(ratio, orient)
        p1, p2 = pre.copy(), pre.copy()
        n1, m1 = m.SearchForInitialization(f1, f2, p1, window)
        n2, m2 = m.SearchForInitialization(f1h, f2h, p2, window)
        on, om, op = oracle.search_for_initialization(ka, da, kb, db, W, H, pre.copy(), window, ratio, orient)
        assert n1 == n2 == on and np.array_equal(m1, m2) and np.array_equal(m1, om) and np.array_equal(p1, p2) and np.array_equal(p1, op)
    rng = np.random.default_rng(3)
    nq = len(ka)
    q_u = (ka["x"] - 7 + rng.normal(0, 1.0, nq)).astype(np.float32); q_v = (ka["y"] - 3 + rng.normal(0, 1.0, nq)).astype(np.float32)
    q_valid = (rng.random(nq) < 0.85).astype(np.uint8); occ = (rng.random(len(kb)) < 0.1).astype(np.uint8)
    for th in (15, 30):                                          # the second, wider search of Tracking.cpp:294 on the same frame object
        q_r = (np.float32(th) * ka["size"]).astype(np.float32)
        m = M(0.9, True)
        got = m.SearchByProjection(q_u, q_v, q_r, ka["octave"], ka["angle"], da, q_valid, f2, occ)
        host = m.SearchByProjection(q_u, q_v, q_r, ka["octave"], ka["angle"], da, q_valid, f2h, occ)
        exp = oracle.search_by_projection(q_u, q_v, q_r, ka["octave"], ka["angle"], da, q_valid, kb, db, W, H, occ, True)
        assert got[0] == host[0] == exp[0] and np.array_equal(got[1], host[1]) and np.array_equal(got[1], exp[1]) and got[0] > 300
    sf = np.array([ctx["ex"].getScaleFactor(int(l)) for l in ka["octave"]], np.float32)
    q_r = (np.float32(2) * np.float32(4.0) * sf).astype(np.float32)
    m = M(0.8, True)
    got = m.SearchLocalPoints(q_u, q_v, q_r, ka["octave"], da, q_valid, f2, occ)
    exp = oracle.search_local_points(q_u, q_v, q_r, ka["octave"], da, q_valid, kb, db, W, H, occ, 0.8)
    assert got[0] == exp[0] and np.array_equal(got[1], exp[1]) and got[0] > 300

    # no host round trip between extractor and matcher: extract on the device, post-process on the device, wrap, search
    ex = ctx["ex"]
    a, b = ctx["synth"].shifted_pair(480, 752, 1000)
    dev = torch.device("cuda", 0)
    frames = torch.from_numpy(np.stack([a, b])).to(dev)
    cap = 2300
    d_kps = torch.zeros((2, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.zeros((2, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(2, dtype=torch.int32, device=dev)
    ex.extract_batch_device(frames, 2, 480, 752, d_kps, d_desc, cap, d_n, sync=True)
    cam = Camera(458.654, 457.296, 367.215, 248.375, [0.0, 0.0, 0.0, 0.0])        # no distortion: the undistorted key points are the raw ones
    from monoorbslam3_b200 import grid_size
    gc, gr = grid_size(752, 480)
    d_un = torch.zeros_like(d_kps); d_goff = torch.zeros((2, gc * gr + 1), dtype=torch.int32, device=dev); d_gidx = torch.zeros((2, cap), dtype=torch.int32, device=dev)
    frame_postprocess_device(ex, cam, d_kps, d_un, d_n, 2, cap, 752, 480, d_goff, d_gidx)
    n_host = d_n.cpu().numpy()
    assert n_host[0] == len(ka) and n_host[1] == len(kb)
    w1 = DeviceFrame.wrap(d_un[0], d_desc[0], int(n_host[0]), 752, 480, d_goff[0], d_gidx[0], handle=ex._h)
    w2 = DeviceFrame.wrap(d_un[1], d_desc[1], int(n_host[1]), 752, 480, handle=ex._h)          # grid built by the wrap
    m = M(0.9, True, handle=ex._h)
    p1 = pre.copy()
    n1, m1 = m.SearchForInitialization(w1, w2, p1, 100)
    on, om, op = oracle.search_for_initialization(ka, da, kb, db, W, H, pre.copy(), 100, 0.9, True)
    assert n1 == on and np.array_equal(m1, om) and np.array_equal(p1, op)
    for f in (f1, f2, w1, w2):
        f.close()


def test_init_search_parallel_resolve_and_its_fallback(ctx, oracle, monkeypatch):
    """SearchForInitialization resolves as a fixed-point iteration (k_resolve_init_par); when a slot collects more acceptors than the
    kernel keeps it hands over to the sequential kernel.  Both paths, and the forced-serial one, equal the oracle; duplicated
    descriptors produce steal chains (several acceptors per slot)."""
    M = ctx["ORBMatcher"]; ka, da, kb, W, H = ctx["ka"], ctx["da"], ctx["kb"], ctx["w"], ctx["h"]
    rng = np.random.default_rng(21)
    db = ctx["db"].copy()
    lvl0 = np.nonzero(kb["octave"] == 0)[0]
    for _ in range(60):                                          # many near-identical level-0 descriptors in frame 2: contested slots
        src = rng.choice(lvl0); dst = rng.choice(lvl0, 4)
        db[dst] = db[src]; db[dst[0], rng.integers(0, 32)] ^= 1
    f1, f2 = ctx["FrameView"](ka, da, W, H), ctx["FrameView"](kb, db, W, H)
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    for window, ratio, orient in ((100, 0.9, True), (200, 0.95, True), (60, 0.8, False)):
        exp = oracle.search_for_initialization(ka, da, kb, db, W, H, pre.copy(), window, ratio, orient)
        for env in ({}, {"ORBFE_INIT_ACC_LIMIT": "1"}, {"ORBFE_SERIAL_RESOLVE": "1"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            p = pre.copy()
            n, m12 = M(ratio, orient).SearchForInitialization(f1, f2, p, window)
            for k in env:
                monkeypatch.delenv(k)
            assert n == exp[0] and np.array_equal(m12, exp[1]) and np.array_equal(p, exp[2]), (window, env)


@pytest.mark.parametrize("n_frames,cap", [(5, 700), (3, 257), (2, 128), (7, 401), (4, 512), (6, 384), (3, 1024)])   # the last three: blocks on tile boundaries, train tiles of padding rows only
def test_allpairs_on_extractor_slabs(ctx, oracle, n_frames, cap):
    """orbfe_hamming_allpairs_slab_device: a key-frame window matched against itself straight from fixed-capacity slabs (counts on the
    device, padding rows in between): equals the compacted table searched with per-row exclusion of the own frame."""
    import torch
    m = ctx["ORBMatcher"]()
    rng = np.random.default_rng(n_frames * 1000 + cap)
    pool = np.concatenate([ctx["da"], ctx["db"]])
    counts = rng.integers(cap // 2, cap + 1, n_frames).astype(np.int32)
    counts[0] = cap; counts[-1] = max(1, cap // 3)                         # a full block and a short one
    slab = rng.integers(0, 256, (n_frames, cap, 32), dtype=np.uint8)       # padding rows hold garbage that must never match
    for f in range(n_frames):
        slab[f, :counts[f]] = pool[rng.integers(0, len(pool), counts[f])]
    dev = torch.device("cuda", 0)
    d_slab = torch.from_numpy(slab).to(dev); d_n = torch.from_numpy(counts).to(dev)
    bi = torch.zeros(n_frames * cap, dtype=torch.int32, device=dev); bd = torch.zeros_like(bi); sd = torch.zeros_like(bi)
    m.hamming_allpairs_slab_device(d_slab, d_n, n_frames, cap, bi, bd, sd)
    bi, bd, sd = (t.cpu().numpy().reshape(n_frames, cap) for t in (bi, bd, sd))
    table = np.concatenate([slab[f, :counts[f]] for f in range(n_frames)])
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    blk = np.repeat(np.arange(n_frames), counts)
    excl = np.stack([starts[blk], starts[blk + 1]], 1).astype(np.int32)
    ebi, ebd, esd = m.hamming_allpairs(table, table, excl)                 # parity-tested against numpy / the oracle above
    row_of = np.concatenate([f * cap + np.arange(counts[f]) for f in range(n_frames)])          # compact index -> slab row
    for f in range(n_frames):
        sl = slice(starts[f], starts[f + 1])
        exp_idx = np.where(ebi[sl] >= 0, row_of[np.maximum(ebi[sl], 0)], -1)
        assert np.array_equal(bi[f, :counts[f]], exp_idx) and np.array_equal(bd[f, :counts[f]], ebd[sl]) and np.array_equal(sd[f, :counts[f]], esd[sl]), f
        assert (bi[f, counts[f]:] == -1).all() and (bd[f, counts[f]:] == 257).all()
