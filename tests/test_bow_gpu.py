"""GPU: DBoW2 vocabulary descent (csrc/orbfe_bow.cu) through the C-ABI against golden vectors of the reference's own vendored
DBoW2 (tests/golden/dbow_ref.npz, thirdParty/DBoW2 compiled verbatim) and against the restatement of
TemplatedVocabulary::transform (oracle/bow.py) on further synthetic vocabulary trees — word ids, node ids, weights, the
FeatureVector and the BowVector bit-exact; then SearchByBow on feature vectors produced by the device."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from monoorbslam3_b200 import ORBExtractor, ORBVocabulary, ORBMatcher, synth
    from oracle import bow
    ex = ORBExtractor(1500, 1.2, 8, 20, 7)
    a, b = synth.shifted_pair(480, 752, 1000)
    ka, da = ex(a); kb, db = ex(b)
    yield ex, ORBVocabulary, ORBMatcher, bow, ka, da, kb, db
    ex.close()


@pytest.mark.parametrize("name", ["k10L3", "k6L4", "k3L6", "k10L2_root"])
def test_transform_matches_reference_dbow2_golden(env, name):
    ex, ORBVocabulary = env[0], env[1]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dbow_ref.npz"))
    c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    gv = ORBVocabulary(ex, int(c["k"]), int(c["L"]), c["parent"], c["leaf"], c["desc"], c["weight"])
    assert gv.n_words == int(c["n_words"])
    wid, nid, w, (fnode, foff, fidx) = gv.transform(c["feats"], int(c["levelsup"]))
    assert np.array_equal(wid, c["word_id"]) and np.array_equal(nid, c["node_id"]) and np.array_equal(w, c["word_weight"])
    assert np.array_equal(fnode, c["fv_node"]) and np.array_equal(foff, c["fv_off"]) and np.array_equal(fidx, c["fv_idx"])
    bid, bval = gv.bow_vector(wid, w)
    assert np.array_equal(bid, c["bow_id"]) and np.array_equal(bval, c["bow_val"])
    gv.close()


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (6, 4, 2), (3, 6, 4), (20, 2, 4), (33, 2, 1)])
def test_transform_matches_restatement(env, k, L, levelsup):
    ex, ORBVocabulary, _, bow, ka, da, kb, db = env
    kk, LL, parent, leaf, desc, w = bow.synthetic_vocabulary(k, L, seed=k * 10 + L)
    ov = bow.Vocabulary(kk, LL, parent, leaf, desc, w)
    rng = np.random.default_rng(k)
    feats = np.concatenate([da[:400], desc[rng.integers(1, ov.n, 200)] ^ np.packbits(rng.random((200, 256)) < 0.04, axis=1),
                            desc[rng.integers(1, ov.n, 50)]])                         # real descriptors, noisy node descriptors, exact ties
    gv = ORBVocabulary(ex, kk, LL, parent, leaf, desc, w)
    assert gv.n_words == ov.n_words
    wid, nid, ww, fv = gv.transform(feats, levelsup)
    owid, onid, oww, ofv = bow.transform(ov, feats, levelsup)
    assert np.array_equal(wid, owid) and np.array_equal(nid, onid) and np.array_equal(ww, oww)
    for g, o in zip(fv, ofv):
        assert np.array_equal(g, o)
    assert (ww == 0).any() or k * L < 12                                              # stopped words are exercised
    gv.close()


def test_search_by_bow_on_device_feature_vectors(env, oracle):
    ex, ORBVocabulary, ORBMatcher, bow, ka, da, kb, db = env
    kk, LL, parent, leaf, desc, w = bow.synthetic_vocabulary(8, 3, seed=5, stop_fraction=0.0)
    gv = ORBVocabulary(ex, kk, LL, parent, leaf, desc, w)
    _, _, _, fv1 = gv.transform(da, 2); _, _, _, fv2 = gv.transform(db, 2)
    valid1 = np.ones(len(da), np.uint8); occ2 = np.zeros(len(db), np.uint8)
    m = ORBMatcher(0.8, True, handle=ex._h)
    n, asg = m.SearchByBow(da, ka["angle"], valid1, fv1, db, kb["angle"], occ2, fv2)
    on, oasg = oracle.search_by_bow(da, ka["angle"], valid1, fv1, db, kb["angle"], occ2, fv2, 0.8, True)
    assert n == on and n > 30 and np.array_equal(asg, oasg)
    gv.close()
