"""The vocabulary-descent restatement (oracle/bow.py) against the reference's own vendored DBoW2: golden vectors produced by
thirdParty/DBoW2 compiled verbatim (tools/gen_golden_dbow.py -> tests/golden/dbow_ref.npz), and, where oracle/_ref/libref_dbow.so
is present, the live library on fresh random inputs.  Word ids, node ids and weights per feature, the BowVector (L1-normalised
doubles, bit-exact) and the FeatureVector in std::map order."""
import os
import tempfile

import numpy as np
import pytest

from oracle import bow

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dbow_ref.npz")
CASES = ["k10L3", "k6L4", "k3L6", "k10L2_root"]


def _case(g, name):
    return {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(name):
    c = _case(np.load(GOLD), name)
    voc = bow.Vocabulary(int(c["k"]), int(c["L"]), c["parent"], c["leaf"], c["desc"], c["weight"])
    assert voc.n_words == int(c["n_words"])
    wid, nid, w, (fnode, foff, fidx) = bow.transform(voc, c["feats"], int(c["levelsup"]))
    assert np.array_equal(wid, c["word_id"]) and np.array_equal(nid, c["node_id"]) and np.array_equal(w, c["word_weight"])
    assert np.array_equal(fnode, c["fv_node"]) and np.array_equal(foff, c["fv_off"]) and np.array_equal(fidx, c["fv_idx"])
    bid, bval = bow.bow_vector(wid, w)
    assert np.array_equal(bid, c["bow_id"]) and np.array_equal(bval, c["bow_val"])          # doubles, bit-exact
    assert abs(bval.sum() - 1.0) < 1e-12 and (w == 0).any()                                  # L1 norm; stopped words are exercised
    if name == "k10L2_root":
        assert fnode.tolist() == [0]                                                         # levelsup >= L: every feature files under the root


def test_restatement_matches_live_reference_build():
    if not os.path.exists(os.path.join(os.path.dirname(bow.__file__), "_ref", "libref_dbow.so")):
        pytest.skip("oracle/_ref/libref_dbow.so not built (needs the reference sources)")
    rng = np.random.default_rng(77)
    for (k, L, levelsup) in [(4, 5, 3), (9, 3, 2), (2, 7, 4), (20, 2, 1)]:
        kk, LL, parent, leaf, desc, weight = bow.synthetic_vocabulary(k, L, seed=k + 100 * L, stop_fraction=0.1)
        feats = np.concatenate([rng.integers(0, 256, (120, 32), dtype=np.uint8), desc[rng.integers(1, len(parent), 80)]])
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            bow.write_text_file(path, kk, LL, parent, leaf, desc, weight)
            rv = bow.ReferenceVocabulary(path)
            rw, rn, rwt = rv.transform_each(feats, levelsup)
            (rbid, rbval), (rfn, rfo, rfi) = rv.transform(feats, levelsup)
            n_words = rv.n_words
            rv.close()
        voc = bow.Vocabulary(kk, LL, parent, leaf, desc, weight)
        wid, nid, w, (fnode, foff, fidx) = bow.transform(voc, feats, levelsup)
        assert voc.n_words == n_words
        assert np.array_equal(wid, rw) and np.array_equal(nid, rn) and np.array_equal(w, rwt)
        assert np.array_equal(fnode, rfn) and np.array_equal(foff, rfo) and np.array_equal(fidx, rfi)
        bid, bval = bow.bow_vector(wid, w)
        assert np.array_equal(bid, rbid) and np.array_equal(bval, rbval)
