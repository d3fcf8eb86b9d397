"""CPU pins of the round-2 fixtures (tools/gen_golden_r2.py): the oracle restatement reproduces what the reference's own
ORBExtractor.cpp / ORBMatcher.cpp (compiled verbatim into oracle/_ref) produced on two real photographs and on BASELINE config 3's
1920x1080 / 8000-feature SearchForInitialization."""
import hashlib
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def c3_inputs(oracle):
    from monoorbslam3_b200 import synth
    g = np.load(os.path.join(GOLD, "matcher_c3_ref.npz"))
    a, b = synth.shifted_pair(1080, 1920, int(g["seed"]))
    ex = oracle.Extractor(8000, 1.2, 8, 20, 7)
    ka, da = ex(a); kb, db = ex(b)
    h = hashlib.sha256()
    for x in (ka, da, kb, db):
        h.update(np.ascontiguousarray(x).tobytes())
    assert h.hexdigest() == str(g["inputs"]), "the oracle extractor no longer reproduces the inputs of the committed fixture"
    return g, a, b, ka, da, kb, db


def test_oracle_extractor_equals_reference_on_photographs(oracle):
    g = np.load(os.path.join(GOLD, "photos_ref.npz"))
    for name in ("china", "flower"):
        kps, desc = oracle.Extractor(int(g["nf_" + name]), 1.2, 8, 20, 7)(g["img_" + name])
        assert kps.tobytes() == g["kps_" + name].tobytes() and np.array_equal(desc, g["desc_" + name]), name


def test_restated_init_search_equals_reference_on_config3(oracle):
    g, _, _, ka, da, kb, db = c3_inputs(oracle)
    pre = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    n, m12, pre2 = oracle.search_for_initialization(ka, da, kb, db, 1920, 1080, pre, 100, 0.9, True)
    assert n == int(g["n"]) and np.array_equal(m12, g["m12"]) and np.array_equal(pre2, g["pre"])
