#!/usr/bin/env python3
"""Golden vectors of the reference's vendored DBoW2 (oracle/_ref/libref_dbow.so = thirdParty/DBoW2 compiled verbatim, see
oracle/Makefile) for the vocabulary descent: synthetic vocabularies written in the ORBvoc.txt text format, loaded by the
reference's loadFromTextFile, and the outputs of its transform() on key-point-like descriptors.  Writes tests/golden/dbow_ref.npz.
Run in the build container (needs /root/reference for the _ref build)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bow

out = {}
cases = [("k10L3", 10, 3, 1, 11), ("k6L4", 6, 4, 2, 12), ("k3L6", 3, 6, 4, 13), ("k10L2_root", 10, 2, 4, 14)]    # name, k, L, levelsup, seed
for name, k, L, levelsup, seed in cases:
    kk, LL, parent, leaf, desc, w = bow.synthetic_vocabulary(k, L, seed, stop_fraction=0.05)
    rng = np.random.default_rng(seed)
    n = len(parent)
    feats = np.concatenate([rng.integers(0, 256, (150, 32), dtype=np.uint8),
                            desc[rng.integers(1, n, 200)] ^ np.packbits(rng.random((200, 256)) < 0.04, axis=1),       # noisy node descriptors
                            desc[rng.integers(1, n, 50)]])                                                             # exact node descriptors: ties
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "voc.txt")
        bow.write_text_file(path, kk, LL, parent, leaf, desc, w)
        rv = bow.ReferenceVocabulary(path)
        assert (rv.k, rv.L) == (kk, LL)
        wid, nid, ww = rv.transform_each(feats, levelsup)
        (bid, bval), (fnode, foff, fidx) = rv.transform(feats, levelsup)
        n_words = rv.n_words
        rv.close()
    for key, val in dict(k=kk, L=LL, levelsup=levelsup, parent=parent, leaf=leaf, desc=desc, weight=w, feats=feats, n_words=n_words,
                         word_id=wid, node_id=nid, word_weight=ww, bow_id=bid, bow_val=bval, fv_node=fnode, fv_off=foff, fv_idx=fidx).items():
        out[name + "/" + key] = np.asarray(val)
    print(name, "nodes", n, "words", n_words, "features", len(feats), "bow entries", len(bid), "fv nodes", len(fnode), "stopped", int((ww == 0).sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dbow_ref.npz"), **out)
print("wrote tests/golden/dbow_ref.npz")
