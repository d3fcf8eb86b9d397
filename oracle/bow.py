"""CPU oracle of the DBoW2 vocabulary descent (TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header).

Restates, on flat numpy arrays, what Frame::computeBow / KeyFrame::computeBow (BasicObject/Frame.cpp:168-178, KeyFrame.cpp:213-223)
ask of the vendored DBoW2 (thirdParty/DBoW2/DBoW2/TemplatedVocabulary.h):
  * loadFromTextFile (:1338-1420): node ids in file order, children appended to their parent in file order, word ids in leaf order;
  * transform(feature, word_id, weight, nid, levelsup) (:1217-1259): descend from the root, at every level take the child with
    the strictly smallest FORB::distance (FORB.cpp:81-101; the first minimum wins), remember the node reached at level
    L - levelsup, stop at a node without children;
  * transform(features, BowVector, FeatureVector, levelsup) (:1127-1172) for TF_IDF weighting: features whose word weight is > 0
    add their weight to the word and their index to the node's list (std::map order = ascending node id, ascending index).
The reference repository ships no vocabulary file (ORBvoc.txt is git-ignored), and the vendored headers need boost::serialization
and cv::FileStorage to compile, neither of which is in this image: parity for this piece is pinned to this restatement only
("parity unpinned" in the strict sense), on synthetic vocabulary trees."""
import numpy as np


class Vocabulary:
    def __init__(self, k, L, parent, is_leaf, desc, weight):
        """parent[i], is_leaf[i], desc[i], weight[i] for node i = 1..n-1 in file order (index 0 is the root and ignored)."""
        n = len(parent)
        self.k, self.L, self.n = k, L, n
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(n, 32)
        self.weight = np.asarray(weight, np.float64)
        self.children = [[] for _ in range(n)]
        self.word_id = np.zeros(n, np.int32)
        n_words = 0
        for i in range(1, n):
            self.children[int(parent[i])].append(i)                 # :1390
            if is_leaf[i] > 0:                                      # :1407-1414
                self.word_id[i] = n_words; n_words += 1
        self.n_words = n_words


def distance(a, b):
    return int(np.unpackbits(a ^ b).sum())                         # FORB::distance


def transform_one(voc, feature, levelsup):
    nid_level = voc.L - levelsup
    nid = 0                                                         # root if nid_level <= 0 (:1224)
    final, level = 0, 0
    while True:
        level += 1
        ch = voc.children[final]
        final = ch[0]
        best = distance(feature, voc.desc[final])
        for c in ch[1:]:
            d = distance(feature, voc.desc[c])
            if d < best:
                best, final = d, c
        if level == nid_level:
            nid = final
        if not voc.children[final]:
            break
    return int(voc.word_id[final]), nid, float(voc.weight[final])


def transform(voc, descs, levelsup):
    """-> word_id[n], node_id[n], weight[n], and the FeatureVector as (node ids ascending, offsets, indices)."""
    descs = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32)
    n = len(descs)
    wid = np.zeros(n, np.int32); nid = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
    for i in range(n):
        wid[i], nid[i], w[i] = transform_one(voc, descs[i], levelsup)
    keep = np.nonzero(w > 0)[0]                                     # :1156 "not stopped"
    ids = np.unique(nid[keep])
    off = [0]; idx = []
    for v in ids:
        idx.extend(keep[nid[keep] == v].tolist()); off.append(len(idx))
    return wid, nid, w, (ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32))


def synthetic_vocabulary(k, L, seed, stop_fraction=0.02):
    """A full k-ary tree of depth L with random 256-bit node descriptors derived from their parent (so that descents are not
    degenerate), idf-like positive leaf weights and a few stopped (weight 0) words; breadth-first file order like ORBvoc.txt."""
    rng = np.random.default_rng(seed)
    parent = [0]; is_leaf = [0]; desc = [np.zeros(32, np.uint8)]; weight = [0.0]
    frontier = [0]
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            for _ in range(k):
                i = len(parent)
                flips = np.packbits(rng.random(256) < (0.5 if level == 1 else 0.18)).astype(np.uint8)
                parent.append(p); is_leaf.append(1 if level == L else 0)
                desc.append(desc[p] ^ flips if level > 1 else flips)
                weight.append(0.0 if level < L else (0.0 if rng.random() < stop_fraction else float(rng.uniform(0.5, 9.0))))
                nxt.append(i)
        frontier = nxt
    return k, L, np.array(parent, np.int32), np.array(is_leaf, np.uint8), np.stack(desc), np.array(weight, np.float64)
