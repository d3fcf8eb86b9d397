"""CPU oracle of the DBoW2 vocabulary descent (TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header).

Restates, on flat numpy arrays, what Frame::computeBow / KeyFrame::computeBow (BasicObject/Frame.cpp:168-178, KeyFrame.cpp:213-223)
ask of the vendored DBoW2 (thirdParty/DBoW2/DBoW2/TemplatedVocabulary.h):
  * loadFromTextFile (:1338-1420): node ids in file order, children appended to their parent in file order, word ids in leaf order;
  * transform(feature, word_id, weight, nid, levelsup) (:1217-1259): descend from the root, at every level take the child with
    the strictly smallest FORB::distance (FORB.cpp:81-101; the first minimum wins), remember the node reached at level
    L - levelsup, stop at a node without children;
  * transform(features, BowVector, FeatureVector, levelsup) (:1127-1172) for TF_IDF weighting: features whose word weight is > 0
    add their weight to the word and their index to the node's list (std::map order = ascending node id, ascending index).
The reference repository ships no vocabulary file (ORBvoc.txt is git-ignored), so the trees are synthetic — written in the
ORBvoc.txt text format (write_text_file) and read back by the reference's own loader.  The restatement is PINNED to the vendored
DBoW2 itself: oracle/Makefile compiles thirdParty/DBoW2 verbatim (oracle/dbow_harness.cpp, cv:: shim, name-only boost stand-ins)
into oracle/_ref/libref_dbow.so; ReferenceVocabulary wraps it, tools/gen_golden_dbow.py stores its outputs in
tests/golden/dbow_ref.npz, and tests/test_oracle_bow_ref.py compares this file with both."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


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


def bow_vector(word_id, weight):
    """The BowVector of transform(features, v, fv, levelsup) (:1127-1172) for TF_IDF weighting and L1_NORM scoring (ORBvoc.txt's
    header "10 6 0 0"): weights accumulate per word in feature order (BowVector::addWeight), then BowVector::normalize(L1) divides
    by the sum of absolute values taken in std::map (ascending word id) order.  -> (word ids ascending, values)."""
    acc = {}
    for w, v in zip(np.asarray(word_id).tolist(), np.asarray(weight, np.float64).tolist()):
        if v > 0:                                                   # :1156 "not stopped"
            acc[w] = acc.get(w, 0.0) + v
    ids = sorted(acc)
    norm = 0.0
    for w in ids:
        norm += abs(acc[w])
    vals = [acc[w] / norm for w in ids] if norm > 0.0 else [acc[w] for w in ids]
    return np.array(ids, np.int32), np.array(vals, np.float64)


def write_text_file(path, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
    """The ORBvoc.txt format loadFromTextFile (:1338-1420) parses: "k L scoring weighting", then one line per node in id order:
    parent id, leaf flag, the 32 descriptor bytes in decimal, the weight.  No trailing newline: the loader's `while(!f.eof())`
    would otherwise append a node for the empty last line."""
    lines = ["%d %d %d %d" % (k, L, scoring, weighting)]
    for i in range(1, len(parent)):
        lines.append("%d %d %s %s" % (parent[i], 1 if is_leaf[i] else 0, " ".join(str(int(b)) for b in desc[i]), repr(float(weight[i]))))
    with open(path, "w") as f:
        f.write("\n".join(lines))


class ReferenceVocabulary:
    """The reference's vendored DBoW2 (TemplatedVocabulary<FORB::TDescriptor, FORB>, ORBVocabulary.h:12) compiled verbatim."""

    def __init__(self, path):
        so = os.path.join(HERE, "_ref", "libref_dbow.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        L = self._L = C.CDLL(so)
        L.ref_voc_load.restype = C.c_void_p; L.ref_voc_load.argtypes = [C.c_char_p]
        L.ref_voc_free.argtypes = [C.c_void_p]
        for f in (L.ref_voc_words, L.ref_voc_k, L.ref_voc_depth):
            f.argtypes = [C.c_void_p]; f.restype = C.c_int
        L.ref_voc_transform_each.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                        C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ref_voc_transform.restype = C.c_int
        self._v = L.ref_voc_load(path.encode())
        if not self._v:
            raise RuntimeError("the reference loader rejected " + path)
        self.n_words, self.k, self.L = L.ref_voc_words(self._v), L.ref_voc_k(self._v), L.ref_voc_depth(self._v)

    def close(self):
        if self._v:
            self._L.ref_voc_free(self._v); self._v = None

    def transform_each(self, descs, levelsup):
        d = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32); n = len(d)
        wid = np.zeros(n, np.int32); nid = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
        self._L.ref_voc_transform_each(self._v, d.ctypes.data, n, levelsup, wid.ctypes.data, nid.ctypes.data, w.ctypes.data)
        return wid, nid, w

    def transform(self, descs, levelsup):
        """-> BowVector (ids, values) and FeatureVector (node ids, offsets, indices), both in std::map order."""
        d = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32); n = len(d)
        bid = np.zeros(max(n, 1), np.int32); bval = np.zeros(max(n, 1), np.float64)
        fnode = np.zeros(max(n, 1), np.int32); foff = np.zeros(n + 2, np.int32); fidx = np.zeros(max(n, 1), np.int32)
        nb, nn, ni = C.c_int(), C.c_int(), C.c_int()
        rc = self._L.ref_voc_transform(self._v, d.ctypes.data, n, levelsup, bid.ctypes.data, bval.ctypes.data, len(bid), C.byref(nb),
                                       fnode.ctypes.data, foff.ctypes.data, len(fnode), C.byref(nn), fidx.ctypes.data, len(fidx), C.byref(ni))
        assert rc == 0
        return (bid[:nb.value], bval[:nb.value]), (fnode[:nn.value], foff[:nn.value + 1], fidx[:ni.value])
