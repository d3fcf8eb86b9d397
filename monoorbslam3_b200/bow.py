"""Host mirror of ORBVocabulary::transform (thirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1172, 1217-1259) over the C-ABI:
the tree descent runs in liborbfe.so on the GPU (csrc/orbfe_bow.cu)."""
import ctypes as C

import numpy as np

from . import _capi


class ORBVocabulary:
    """Built from the columns of ORBvoc.txt (loadFromTextFile, :1338-1420): for node i = 1..n-1 in file order its parent id, leaf
    flag, 32-byte descriptor and weight (index 0 = root, ignored)."""

    def __init__(self, extractor, k, L, parent, is_leaf, desc, weight):
        self._lib = _capi.lib()
        self._ex = extractor
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(len(parent), 32); weight = np.ascontiguousarray(weight, np.float64)
        self.k, self.L = int(k), int(L)
        v = C.c_void_p()
        _capi.check(extractor._h, self._lib.orbfe_vocab_create(extractor._h, self.k, self.L, len(parent), _capi.ptr(parent), _capi.ptr(is_leaf), _capi.ptr(desc),
                                                               _capi.ptr(weight), C.byref(v)))
        self._v = v

    def close(self):
        if self._v:
            self._lib.orbfe_vocab_destroy(self._v); self._v = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_words(self):
        return self._lib.orbfe_vocab_words(self._v)

    def transform(self, descriptors, levelsup=4):
        """-> (word_id[n], node_id[n], weight[n], feature_vector = (node ids ascending, CSR offsets, feature indices))."""
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = len(d)
        wid = np.zeros(max(n, 1), np.int32); nid = np.zeros(max(n, 1), np.int32); w = np.zeros(max(n, 1), np.float64)
        fid = np.zeros(max(n, 1), np.int32); foff = np.zeros(n + 1, np.int32); fidx = np.zeros(max(n, 1), np.int32)
        nn = C.c_int()
        _capi.check(self._ex._h, self._lib.orbfe_vocab_transform(self._v, _capi.ptr(d), n, int(levelsup), _capi.ptr(wid), _capi.ptr(nid), _capi.ptr(w),
                                                                 _capi.ptr(fid), _capi.ptr(foff), _capi.ptr(fidx), C.byref(nn)))
        k = nn.value
        return wid[:n], nid[:n], w[:n], (fid[:k].copy(), foff[:k + 1].copy(), fidx[:foff[k]].copy())

    @staticmethod
    def bow_vector(word_id, weight):
        """The adapter's half of transform(features, BowVector, FeatureVector, levelsup) (:1150-1166) for ORBvoc.txt's TF_IDF /
        L1_NORM header: BowVector::addWeight in feature order (stopped words, weight 0, left out), then BowVector::normalize(L1)
        in std::map order.  -> (word ids ascending, values)."""
        acc = {}
        for w, v in zip(np.asarray(word_id).tolist(), np.asarray(weight, np.float64).tolist()):
            if v > 0:
                acc[w] = acc.get(w, 0.0) + v
        ids = sorted(acc)
        norm = 0.0
        for w in ids:
            norm += abs(acc[w])
        vals = [acc[w] / norm for w in ids] if norm > 0.0 else [acc[w] for w in ids]
        return np.array(ids, np.int32), np.array(vals, np.float64)
