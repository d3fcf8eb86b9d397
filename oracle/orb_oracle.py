"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so) and of the verbatim reference build
(oracle/_ref/libref_orb*.so).  TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg — never from the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
CORNER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])


def build(force=False):
    """(Re)build the oracle and, when /root/reference exists, the verbatim reference libraries."""
    so = os.path.join(HERE, "liborb_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "orb_oracle.c")):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "all"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liborb_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        u8p, i32p, f32p, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_resize_linear_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t]
        L.orc_gaussian_blur7_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t]
        L.orc_fast9_16.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, vp, C.c_int]
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_create.restype = vp
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extractor_quota.argtypes = [vp, C.c_int]
        L.orc_extractor_scale.argtypes = [vp, C.c_int]
        L.orc_extractor_scale.restype = C.c_float
        L.orc_extract.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_size_t, vp, u8p, C.c_int]
        L.orc_level_image.argtypes = [vp, C.c_int, i32p, i32p, vp]
        L.orc_level_image.restype = vp
        L.orc_level_blurred.argtypes = [vp, C.c_int, i32p, i32p, vp]
        L.orc_level_blurred.restype = vp
        L.orc_level_candidates.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_level_keypoints.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_distribute_octree.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
        L.orc_ic_angle.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int]
        L.orc_ic_angle.restype = C.c_float
        L.orc_brief_descriptor.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_float, u8p]
        L.orc_descriptor_distance.argtypes = [u8p, u8p]
        L.orc_grid_build.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_grid_build.restype = vp
        L.orc_grid_destroy.argtypes = [vp]
        L.orc_features_in_area.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
        L.orc_search_for_initialization.argtypes = [vp, u8p, C.c_int, vp, u8p, C.c_int, C.c_int, C.c_int, f32p, i32p,
                                                    C.c_int, C.c_float, C.c_int]
        L.orc_search_by_projection.argtypes = [f32p, f32p, f32p, i32p, f32p, u8p, u8p, C.c_int, vp, u8p, C.c_int, C.c_int,
                                               C.c_int, u8p, i32p, C.c_int]
        L.orc_search_local_points.argtypes = [f32p, f32p, f32p, i32p, u8p, u8p, C.c_int, vp, u8p, C.c_int, C.c_int, C.c_int,
                                              u8p, i32p, C.c_float]
        L.orc_search_fuse.argtypes = [f32p, f32p, f32p, i32p, u8p, u8p, C.c_int, vp, u8p, C.c_int, C.c_int, C.c_int, f32p, i32p, i32p]
        L.orc_compute_descriptors.argtypes = [u8p, i32p, C.c_int, i32p]
        L.orc_search_by_bow.argtypes = [u8p, f32p, u8p, C.c_int, i32p, i32p, i32p, C.c_int, u8p, f32p, u8p, C.c_int, i32p, i32p, i32p, C.c_int,
                                        i32p, C.c_float, C.c_int]
        L.orc_search_for_triangulation.argtypes = [u8p, f32p, u8p, C.c_int, i32p, i32p, i32p, C.c_int,
                                                   u8p, f32p, u8p, C.c_int, i32p, i32p, i32p, C.c_int, i32p, C.c_int]
        L.orc_compute_three_maxima.argtypes = [i32p, C.c_int, i32p, i32p, i32p]
        L.orc_hamming_allpairs.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p, i32p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = _c(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def gaussian_blur7(src):
    src = _c(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian_blur7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9_16(img, threshold, nms=True):
    img = _c(img, np.uint8)
    out = np.empty(img.size // 2 + 16, CORNER_DTYPE)
    n = lib().orc_fast9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), out.size)
    assert n >= 0
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(y), float(x)))


def descriptor_distance(a, b):
    a = _c(a, np.uint8); b = _c(b, np.uint8)
    return int(lib().orc_descriptor_distance(_p(a), _p(b)))


# ---------------------------------------------------------------- extractor
class Extractor:
    """CPU oracle of ORBExtractor (ORBExtractor.h:29-39), canonical octree tie-break."""

    def __init__(self, n_features=1000, scale_factor=1.2, n_levels=8, ini_th=20, min_th=10):
        self.n_levels = n_levels
        self.n_features = n_features
        self._h = lib().orc_extractor_create(n_features, scale_factor, n_levels, ini_th, min_th)
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def quota(self, level):
        return lib().orc_extractor_quota(self._h, level)

    def scale(self, level):
        return lib().orc_extractor_scale(self._h, level)

    def __call__(self, img):
        img = _c(img, np.uint8)
        cap = self.n_features + 64 * self.n_levels + 256        # a level may return more than its quota when its root nodes alone exceed it
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orc_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        assert n >= 0, "oracle capacity"
        return kps[:n].copy(), desc[:n].copy()

    def last_stage_ms(self):
        """Wall-clock split of the last call on this thread: (pyramid, FAST per cell, blur, everything else) in ms."""
        out = (C.c_double * 4)()
        lib().orc_last_stage_ms(out)
        return tuple(out)

    def _img(self, fn, level):
        w, h, st = C.c_int(), C.c_int(), C.c_size_t()
        p = fn(self._h, level, C.byref(w), C.byref(h), C.byref(st))
        if not p:
            return None
        buf = (C.c_uint8 * (st.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, st.value)[:, :w.value].copy()

    def level_image(self, level):
        return self._img(lib().orc_level_image, level)

    def level_blurred(self, level):
        return self._img(lib().orc_level_blurred, level)

    def level_candidates(self, level):
        n = lib().orc_level_candidates(self._h, level, None, 0)
        out = np.empty(max(n, 1), CORNER_DTYPE)
        lib().orc_level_candidates(self._h, level, _p(out), n)
        return out[:n]

    def level_keypoints(self, level):
        n = lib().orc_level_keypoints(self._h, level, None, 0)
        out = np.empty(max(n, 1), KP_DTYPE)
        lib().orc_level_keypoints(self._h, level, _p(out), n)
        return out[:n]


def distribute_octree(cands, min_x, max_x, min_y, max_y, n_features):
    cands = _c(cands, CORNER_DTYPE)
    out = np.empty(max(len(cands), 1), np.int32)
    n = lib().orc_distribute_octree(_p(cands), len(cands), min_x, max_x, min_y, max_y, n_features, _p(out), out.size)
    return out[:n].copy()


def ic_angle(img, x, y):
    img = _c(img, np.uint8)
    return float(lib().orc_ic_angle(_p(img), img.strides[0], x, y))


def brief_descriptor(blurred, x, y, angle):
    blurred = _c(blurred, np.uint8)
    d = np.empty(32, np.uint8)
    lib().orc_brief_descriptor(_p(blurred), blurred.strides[0], x, y, float(angle), _p(d))
    return d


# ---------------------------------------------------------------- matcher
def features_in_area(kps, img_w, img_h, x, y, r, min_level=-1, max_level=-1, strict=False):
    kps = _c(kps, KP_DTYPE)
    g = lib().orc_grid_build(_p(kps), len(kps), img_w, img_h)
    out = np.empty(max(len(kps), 1), np.int32)
    n = lib().orc_features_in_area(g, _p(kps), x, y, r, min_level, max_level, int(strict), _p(out), out.size)
    lib().orc_grid_destroy(g)
    return out[:n].copy()


def search_for_initialization(kps1, desc1, kps2, desc2, img_w, img_h, prematched, window=100, nn_ratio=0.9,
                              check_orientation=True):
    kps1 = _c(kps1, KP_DTYPE); kps2 = _c(kps2, KP_DTYPE)
    desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8)
    pre = _c(prematched, np.float32).copy()
    m12 = np.empty(max(len(kps1), 1), np.int32)
    n = lib().orc_search_for_initialization(_p(kps1), _p(desc1), len(kps1), _p(kps2), _p(desc2), len(kps2), img_w, img_h,
                                            _p(pre), _p(m12), window, nn_ratio, int(check_orientation))
    return n, m12[:len(kps1)].copy(), pre


def search_by_projection(q_u, q_v, q_radius, q_level, q_angle, q_desc, q_valid, kps2, desc2, img_w, img_h, occupied,
                         check_orientation=True):
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_radius = _c(q_radius, np.float32)
    q_level = _c(q_level, np.int32); q_angle = _c(q_angle, np.float32); q_desc = _c(q_desc, np.uint8)
    q_valid = _c(q_valid, np.uint8); kps2 = _c(kps2, KP_DTYPE); desc2 = _c(desc2, np.uint8); occupied = _c(occupied, np.uint8)
    assigned = np.empty(max(len(kps2), 1), np.int32)
    n = lib().orc_search_by_projection(_p(q_u), _p(q_v), _p(q_radius), _p(q_level), _p(q_angle), _p(q_desc), _p(q_valid),
                                       len(q_u), _p(kps2), _p(desc2), len(kps2), img_w, img_h, _p(occupied), _p(assigned),
                                       int(check_orientation))
    return n, assigned[:len(kps2)].copy()


def search_local_points(q_u, q_v, q_radius, q_level, q_desc, q_valid, kps2, desc2, img_w, img_h, occupied, nn_ratio=0.8):
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_radius = _c(q_radius, np.float32)
    q_level = _c(q_level, np.int32); q_desc = _c(q_desc, np.uint8)
    q_valid = _c(q_valid, np.uint8); kps2 = _c(kps2, KP_DTYPE); desc2 = _c(desc2, np.uint8); occupied = _c(occupied, np.uint8)
    assigned = np.empty(max(len(kps2), 1), np.int32)
    n = lib().orc_search_local_points(_p(q_u), _p(q_v), _p(q_radius), _p(q_level), _p(q_desc), _p(q_valid), len(q_u),
                                      _p(kps2), _p(desc2), len(kps2), img_w, img_h, _p(occupied), _p(assigned), nn_ratio)
    return n, assigned[:len(kps2)].copy()


def search_for_triangulation(desc1, angle1, has_mp1, fv1, desc2, angle2, has_mp2, fv2, check_orientation=False):
    """fv = (node_ids ascending, offsets, indices) as int32 arrays."""
    desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8)
    angle1 = _c(angle1, np.float32); angle2 = _c(angle2, np.float32)
    has_mp1 = _c(has_mp1, np.uint8); has_mp2 = _c(has_mp2, np.uint8)
    a = [_c(x, np.int32) for x in fv1]; b = [_c(x, np.int32) for x in fv2]
    m12 = np.empty(max(len(desc1), 1), np.int32)
    n = lib().orc_search_for_triangulation(_p(desc1), _p(angle1), _p(has_mp1), len(desc1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]),
                                           _p(desc2), _p(angle2), _p(has_mp2), len(desc2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                           _p(m12), int(check_orientation))
    return n, m12[:len(desc1)].copy()


def search_by_bow(desc1, angle1, valid1, fv1, desc2, angle2, occupied2, fv2, nn_ratio=0.7, check_orientation=True):
    """ORBMatcher::SearchByBow (ORBMatcher.cpp:118-201) on flat arrays; fv = (node_ids ascending, offsets, indices)."""
    desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8)
    angle1 = _c(angle1, np.float32); angle2 = _c(angle2, np.float32)
    valid1 = _c(valid1, np.uint8); occupied2 = _c(occupied2, np.uint8)
    a = [_c(x, np.int32) for x in fv1]; b = [_c(x, np.int32) for x in fv2]
    asg = np.empty(max(len(desc2), 1), np.int32)
    n = lib().orc_search_by_bow(_p(desc1), _p(angle1), _p(valid1), len(desc1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]),
                                _p(desc2), _p(angle2), _p(occupied2), len(desc2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                _p(asg), float(nn_ratio), int(check_orientation))
    return n, asg[:len(desc2)].copy()


def search_fuse(q_u, q_v, q_radius, q_level, q_desc, q_valid, kps1, desc1, img_w, img_h, square_sigmas):
    """Search half of the fuse SearchByProjection(KeyFrame, mapPoints) (ORBMatcher.cpp:524-571) -> (numMatch, best_idx1[nq], best_dist[nq])."""
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_radius = _c(q_radius, np.float32); q_level = _c(q_level, np.int32)
    q_desc = _c(q_desc, np.uint8); q_valid = _c(q_valid, np.uint8); kps1 = _c(kps1, KP_DTYPE); desc1 = _c(desc1, np.uint8)
    ss = _c(square_sigmas, np.float32)
    bi = np.empty(max(len(q_u), 1), np.int32); bd = np.empty(max(len(q_u), 1), np.int32)
    n = lib().orc_search_fuse(_p(q_u), _p(q_v), _p(q_radius), _p(q_level), _p(q_desc), _p(q_valid), len(q_u), _p(kps1), _p(desc1), len(kps1),
                              int(img_w), int(img_h), _p(ss), _p(bi), _p(bd))
    return n, bi[:len(q_u)].copy(), bd[:len(q_u)].copy()


def compute_descriptors(desc, off):
    """MapPoint::computeDescriptor (MapPoint.cpp:103-152) for groups of descriptor rows [off[g], off[g+1]) -> best index within each group."""
    desc = _c(desc, np.uint8); off = _c(off, np.int32)
    best = np.empty(max(len(off) - 1, 1), np.int32)
    lib().orc_compute_descriptors(_p(desc), _p(off), len(off) - 1, _p(best))
    return best[:len(off) - 1].copy()


def compute_three_maxima(counts):
    counts = _c(counts, np.int32)
    i1, i2, i3 = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    lib().orc_compute_three_maxima(_p(counts), len(counts), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def hamming_allpairs(q, t):
    q = _c(q, np.uint8); t = _c(t, np.uint8)
    bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32); sd = np.empty(len(q), np.int32)
    lib().orc_hamming_allpairs(_p(q), len(q), _p(t), len(t), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def hamming_window(q, t, cand_offsets, cand_idx):
    """Best / second best of each query over its own candidate list, scanned in list order the way ORBMatcher.cpp:60-72 does
    (`if (dist < bestDist) {second = best; best = dist} else if (dist < second) second = dist`); 257 where there is none."""
    q = _c(q, np.uint8); t = _c(t, np.uint8)
    bi = np.full(len(q), -1, np.int32); bd = np.full(len(q), 257, np.int32); sd = np.full(len(q), 257, np.int32)
    for i in range(len(q)):
        for j in cand_idx[cand_offsets[i]:cand_offsets[i + 1]]:
            d = int(np.unpackbits(q[i] ^ t[j]).sum())
            if d < bd[i]:
                sd[i] = bd[i]; bd[i] = d; bi[i] = j
            elif d < sd[i]:
                sd[i] = d
    return bi, bd, sd


# ---------------------------------------------------------------- verbatim reference build (oracle/_ref)
class ReferenceExtractor:
    """The reference's own ORBExtractor.cpp compiled verbatim against oracle/cvshim (see oracle/ref_harness.cpp).
    canonical=True uses the stable-by-size tie-break at ORBExtractor.cpp:757; False keeps the pointer tie-break."""

    def __init__(self, n_features=1000, scale_factor=1.2, n_levels=8, ini_th=20, min_th=10, canonical=True):
        name = "libref_orb_canon.so" if canonical else "libref_orb.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self._L = C.CDLL(path)
        self._L.ref_extractor_create.restype = C.c_void_p
        self._L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        self._L.ref_extractor_destroy.argtypes = [C.c_void_p]
        self._L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        self._L.ref_extract_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self._L.ref_extract_mt.restype = C.c_double
        self.n_features, self.n_levels = n_features, n_levels
        self._h = self._L.ref_extractor_create(n_features, scale_factor, n_levels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img):
        img = _c(img, np.uint8)
        cap = self.n_features + 4 * self.n_levels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self._L.ref_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy()

    def time_batch(self, frames, threads):
        """Run the reference extractor over frames[B,H,W] with `threads` host threads (one private extractor each);
        returns (seconds, keypoint counts)."""
        frames = _c(frames, np.uint8)
        counts = np.zeros(frames.shape[0], np.int32)
        sec = self._L.ref_extract_mt(self._h, _p(frames), frames.shape[0], frames.shape[2], frames.shape[1], threads, _p(counts))
        return float(sec), counts
