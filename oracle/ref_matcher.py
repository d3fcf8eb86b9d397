"""ctypes wrapper of oracle/_ref/libref_matcher.so: the reference's own modules/ORB/ORBMatcher.cpp compiled VERBATIM (recipe in
oracle/Makefile, harness oracle/matcher_harness.cpp, stand-in Frame / KeyFrame / MapPoint headers in oracle/matchshim).
TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header.  The functions take the flat arrays of the restatement's wrappers in
oracle/orb_oracle.py so that tests can call both with the same arguments; where the reference derives a search radius itself
(local map points, fuse) the caller passes the ingredients (th, view cosine, level) instead of the radius."""
import ctypes as C
import os

import numpy as np

from .orb_oracle import KP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libref_matcher.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        for name in ("ref_descriptor_distance", "ref_search_for_initialization", "ref_search_by_projection", "ref_search_local_points",
                     "ref_search_by_bow", "ref_search_for_triangulation", "ref_search_fuse"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def descriptor_distance(a, b):
    a = _c(a, np.uint8); b = _c(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def compute_three_maxima(counts):
    counts = _c(counts, np.int32)
    assert len(counts) == 30                                         # HISTO_LENGTH is a constant of the reference
    i1, i2, i3 = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    lib().ref_compute_three_maxima(_p(counts), 30, C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def search_for_initialization(kps1, desc1, kps2, desc2, img_w, img_h, prematched, window=100, nn_ratio=0.9, check_orientation=True):
    kps1 = _c(kps1, KP_DTYPE); kps2 = _c(kps2, KP_DTYPE); desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8)
    pre = _c(prematched, np.float32).copy()
    m12 = np.empty(max(len(kps1), 1), np.int32)
    n = lib().ref_search_for_initialization(_p(kps1), _p(desc1), len(kps1), _p(kps2), _p(desc2), len(kps2), int(img_w), int(img_h), _p(pre), _p(m12),
                                            int(window), C.c_float(nn_ratio), int(check_orientation))
    return n, m12[:len(kps1)].copy(), pre


def search_by_projection(q_u, q_v, q_radius, q_level, q_angle, q_desc, q_valid, kps2, desc2, img_w, img_h, occupied, check_orientation=True,
                         from_keyframe=False):
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_radius = _c(q_radius, np.float32); q_level = _c(q_level, np.int32)
    q_angle = _c(q_angle, np.float32); q_desc = _c(q_desc, np.uint8); q_valid = _c(q_valid, np.uint8)
    kps2 = _c(kps2, KP_DTYPE); desc2 = _c(desc2, np.uint8); occupied = _c(occupied, np.uint8)
    assigned = np.empty(max(len(kps2), 1), np.int32)
    n = lib().ref_search_by_projection(_p(q_u), _p(q_v), _p(q_radius), _p(q_level), _p(q_angle), _p(q_desc), _p(q_valid), len(q_u), _p(kps2), _p(desc2),
                                       len(kps2), int(img_w), int(img_h), _p(occupied), _p(assigned), int(check_orientation), int(from_keyframe))
    return n, assigned[:len(kps2)].copy()


def search_local_points(q_u, q_v, q_view_cos, q_level, q_desc, q_valid, th, kps2, desc2, img_w, img_h, occupied, nn_ratio=0.8):
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_view_cos = _c(q_view_cos, np.float32); q_level = _c(q_level, np.int32)
    q_desc = _c(q_desc, np.uint8); q_valid = _c(q_valid, np.uint8)
    kps2 = _c(kps2, KP_DTYPE); desc2 = _c(desc2, np.uint8); occupied = _c(occupied, np.uint8)
    assigned = np.empty(max(len(kps2), 1), np.int32)
    n = lib().ref_search_local_points(_p(q_u), _p(q_v), _p(q_view_cos), _p(q_level), _p(q_desc), _p(q_valid), len(q_u), C.c_float(th), _p(kps2), _p(desc2),
                                      len(kps2), int(img_w), int(img_h), _p(occupied), _p(assigned), C.c_float(nn_ratio))
    return n, assigned[:len(kps2)].copy()


def _fv(fv):
    return [_c(x, np.int32) for x in fv]


def search_by_bow(desc1, angle1, valid1, fv1, desc2, angle2, occupied2, fv2, nn_ratio=0.7, check_orientation=True):
    desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8); angle1 = _c(angle1, np.float32); angle2 = _c(angle2, np.float32)
    valid1 = _c(valid1, np.uint8); occupied2 = _c(occupied2, np.uint8); a = _fv(fv1); b = _fv(fv2)
    asg = np.empty(max(len(desc2), 1), np.int32)
    n = lib().ref_search_by_bow(_p(desc1), _p(angle1), _p(valid1), len(desc1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]),
                                _p(desc2), _p(angle2), _p(occupied2), len(desc2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                _p(asg), C.c_float(nn_ratio), int(check_orientation))
    return n, asg[:len(desc2)].copy()


def search_for_triangulation(desc1, angle1, has_mp1, fv1, desc2, angle2, has_mp2, fv2, check_orientation=False):
    desc1 = _c(desc1, np.uint8); desc2 = _c(desc2, np.uint8); angle1 = _c(angle1, np.float32); angle2 = _c(angle2, np.float32)
    has_mp1 = _c(has_mp1, np.uint8); has_mp2 = _c(has_mp2, np.uint8); a = _fv(fv1); b = _fv(fv2)
    m12 = np.empty(max(len(desc1), 1), np.int32)
    n = lib().ref_search_for_triangulation(_p(desc1), _p(angle1), _p(has_mp1), len(desc1), _p(a[0]), _p(a[1]), _p(a[2]), len(a[0]),
                                           _p(desc2), _p(angle2), _p(has_mp2), len(desc2), _p(b[0]), _p(b[1]), _p(b[2]), len(b[0]),
                                           _p(m12), int(check_orientation))
    return n, m12[:len(desc1)].copy()


def search_fuse(q_u, q_v, q_level, q_desc, q_valid, th, kps1, desc1, img_w, img_h):
    q_u = _c(q_u, np.float32); q_v = _c(q_v, np.float32); q_level = _c(q_level, np.int32); q_desc = _c(q_desc, np.uint8); q_valid = _c(q_valid, np.uint8)
    kps1 = _c(kps1, KP_DTYPE); desc1 = _c(desc1, np.uint8)
    bi = np.empty(max(len(q_u), 1), np.int32)
    n = lib().ref_search_fuse(_p(q_u), _p(q_v), _p(q_level), _p(q_desc), _p(q_valid), len(q_u), C.c_float(th), _p(kps1), _p(desc1), len(kps1),
                              int(img_w), int(img_h), _p(bi))
    return n, bi[:len(q_u)].copy()
