"""ctypes wrapper of oracle/_ref/libref_twoview.so: the reference's own Frontend/TwoViewReconstruction.cpp compiled VERBATIM against a
name-level Eigen stand-in (recipe in oracle/Makefile, harness oracle/twoview_harness.cpp).  TEST INFRASTRUCTURE ONLY — see
oracle/orb_oracle.py's header.  Pins the RANSAC scoring (CheckHomography :226-288, CheckFundamental :290-345)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libref_twoview.so")
_lib = None


def available():
    return os.path.exists(PATH)


def _l():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        _lib.ref_check_homography.restype = C.c_float
        _lib.ref_check_fundamental.restype = C.c_float
    return _lib


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def check_homography(H21, pts1, pts2, sigma):
    """-> (score float32, inliers bool[n], H12 float32[3,3] the function worked with)."""
    H21 = _f(H21).reshape(3, 3); p1 = _f(pts1).reshape(-1, 2); p2 = _f(pts2).reshape(-1, 2)
    inl = np.zeros(max(len(p1), 1), np.uint8); H12 = np.zeros((3, 3), np.float32)
    s = _l().ref_check_homography(H21.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), len(p1), C.c_float(sigma),
                                  inl.ctypes.data_as(C.c_void_p), H12.ctypes.data_as(C.c_void_p))
    return np.float32(s), inl[:len(p1)].astype(bool), H12


def check_fundamental(F21, pts1, pts2, sigma):
    F21 = _f(F21).reshape(3, 3); p1 = _f(pts1).reshape(-1, 2); p2 = _f(pts2).reshape(-1, 2)
    inl = np.zeros(max(len(p1), 1), np.uint8)
    s = _l().ref_check_fundamental(F21.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), len(p1), C.c_float(sigma),
                                   inl.ctypes.data_as(C.c_void_p))
    return np.float32(s), inl[:len(p1)].astype(bool)
