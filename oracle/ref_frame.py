"""ctypes wrapper of oracle/_ref/libref_frame.so: the reference's own BasicObject/Frame.cpp (with its own Frame.h) compiled VERBATIM
after oracle/frameshim/prelude.h replaced the Eigen / g2o dependent headers by stand-ins (recipe in oracle/Makefile, harness
oracle/frame_harness.cpp).  TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header.  Pins the 40-px grid of Frame::Frame
(Frame.cpp:32-51) and Frame::getFeaturesInArea (:97-127)."""
import ctypes as C
import os

import numpy as np

from .orb_oracle import KP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libref_frame.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        _lib.ref_frame_features_in_area.restype = C.c_int
        _lib.ref_frame_grid.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def features_in_area(kps, img_w, img_h, qx, qy, qr, qmin, qmax):
    """-> (offsets[nq + 1], indices) of Frame::getFeaturesInArea for every query, and (GRID_COLS, GRID_ROWS)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    qmin = np.ascontiguousarray(qmin, np.int32); qmax = np.ascontiguousarray(qmax, np.int32)
    nq = len(qx)
    cap = max(len(kps) * max(nq, 1), 1)
    idx = np.empty(min(cap, 1 << 26), np.int32); off = np.zeros(nq + 1, np.int32)
    cols, rows = C.c_int(), C.c_int()
    total = lib().ref_frame_features_in_area(_p(kps), len(kps), int(img_w), int(img_h), _p(qx), _p(qy), _p(qr), _p(qmin), _p(qmax), nq,
                                             _p(idx), len(idx), _p(off), C.byref(cols), C.byref(rows))
    assert total >= 0
    return off, idx[:total].copy(), (cols.value, rows.value)


KF_PATH = os.path.join(HERE, "_ref", "libref_keyframe.so")
_kf = None


def keyframe_available():
    return os.path.exists(KF_PATH)


def keyframe_features_in_area(kps, img_w, img_h, qx, qy, qr, qmin, qmax):
    """KeyFrame::getFeaturesInArea (KeyFrame.cpp:181-211, strict `<`) of the reference's own KeyFrame.cpp compiled verbatim
    (oracle/keyframe_harness.cpp): -> (offsets[nq + 1], indices)."""
    global _kf
    if _kf is None:
        _kf = C.CDLL(KF_PATH)
        _kf.ref_keyframe_features_in_area.restype = C.c_int
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    qmin = np.ascontiguousarray(qmin, np.int32); qmax = np.ascontiguousarray(qmax, np.int32)
    nq = len(qx)
    idx = np.empty(min(max(len(kps) * max(nq, 1), 1), 1 << 26), np.int32); off = np.zeros(nq + 1, np.int32)
    total = _kf.ref_keyframe_features_in_area(_p(kps), len(kps), int(img_w), int(img_h), _p(qx), _p(qy), _p(qr), _p(qmin), _p(qmax), nq,
                                              _p(idx), len(idx), _p(off))
    assert total >= 0
    return off, idx[:total].copy()


def grid(kps, img_w, img_h):
    """-> (grid_off[cols * rows + 1], grid_idx, (cols, rows)): Frame::grid flattened with cell = cx * rows + cy."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    off = np.zeros(((img_w + 39) // 40) * ((img_h + 39) // 40) + 8, np.int32); idx = np.empty(max(len(kps), 1), np.int32)
    cols, rows = C.c_int(), C.c_int()
    total = lib().ref_frame_grid(_p(kps), len(kps), int(img_w), int(img_h), _p(off), len(off), _p(idx), C.byref(cols), C.byref(rows))
    assert total >= 0
    return off[:cols.value * rows.value + 1].copy(), idx[:total].copy(), (cols.value, rows.value)
