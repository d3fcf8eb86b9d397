"""ctypes wrapper of oracle/_ref/libref_mappoint.so: the reference's own BasicObject/MapPoint.cpp (with its own MapPoint.h) compiled
VERBATIM after oracle/mappointshim/prelude.h replaced KeyFrame / Map / ORBMatcher by stand-ins (recipe in oracle/Makefile, harness
oracle/mappoint_harness.cpp).  TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header.  Pins MapPoint::computeDescriptor
(MapPoint.cpp:103-152)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libref_mappoint.so")
_lib = None


def available():
    return os.path.exists(PATH)


def compute_descriptors(desc, off, bad=None):
    """-> (order, n_order, chosen): per group the row indices in the reference's own iteration order (std::map over key-frame
    addresses, bad key frames left out), their count, and the 32-byte descriptor MapPoint::computeDescriptor picked."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); off = np.ascontiguousarray(off, np.int32)
    bad = np.zeros(len(desc), np.uint8) if bad is None else np.ascontiguousarray(bad, np.uint8)
    ng = len(off) - 1
    order = np.full(max(len(desc), 1), -1, np.int32); n_order = np.zeros(max(ng, 1), np.int32); chosen = np.zeros((max(ng, 1), 32), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.ref_compute_descriptors(p(desc), p(bad), p(off), ng, p(order), p(n_order), p(chosen))
    return order, n_order[:ng], chosen[:ng]
