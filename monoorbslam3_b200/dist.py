"""ctypes mirror of include/orbfe_dist.h (liborbfe_dist.so): the multi-GPU entry points for a single-process host — one host thread
and one NCCL communicator per GPU inside the library.  (bench.py and the torchrun tools use one process per GPU with torch.distributed
instead: see sharding.py.)  No CPU fallback: creating a group without CUDA devices raises."""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import KP_DTYPE

LIB_PATH = os.path.join(_capi.HERE, "lib", "liborbfe_dist.so")
_lib = None
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t


def lib():
    global _lib
    if _lib is None:
        _capi.lib()                                            # liborbfe.so first: the dist library links against it
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s not built: run `python -m monoorbslam3_b200.build`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.orbfe_dist_init.restype = _i; L.orbfe_dist_init.argtypes = [C.POINTER(_capi.Config), _i, _vp, C.POINTER(_vp)]
        L.orbfe_dist_destroy.restype = None; L.orbfe_dist_destroy.argtypes = [_vp]
        L.orbfe_dist_size.restype = _i; L.orbfe_dist_size.argtypes = [_vp]
        L.orbfe_dist_last_error.restype = C.c_char_p; L.orbfe_dist_last_error.argtypes = [_vp]
        L.orbfe_extract_batch_sharded.restype = _i
        L.orbfe_extract_batch_sharded.argtypes = [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp]
        L.orbfe_allpairs_sharded.restype = _i
        L.orbfe_allpairs_sharded.argtypes = [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]
        _lib = L
    return _lib


class DistGroup:
    """orbfe_dist_init(cfg, n_gpus, devices): ORBExtractor constructor arguments + the GPUs of the group."""

    def __init__(self, n_gpus, nFeatures=1000, scaleFactor=1.2, nLevels=8, iniThFast=20, minThFast=10, devices=None, max_batch=64):
        self._lib = lib()
        cfg = _capi.Config(int(nFeatures), float(scaleFactor), int(nLevels), int(iniThFast), int(minThFast), 0, int(max_batch), 0)
        dev = np.ascontiguousarray(devices, np.int32) if devices is not None else None
        self._d = C.c_void_p()
        rc = self._lib.orbfe_dist_init(C.byref(cfg), int(n_gpus), _capi.ptr(dev), C.byref(self._d))
        if rc != _capi.ORBFE_OK:
            raise _capi.OrbfeError(rc, (self._lib.orbfe_dist_last_error(None) or b"?").decode())
        self.n_features, self.n_levels = int(nFeatures), int(nLevels)

    def close(self):
        if getattr(self, "_d", None):
            self._lib.orbfe_dist_destroy(self._d)
            self._d = None

    __del__ = close

    def size(self):
        return int(self._lib.orbfe_dist_size(self._d))

    def _check(self, rc):
        if rc != _capi.ORBFE_OK:
            raise _capi.OrbfeError(rc, (self._lib.orbfe_dist_last_error(self._d) or b"?").decode())

    def extract_batch(self, frames, cap=None):
        """frames [B, H, W] uint8 (host) -> (n[B], kps[B, cap] KP_DTYPE, desc[B, cap, 32]); frame blocks are spread over the GPUs."""
        frames = np.ascontiguousarray(frames, np.uint8)
        B, H, W = frames.shape
        cap = int(cap or self.n_features + 40 * self.n_levels + 64)
        n = np.zeros(B, np.int32); kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8)
        self._check(self._lib.orbfe_extract_batch_sharded(self._d, _capi.ptr(frames), B, W, H, W, W * H, _capi.ptr(kps), _capi.ptr(desc), cap, _capi.ptr(n)))
        return n, kps, desc

    def hamming_allpairs(self, q, t, excl=None):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        ex = np.ascontiguousarray(excl, np.int32).reshape(-1, 2) if excl is not None else None
        bi = np.zeros(len(q), np.int32); bd = np.zeros(len(q), np.int32); sd = np.zeros(len(q), np.int32)
        self._check(self._lib.orbfe_allpairs_sharded(self._d, _capi.ptr(q), len(q), _capi.ptr(t), len(t), _capi.ptr(ex), _capi.ptr(bi), _capi.ptr(bd), _capi.ptr(sd)))
        return bi, bd, sd
