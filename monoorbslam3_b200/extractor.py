"""Host-side mirror of the reference's ORBExtractor (modules/ORB/ORBExtractor.h:26-98) on top of the C-ABI.

Same constructor arguments and the same call contract as `ORBExtractor::operator()(image, keyPoints, descriptors)`:
an 8-bit single-channel image in, a key-point array (cv::KeyPoint layout) and an N x 32 uint8 descriptor matrix out."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import KP_DTYPE


class ORBExtractor:
    """ORBExtractor(nFeatures=1000, scaleFactor=1.2, nLevels=8, iniThFast=20, minThFast=10) — ORBExtractor.h:29-30."""

    def __init__(self, nFeatures=1000, scaleFactor=1.2, nLevels=8, iniThFast=20, minThFast=10, device=0, max_batch=1,
                 use_tma=True, keep_stages=False):
        flags = (0 if use_tma else _capi.FLAG_NO_TMA) | (_capi.FLAG_KEEP_STAGES if keep_stages else 0)
        self.n_features, self.n_levels = int(nFeatures), int(nLevels)
        self.ini_th_fast, self.min_th_fast = int(iniThFast), int(minThFast)
        self.device, self.max_batch = int(device), int(max_batch)
        self._h = _capi.create(self.n_features, float(scaleFactor), self.n_levels, self.ini_th_fast, self.min_th_fast,
                               self.device, self.max_batch, flags)
        self._lib = _capi.lib()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbfe_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- static getters of the reference (ORBExtractor.h:44-86)
    def getScaleFactor(self, level=0):
        return float(self._lib.orbfe_scale_factor(self._h, level))

    def getScaleFactors(self):
        return [self.getScaleFactor(l) for l in range(self.n_levels)]

    def getSquareSigmas(self):
        """square_sigmas[level] = scale_factors[level]^2 in float32 (ORBExtractor.cpp:432-436)."""
        sf = np.array(self.getScaleFactors(), np.float32)
        return (sf * sf).astype(np.float32)

    def getMaxScaleFactor(self):
        return self.getScaleFactor(self.n_levels - 1)

    def getLevels(self):
        return self.n_levels

    def getFeaturesPerLevel(self, level):
        return int(self._lib.orbfe_features_per_level(self._h, level))

    def capacity(self):
        """Upper bound of key points per frame for the current geometry (sum of per-level list bounds)."""
        return int(self._lib.orbfe_max_keypoints(self._h))

    def launch_count(self):
        return int(self._lib.orbfe_launch_count(self._h))

    def profile(self, enable=True):
        """Record CUDA events around every stage of every pass (on the launching stream)."""
        _capi.check(self._h, self._lib.orbfe_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        """-> ({stage: accumulated ms}, passes) since the last reset."""
        ms = np.zeros(len(_capi.STAGES), np.float32)
        n = C.c_int()
        _capi.check(self._h, self._lib.orbfe_profile_read(self._h, _capi.ptr(ms), C.byref(n), int(reset)))
        return dict(zip(_capi.STAGES, ms.tolist())), n.value

    # ---- operator()
    def __call__(self, image):
        """image: HxW uint8 (CV_8UC1).  Returns (keypoints[KP_DTYPE], descriptors[N,32] uint8).  An empty image or an image
        without key points returns empty arrays (the reference leaves its outputs untouched, ORBExtractor.cpp:497,512)."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be a 2-D uint8 array (CV_8UC1, ORBExtractor.cpp:499)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.n_features + 40 * self.n_levels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = self._lib.orbfe_extract(self._h, _capi.ptr(image), image.shape[1], image.shape[0], image.strides[0], _capi.ptr(kps),
                                     _capi.ptr(desc), cap, C.byref(n))
        _capi.check(self._h, rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, frames, cap=None, out=None):
        """frames: [B,H,W] uint8 host array (pinned or pageable).  Returns (n[B], kps[B,cap], desc[B,cap,32]) host arrays."""
        if frames.dtype != np.uint8 or frames.ndim != 3:
            raise TypeError("frames must be a [B,H,W] uint8 array")
        if frames.strides[2] != 1 or frames.strides[1] < frames.shape[2]:
            frames = np.ascontiguousarray(frames)
        B, H, W = frames.shape
        if cap is None:
            cap = self.n_features + 40 * self.n_levels + 64
        if out is None:
            out = (np.zeros(B, np.int32), np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8))
        n, kps, desc = out
        rc = self._lib.orbfe_extract_batch(self._h, _capi.ptr(frames), B, W, H, frames.strides[1], frames.strides[0], _capi.ptr(kps),
                                           _capi.ptr(desc), cap, _capi.ptr(n))
        _capi.check(self._h, rc)
        return n, kps, desc

    def extract_batch_submit(self, frames, out, cap=None):
        """Asynchronous extract_batch for streams of batches: enqueues the batch and returns a ticket; `frames` and the arrays of
        `out` = (n[B], kps[B,cap], desc[B,cap,32]) must be pinned host arrays that stay alive until extract_batch_wait(ticket)."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or frames.strides[2] != 1 or frames.strides[1] < frames.shape[2]:
            raise TypeError("frames must be a [B,H,W] uint8 array with unit pixel stride")
        B, H, W = frames.shape
        n, kps, desc = out
        if cap is None:
            cap = kps.shape[1]
        if kps.shape[0] < B or kps.shape[1] != cap or desc.shape[:2] != kps.shape[:2] or len(n) < B:
            raise ValueError("output arrays do not match the batch")
        t = C.c_longlong(-1)
        rc = self._lib.orbfe_extract_batch_submit(self._h, _capi.ptr(frames), B, W, H, frames.strides[1], frames.strides[0], _capi.ptr(kps),
                                                  _capi.ptr(desc), cap, _capi.ptr(n), C.byref(t))
        _capi.check(self._h, rc)
        return t.value

    def extract_batch_wait(self, ticket=-1):
        """Blocks until the batch of `ticket` (default: every submitted batch) has its results in the caller's arrays."""
        _capi.check(self._h, self._lib.orbfe_extract_batch_wait(self._h, ticket))

    def extract_batch_device(self, d_frames, B, H, W, d_kps, d_desc, cap, d_n, row_stride=None, frame_stride=None, stream=None, sync=True):
        """Device-resident variant: arguments are torch CUDA tensors (or raw device pointers as ints)."""
        row_stride = W if row_stride is None else row_stride
        frame_stride = row_stride * H if frame_stride is None else frame_stride
        rc = self._lib.orbfe_extract_batch_device(self._h, _capi.ptr(d_frames), B, W, H, row_stride, frame_stride, _capi.ptr(d_kps),
                                                  _capi.ptr(d_desc), cap, _capi.ptr(d_n), C.c_void_p(stream) if stream else None, int(sync))
        _capi.check(self._h, rc)

    # ---- per-stage outputs of the last pass (parity tests)
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _capi.check(self._h, self._lib.orbfe_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level_image(self, level, frame=0, blurred=False):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        fn = self._lib.orbfe_get_level_blurred if blurred else self._lib.orbfe_get_level_image
        _capi.check(self._h, fn(self._h, frame, level, _capi.ptr(out)))
        return out

    def _packed(self, fn, level, frame):
        n = C.c_int()
        _capi.check(self._h, fn(self._h, frame, level, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        _capi.check(self._h, fn(self._h, frame, level, _capi.ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def level_candidates(self, level, frame=0):
        """FAST candidates (x, y relative to the 19-px border, score) in reference order."""
        return self._packed(self._lib.orbfe_get_level_candidates, level, frame)

    def level_keypoints(self, level, frame=0):
        """Quadtree-selected key points (x, y in level pixels, score) in list order."""
        return self._packed(self._lib.orbfe_get_level_keypoints, level, frame)
