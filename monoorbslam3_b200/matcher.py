"""Host-side mirror of the reference's ORBMatcher (modules/ORB/ORBMatcher.h:12-52) on top of the C-ABI.

The reference's methods take Frame/KeyFrame objects; here a frame is anything with `key_points` (KP_DTYPE array),
`descriptors` ([N,32] uint8) and `width`/`height` — see FrameView.  The adapters do what the C++ adapters of INTEGRATION.md do:
flatten the objects, call the C-ABI, write the results back in the reference's order."""
import ctypes as C
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import KP_DTYPE

TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30      # ORBMatcher.cpp:13-15


@dataclass
class FrameView:
    """The part of Frame/KeyFrame the matcher reads (BasicObject/Frame.h): undistorted key points, descriptors, image size,
    and the map-point slots (None = empty)."""
    key_points: np.ndarray
    descriptors: np.ndarray
    width: int
    height: int
    map_points: list = field(default_factory=list)

    @property
    def num_kps(self):
        return len(self.key_points)


class DeviceFrame:
    """orbfe_frame: a frame whose key points, descriptors and 40-px grid stay in HBM between matcher calls (Tracking calls the matcher
    several times per frame: Tracking.cpp:284-296, 412-425).  Built from host arrays (upload) or from device buffers
    (wrap: torch tensors / raw pointers, e.g. one frame's slab of ORBExtractor.extract_batch_device + frame_postprocess_device)."""

    def __init__(self, handle, ptr, n, width, height, keep=()):
        self._h, self._f, self.num_kps, self.width, self.height, self._keep = handle, ptr, int(n), int(width), int(height), keep
        self._lib = _capi.lib()

    @classmethod
    def upload(cls, key_points, descriptors, width, height, handle=None, device=0):
        h = handle if handle is not None else _handle(device)
        k = _c(key_points, KP_DTYPE); d = _c(descriptors, np.uint8)
        f = C.c_void_p()
        _capi.check(h, _capi.lib().orbfe_frame_upload(h, _capi.ptr(k), _capi.ptr(d), len(k), int(width), int(height), C.byref(f)))
        return cls(h, f, len(k), width, height)

    @classmethod
    def wrap(cls, d_kps, d_desc, n, width, height, d_grid_off=None, d_grid_idx=None, handle=None, device=0):
        h = handle if handle is not None else _handle(device)
        f = C.c_void_p()
        _capi.check(h, _capi.lib().orbfe_frame_wrap_device(h, _capi.ptr(d_kps), _capi.ptr(d_desc), int(n), int(width), int(height), _capi.ptr(d_grid_off),
                                                           _capi.ptr(d_grid_idx), C.byref(f)))
        return cls(h, f, n, width, height, keep=(d_kps, d_desc, d_grid_off, d_grid_idx))

    def close(self):
        if getattr(self, "_f", None):
            self._lib.orbfe_frame_destroy(self._f)
            self._f = None

    __del__ = close


_tls = threading.local()


def _handle(device=0):
    """One handle (= one CUDA stream + scratch) per host thread and device, like the C++ adapter's thread_local holder: the
    reference runs ORBMatcher on the tracking and the local-mapping thread concurrently (System.cpp:55)."""
    handles = getattr(_tls, "handles", None)
    if handles is None:
        handles = _tls.handles = {}
    if device not in handles:
        handles[device] = _capi.create(1000, 1.2, 8, 20, 7, device, 1, 0)
    return handles[device]


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class ORBMatcher:
    """ORBMatcher(nnRatio=0.6, checkOrientation=True) — ORBMatcher.h:14."""

    def __init__(self, nnRatio=0.6, checkOrientation=True, handle=None, device=0):
        self.nn_ratio = float(nnRatio)
        self.be_check_orientation = bool(checkOrientation)
        self._h = handle if handle is not None else _handle(device)
        self._lib = _capi.lib()

    # ---- static int DescriptorDistance(a, b) — ORBMatcher.cpp:17-31
    def DescriptorDistance(self, a, b):
        a = _c(a, np.uint8).reshape(1, 32); b = _c(b, np.uint8).reshape(1, 32)
        return int(self.descriptor_distances(a, b, [0], [0])[0])

    def descriptor_distances(self, a, b, ia, ib):
        a = _c(a, np.uint8); b = _c(b, np.uint8); ia = _c(ia, np.int32); ib = _c(ib, np.int32)
        out = np.zeros(len(ia), np.int32)
        _capi.check(self._h, self._lib.orbfe_descriptor_distance(self._h, _capi.ptr(a), len(a), _capi.ptr(b), len(b), _capi.ptr(ia), _capi.ptr(ib),
                                                                 len(ia), _capi.ptr(out)))
        return out

    def popc_peak(self):
        """Measured 32-bit popc throughput of the device in 10^9/s (the matching roofline denominator)."""
        v = C.c_double()
        _capi.check(self._h, self._lib.orbfe_popc_peak(self._h, C.byref(v)))
        return v.value

    def imma_peak(self):
        """Measured int8 tensor-core (mma.sync m16n8k32) throughput in 10^9 descriptor pairs/s: the roofline denominator of the
        large all-pairs searches."""
        v = C.c_double()
        _capi.check(self._h, self._lib.orbfe_imma_peak(self._h, C.byref(v)))
        return v.value

    # ---- brute force best / second best (BASELINE configs 4/5)
    def hamming_allpairs(self, q, t, excl=None):
        """Best index / best distance / second-best distance of every query row over all train rows.  excl ([nq, 2] int32, optional):
        train indices [lo, hi) query i skips — its own key frame's block when a key-frame window is matched against itself."""
        q = _c(q, np.uint8); t = _c(t, np.uint8)
        bi = np.zeros(len(q), np.int32); bd = np.zeros(len(q), np.int32); sd = np.zeros(len(q), np.int32)
        ex = None
        if excl is not None:
            ex = _c(excl, np.int32).reshape(-1, 2)
            if len(ex) != len(q):
                raise ValueError("excl must have one [lo, hi) pair per query")
        _capi.check(self._h, self._lib.orbfe_hamming_allpairs_excl(self._h, _capi.ptr(q), len(q), _capi.ptr(t), len(t), _capi.ptr(ex), _capi.ptr(bi),
                                                                   _capi.ptr(bd), _capi.ptr(sd)))
        return bi, bd, sd

    # ---- best / second best over caller-supplied candidate lists (CSR)
    def hamming_window(self, q, t, cand_offsets, cand_idx):
        q = _c(q, np.uint8); t = _c(t, np.uint8)
        off = _c(cand_offsets, np.int32); idx = _c(cand_idx, np.int32)
        if len(off) != len(q) + 1:
            raise ValueError("cand_offsets must have len(q) + 1 entries")
        if len(q) and len(idx) < off[-1]:
            raise ValueError("cand_idx shorter than cand_offsets[-1]")
        bi = np.zeros(len(q), np.int32); bd = np.zeros(len(q), np.int32); sd = np.zeros(len(q), np.int32)
        _capi.check(self._h, self._lib.orbfe_hamming_window(self._h, _capi.ptr(q), len(q), _capi.ptr(t), len(t), _capi.ptr(off), _capi.ptr(idx),
                                                            _capi.ptr(bi), _capi.ptr(bd), _capi.ptr(sd)))
        return bi, bd, sd

    def hamming_allpairs_device(self, d_q, nq, d_t, nt, d_bi, d_bd, d_sd, stream=None, sync=True, d_excl=None):
        _capi.check(self._h, self._lib.orbfe_hamming_allpairs_excl_device(self._h, _capi.ptr(d_q), nq, _capi.ptr(d_t), nt, _capi.ptr(d_excl), _capi.ptr(d_bi),
                                                                          _capi.ptr(d_bd), _capi.ptr(d_sd), C.c_void_p(stream) if stream else None, int(sync)))

    def hamming_allpairs_slab_device(self, d_desc, d_n, n_frames, cap, d_bi, d_bd, d_sd, stream=None, sync=True):
        """A key-frame window matched against itself straight from the extractor's slabs (orbfe_hamming_allpairs_slab_device)."""
        _capi.check(self._h, self._lib.orbfe_hamming_allpairs_slab_device(self._h, _capi.ptr(d_desc), _capi.ptr(d_n), int(n_frames), int(cap), _capi.ptr(d_bi),
                                                                          _capi.ptr(d_bd), _capi.ptr(d_sd), C.c_void_p(stream) if stream else None, int(sync)))

    # ---- int SearchForInitialization(frame1, frame2, vecPreMatched, matches12, windowSize=100) — ORBMatcher.cpp:33-116
    def SearchForInitialization(self, frame1, frame2, vecPreMatched, windowSize=100):
        """Returns (numMatches, matches12); vecPreMatched ([n1,2] float32) is updated in place like the reference's reference argument."""
        pre = _c(vecPreMatched, np.float32).reshape(-1, 2).copy()
        n = C.c_int()
        if isinstance(frame1, DeviceFrame) or isinstance(frame2, DeviceFrame):
            if not (isinstance(frame1, DeviceFrame) and isinstance(frame2, DeviceFrame)):
                raise TypeError("SearchForInitialization: both frames must be DeviceFrame objects, or neither")
            m12 = np.full(max(frame1.num_kps, 1), -1, np.int32)
            _capi.check(self._h, self._lib.orbfe_search_for_initialization_f(self._h, frame1._f, frame2._f, _capi.ptr(pre), _capi.ptr(m12), int(windowSize),
                                                                             self.nn_ratio, int(self.be_check_orientation), C.byref(n)))
            vecPreMatched[...] = pre.reshape(np.shape(vecPreMatched))
            return n.value, m12[:frame1.num_kps]
        k1 = _c(frame1.key_points, KP_DTYPE); k2 = _c(frame2.key_points, KP_DTYPE)
        d1 = _c(frame1.descriptors, np.uint8); d2 = _c(frame2.descriptors, np.uint8)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        _capi.check(self._h, self._lib.orbfe_search_for_initialization(self._h, _capi.ptr(k1), _capi.ptr(d1), len(k1), _capi.ptr(k2), _capi.ptr(d2), len(k2),
                                                                       frame2.width, frame2.height, _capi.ptr(pre), _capi.ptr(m12), int(windowSize),
                                                                       self.nn_ratio, int(self.be_check_orientation), C.byref(n)))
        vecPreMatched[...] = pre.reshape(np.shape(vecPreMatched))
        return n.value, m12[:len(k1)]

    # ---- SearchByProjection(lastFrame|lastKF, curFrame, th) — ORBMatcher.cpp:203-348, after the adapter projected the map points
    def SearchByProjection(self, q_u, q_v, q_radius, q_level, q_angle, q_desc, q_valid, curFrame, occupied):
        """Returns (numMatch, assigned) where assigned[j] = query index written into curFrame.map_points[j] or -1."""
        args = [_c(q_u, np.float32), _c(q_v, np.float32), _c(q_radius, np.float32), _c(q_level, np.int32), _c(q_angle, np.float32),
                _c(q_desc, np.uint8), _c(q_valid, np.uint8)]
        occ = _c(occupied, np.uint8)
        if isinstance(curFrame, DeviceFrame):
            assigned = np.full(max(curFrame.num_kps, 1), -1, np.int32)
            n = C.c_int()
            _capi.check(self._h, self._lib.orbfe_search_by_projection_f(self._h, *[_capi.ptr(a) for a in args], len(args[0]), curFrame._f, _capi.ptr(occ),
                                                                        _capi.ptr(assigned), int(self.be_check_orientation), C.byref(n)))
            return n.value, assigned[:curFrame.num_kps]
        k2 = _c(curFrame.key_points, KP_DTYPE); d2 = _c(curFrame.descriptors, np.uint8)
        assigned = np.full(max(len(k2), 1), -1, np.int32)
        n = C.c_int()
        _capi.check(self._h, self._lib.orbfe_search_by_projection(self._h, *[_capi.ptr(a) for a in args], len(args[0]), _capi.ptr(k2), _capi.ptr(d2), len(k2),
                                                                  curFrame.width, curFrame.height, _capi.ptr(occ), _capi.ptr(assigned),
                                                                  int(self.be_check_orientation), C.byref(n)))
        return n.value, assigned[:len(k2)]

    # ---- SearchByProjection(frame, mapPoints, th) — ORBMatcher.cpp:350-415
    def SearchLocalPoints(self, q_u, q_v, q_radius, q_level, q_desc, q_valid, frame, occupied):
        args = [_c(q_u, np.float32), _c(q_v, np.float32), _c(q_radius, np.float32), _c(q_level, np.int32), _c(q_desc, np.uint8), _c(q_valid, np.uint8)]
        occ = _c(occupied, np.uint8)
        if isinstance(frame, DeviceFrame):
            assigned = np.full(max(frame.num_kps, 1), -1, np.int32)
            n = C.c_int()
            _capi.check(self._h, self._lib.orbfe_search_local_points_f(self._h, *[_capi.ptr(a) for a in args], len(args[0]), frame._f, _capi.ptr(occ),
                                                                       _capi.ptr(assigned), self.nn_ratio, C.byref(n)))
            return n.value, assigned[:frame.num_kps]
        k2 = _c(frame.key_points, KP_DTYPE); d2 = _c(frame.descriptors, np.uint8)
        assigned = np.full(max(len(k2), 1), -1, np.int32)
        n = C.c_int()
        _capi.check(self._h, self._lib.orbfe_search_local_points(self._h, *[_capi.ptr(a) for a in args], len(args[0]), _capi.ptr(k2), _capi.ptr(d2), len(k2),
                                                                 frame.width, frame.height, _capi.ptr(occ), _capi.ptr(assigned), self.nn_ratio, C.byref(n)))
        return n.value, assigned[:len(k2)]

    # ---- int SearchForTriangulation(keyFrame1, keyFrame2, matches12) — ORBMatcher.cpp:417-522
    def SearchByBow(self, desc1, angle1, valid1, fv1, desc2, angle2, occupied2, fv2):
        """ORBMatcher::SearchByBow(KeyFrame, Frame) (ORBMatcher.cpp:118-201).  valid1[i]: key-frame key point i has a good map point;
        occupied2[j]: frame slot j already holds a map point.  -> (numMatch, assigned[n2]: key-frame index placed in slot j or -1)."""
        d1 = _c(desc1, np.uint8); d2 = _c(desc2, np.uint8)
        a1 = _c(angle1, np.float32); a2 = _c(angle2, np.float32); v1 = _c(valid1, np.uint8); o2 = _c(occupied2, np.uint8)
        f1 = [_c(x, np.int32) for x in fv1]; f2 = [_c(x, np.int32) for x in fv2]
        asg = np.full(max(len(d2), 1), -1, np.int32)
        n = C.c_int()
        _capi.check(self._h, self._lib.orbfe_search_by_bow(self._h, _capi.ptr(d1), _capi.ptr(a1), _capi.ptr(v1), len(d1), _capi.ptr(f1[0]), _capi.ptr(f1[1]),
                                                           _capi.ptr(f1[2]), len(f1[0]), _capi.ptr(d2), _capi.ptr(a2), _capi.ptr(o2), len(d2), _capi.ptr(f2[0]),
                                                           _capi.ptr(f2[1]), _capi.ptr(f2[2]), len(f2[0]), _capi.ptr(asg), float(self.nn_ratio),
                                                           int(self.be_check_orientation), C.byref(n)))
        return n.value, asg[:len(d2)]

    def SearchFuse(self, keyframe, q_u, q_v, q_radius, q_level, q_desc, q_valid, square_sigmas=None):
        """Search half of the fuse SearchByProjection(KeyFrame, mapPoints, Map*, th) (ORBMatcher.cpp:524-571): per projected map point the
        best key-frame key point (or -1) and its distance.  `keyframe` is a FrameView of the key frame (undistorted key points).
        square_sigmas = ORBExtractor.getSquareSigmas() of the extractor that produced the key points (the chi-square gate of :564 is
        5.991 * square_sigmas[octave]); without it the table of the matcher's own handle (scale 1.2, 8 levels by default) is used."""
        u = _c(q_u, np.float32); v = _c(q_v, np.float32); r = _c(q_radius, np.float32); lv = _c(q_level, np.int32)
        qd = _c(q_desc, np.uint8); qv = _c(q_valid, np.uint8)
        bi = np.full(max(len(u), 1), -1, np.int32); bd = np.zeros(max(len(u), 1), np.int32)
        n = C.c_int()
        kp = _c(keyframe.key_points, KP_DTYPE); kd = _c(keyframe.descriptors, np.uint8)
        if square_sigmas is None:
            _capi.check(self._h, self._lib.orbfe_search_fuse(self._h, _capi.ptr(u), _capi.ptr(v), _capi.ptr(r), _capi.ptr(lv), _capi.ptr(qd), _capi.ptr(qv), len(u),
                                                             _capi.ptr(kp), _capi.ptr(kd), keyframe.num_kps, keyframe.width, keyframe.height,
                                                             _capi.ptr(bi), _capi.ptr(bd), C.byref(n)))
        else:
            s2 = _c(square_sigmas, np.float32)
            _capi.check(self._h, self._lib.orbfe_search_fuse_sigma(self._h, _capi.ptr(u), _capi.ptr(v), _capi.ptr(r), _capi.ptr(lv), _capi.ptr(qd), _capi.ptr(qv),
                                                                   len(u), _capi.ptr(kp), _capi.ptr(kd), keyframe.num_kps, keyframe.width, keyframe.height,
                                                                   _capi.ptr(s2), len(s2), _capi.ptr(bi), _capi.ptr(bd), C.byref(n)))
        return n.value, bi[:len(u)], bd[:len(u)]

    def compute_descriptors(self, desc, group_off):
        """MapPoint::computeDescriptor (MapPoint.cpp:103-152) for a batch of map points: index (within each group of observation
        descriptors) of the descriptor with the least median distance to the rest."""
        d = _c(desc, np.uint8); off = _c(group_off, np.int32)
        best = np.full(max(len(off) - 1, 1), -1, np.int32)
        _capi.check(self._h, self._lib.orbfe_compute_descriptors(self._h, _capi.ptr(d), _capi.ptr(off), len(off) - 1, _capi.ptr(best)))
        return best[:len(off) - 1]

    def SearchForTriangulation(self, desc1, angle1, has_mp1, fv1, desc2, angle2, has_mp2, fv2):
        """fv = (node ids ascending, CSR offsets, key-point indices): the DBoW2 FeatureVector of a key frame."""
        d1 = _c(desc1, np.uint8); d2 = _c(desc2, np.uint8)
        a1 = _c(angle1, np.float32); a2 = _c(angle2, np.float32); m1 = _c(has_mp1, np.uint8); m2 = _c(has_mp2, np.uint8)
        f1 = [_c(x, np.int32) for x in fv1]; f2 = [_c(x, np.int32) for x in fv2]
        m12 = np.full(max(len(d1), 1), -1, np.int32)
        n = C.c_int()
        _capi.check(self._h, self._lib.orbfe_search_for_triangulation(self._h, _capi.ptr(d1), _capi.ptr(a1), _capi.ptr(m1), len(d1), _capi.ptr(f1[0]),
                                                                      _capi.ptr(f1[1]), _capi.ptr(f1[2]), len(f1[0]), _capi.ptr(d2), _capi.ptr(a2), _capi.ptr(m2),
                                                                      len(d2), _capi.ptr(f2[0]), _capi.ptr(f2[1]), _capi.ptr(f2[2]), len(f2[0]), _capi.ptr(m12),
                                                                      int(self.be_check_orientation), C.byref(n)))
        return n.value, m12[:len(d1)]
