"""Host mirror of the Frame constructor's post-processing (modules/BasicObject/Frame.cpp:22-51) over the C-ABI:
kp.size *= camera->uncertainty(kp.pt), camera->undistortKeyPoints(), and the 40-px grid the matchers index.
The arithmetic runs in liborbfe.so on the GPU (csrc/orbfe_frame.cu); nothing here computes."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import KP_DTYPE, CAMERA_PINHOLE, CAMERA_FISHEYE


class Camera:
    """Camera::create's fields (modules/Sensor/Camera.cpp:27-52): CameraMatrix, Distortion, DistortionModel
    ("radtan" -> Pinhole, "equidistant" -> Fisheye).  `uncertainty_map` is Fisheye::scale_mat (height x width float32)."""

    def __init__(self, fx, fy, cx, cy, dist=(), model="radtan", uncertainty_map=None):
        if model not in ("radtan", "equidistant"):
            raise ValueError("un-recognition distort model: %s" % model)              # Camera.cpp:48
        self.model = CAMERA_PINHOLE if model == "radtan" else CAMERA_FISHEYE
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.dist = [float(v) for v in dist]
        if len(self.dist) > 12:
            raise ValueError("at most 12 distortion coefficients")
        self.uncertainty_map = None if uncertainty_map is None else np.ascontiguousarray(uncertainty_map, np.float32)

    def c_struct(self):
        c = _capi.Camera()
        c.model, c.fx, c.fy, c.cx, c.cy = self.model, self.fx, self.fy, self.cx, self.cy
        for i, v in enumerate(self.dist):
            c.dist[i] = v
        c.n_dist = len(self.dist)
        if self.uncertainty_map is not None:
            c.uncertainty_map = self.uncertainty_map.ctypes.data
            c.uncertainty_h, c.uncertainty_w = self.uncertainty_map.shape
        return c


def grid_size(img_w, img_h):
    cols, rows = C.c_int(), C.c_int()
    rc = _capi.lib().orbfe_grid_size(img_w, img_h, C.byref(cols), C.byref(rows))
    if rc:
        raise ValueError("bad image size")
    return cols.value, rows.value


class FramePost:
    """raw_key_points (size scaled), key_points (undistorted) and the grid of one frame — the members Frame::Frame fills."""

    def __init__(self, raw, un, grid_off, grid_idx, cols, rows):
        self.raw_key_points, self.key_points, self.grid_off, self.grid_idx, self.cols, self.rows = raw, un, grid_off, grid_idx, cols, rows

    def cell(self, cx, cy):
        """grid[cx][cy] (Frame.cpp:43-51)."""
        c = cx * self.rows + cy
        return self.grid_idx[self.grid_off[c]:self.grid_off[c + 1]]


def frame_postprocess(extractor, kps, camera, img_w, img_h):
    """kps: KP_DTYPE array from ORBExtractor.__call__.  Uses the extractor's handle (stream, scratch)."""
    lib = _capi.lib()
    raw = np.ascontiguousarray(kps, KP_DTYPE).copy()
    n = len(raw)
    cols, rows = grid_size(img_w, img_h)
    un = np.zeros(n, KP_DTYPE)
    off = np.zeros(cols * rows + 1, np.int32); idx = np.zeros(max(n, 1), np.int32)
    nin = C.c_int()
    cam = camera.c_struct()
    _capi.check(extractor._h, lib.orbfe_frame_postprocess(extractor._h, C.byref(cam), _capi.ptr(raw), n, img_w, img_h, _capi.ptr(un),
                                                          _capi.ptr(off), _capi.ptr(idx), C.byref(nin)))
    return FramePost(raw, un, off, idx[:nin.value], cols, rows)


def frame_postprocess_device(extractor, camera, d_kps_raw, d_kps_un, d_n, n_frames, cap, img_w, img_h, d_grid_off, d_grid_idx, d_n_in_grid=None,
                             stream=None, sync=True):
    """Device-resident batch variant, chained after ORBExtractor.extract_batch_device (torch CUDA tensors or raw pointers)."""
    cam = camera.c_struct()
    rc = _capi.lib().orbfe_frame_postprocess_device(extractor._h, C.byref(cam), _capi.ptr(d_kps_raw), _capi.ptr(d_kps_un), _capi.ptr(d_n), n_frames, cap,
                                                    img_w, img_h, _capi.ptr(d_grid_off), _capi.ptr(d_grid_idx), _capi.ptr(d_n_in_grid),
                                                    C.c_void_p(stream) if stream else None, int(sync))
    _capi.check(extractor._h, rc)
