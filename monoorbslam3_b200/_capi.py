"""ctypes binding of liborbfe.so (include/orbfe.h).  The library is the product: if it is missing or no CUDA device is
present, loading / handle creation raises — there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liborbfe.so")

ORBFE_OK, ORBFE_E_ARG, ORBFE_E_CUDA, ORBFE_E_CAPACITY, ORBFE_E_INTERNAL = 0, -1, -2, -3, -4
ORBFE_MAX_LEVELS = 16
FLAG_NO_TMA, FLAG_KEEP_STAGES = 1, 2
STAGES = ("pyramid", "fast", "quadtree", "blur", "describe")

# cv::KeyPoint layout (7 x 4 bytes), see include/orbfe.h
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


class Config(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("scale_factor", C.c_float), ("n_levels", C.c_int32), ("ini_th_fast", C.c_int32),
                ("min_th_fast", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32), ("flags", C.c_uint32)]


class Camera(C.Structure):
    """orbfe_camera (include/orbfe.h): Camera::create's yaml fields (Camera.cpp:27-52)."""
    _fields_ = [("model", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 12),
                ("n_dist", C.c_int32), ("uncertainty_map", C.c_void_p), ("uncertainty_w", C.c_int32), ("uncertainty_h", C.c_int32)]


CAMERA_PINHOLE, CAMERA_FISHEYE = 0, 1


class OrbfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbfe error %d: %s" % (code, msg))
        self.code = code


_lib = None
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/orbfe.h declares
SIGNATURES = {
    "orbfe_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "orbfe_destroy": (None, [_vp]),
    "orbfe_last_error": (C.c_char_p, [_vp]),
    "orbfe_version": (C.c_char_p, []),
    "orbfe_scale_factor": (_f, [_vp, _i]),
    "orbfe_features_per_level": (_i, [_vp, _i]),
    "orbfe_max_keypoints": (_i, [_vp]),
    "orbfe_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "orbfe_host_free": (None, [_vp]),
    "orbfe_extract": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _i, C.POINTER(_i)]),
    "orbfe_extract_batch": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp]),
    "orbfe_extract_batch_submit": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp, C.POINTER(C.c_longlong)]),
    "orbfe_extract_batch_wait": (_i, [_vp, C.c_longlong]),
    "orbfe_extract_batch_device": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp, _vp, _i]),
    "orbfe_level_size": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "orbfe_get_level_image": (_i, [_vp, _i, _i, _vp]),
    "orbfe_get_level_blurred": (_i, [_vp, _i, _i, _vp]),
    "orbfe_get_level_candidates": (_i, [_vp, _i, _i, _vp, _i, C.POINTER(_i)]),
    "orbfe_get_level_keypoints": (_i, [_vp, _i, _i, _vp, _i, C.POINTER(_i)]),
    "orbfe_launch_count": (C.c_longlong, [_vp]),
    "orbfe_profile": (_i, [_vp, _i]),
    "orbfe_profile_read": (_i, [_vp, _vp, C.POINTER(_i), _i]),
    "orbfe_grid_size": (_i, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "orbfe_frame_postprocess": (_i, [_vp, C.POINTER(Camera), _vp, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_i)]),
    "orbfe_frame_postprocess_device": (_i, [_vp, C.POINTER(Camera), _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "orbfe_popc_peak": (_i, [_vp, C.POINTER(C.c_double)]),
    "orbfe_imma_peak": (_i, [_vp, C.POINTER(C.c_double)]),
    "orbfe_descriptor_distance": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "orbfe_hamming_allpairs": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "orbfe_hamming_allpairs_excl": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "orbfe_hamming_allpairs_excl_device": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "orbfe_hamming_allpairs_slab_device": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "orbfe_hamming_window": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "orbfe_hamming_allpairs_device": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i]),
    "orbfe_search_for_initialization": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _f, _i, C.POINTER(_i)]),
    "orbfe_frame_upload": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(_vp)]),
    "orbfe_frame_wrap_device": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, C.POINTER(_vp)]),
    "orbfe_frame_destroy": (None, [_vp]),
    "orbfe_frame_size": (_i, [_vp]),
    "orbfe_search_for_initialization_f": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _i, C.POINTER(_i)]),
    "orbfe_search_by_projection_f": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(_i)]),
    "orbfe_search_local_points_f": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _f, C.POINTER(_i)]),
    "orbfe_search_by_projection": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, C.POINTER(_i)]),
    "orbfe_search_local_points": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _f, C.POINTER(_i)]),
    "orbfe_check_homography": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "orbfe_check_fundamental": (_i, [_vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "orbfe_vocab_create": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "orbfe_vocab_destroy": (None, [_vp]),
    "orbfe_vocab_words": (_i, [_vp]),
    "orbfe_vocab_transform": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "orbfe_search_fuse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, C.POINTER(_i)]),
    "orbfe_search_fuse_sigma": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, C.POINTER(_i)]),
    "orbfe_compute_descriptors": (_i, [_vp, _vp, _vp, _i, _vp]),
    "orbfe_search_by_bow": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _f, _i, C.POINTER(_i)]),
    "orbfe_search_for_triangulation": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i,
                                            C.POINTER(_i)]),
}


def lib():
    """Load liborbfe.so (built by monoorbslam3_b200.build / __graft_entry__.build()).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s not built: run `python -m monoorbslam3_b200.build` (needs nvcc); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def ptr(a):
    """void* of a numpy array, torch tensor (data_ptr) or raw int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def check(handle, rc):
    if rc != ORBFE_OK:
        msg = lib().orbfe_last_error(handle)
        raise OrbfeError(rc, msg.decode() if msg else "?")


def create(n_features, scale_factor, n_levels, ini_th, min_th, device=0, max_batch=1, flags=0):
    cfg = Config(n_features, scale_factor, n_levels, ini_th, min_th, device, max_batch, flags)
    h = C.c_void_p()
    rc = lib().orbfe_create(C.byref(cfg), C.byref(h))
    if rc != ORBFE_OK:
        msg = lib().orbfe_last_error(None)
        raise OrbfeError(rc, msg.decode() if msg else "?")
    return h
