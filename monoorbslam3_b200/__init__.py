"""monoorbslam3_b200 — B200-native ORB front-end (extractor + Hamming matcher) for monoORBSLAM3's hot path.

Only the path is here: csrc/ (CUDA kernels + C-ABI, built into lib/liborbfe.so), the ctypes binding, and host-side mirrors
of the reference's ORBExtractor / ORBMatcher interfaces.  No CPU fallback: without the built library or a CUDA device the
entry points raise."""
from ._capi import KP_DTYPE, OrbfeError, lib      # noqa: F401
from .extractor import ORBExtractor               # noqa: F401
from .matcher import ORBMatcher, FrameView, DeviceFrame   # noqa: F401
from .bow import ORBVocabulary                   # noqa: F401
from .frame import Camera, FramePost, frame_postprocess, frame_postprocess_device, grid_size   # noqa: F401

__all__ = ["ORBExtractor", "ORBMatcher", "FrameView", "DeviceFrame", "KP_DTYPE", "OrbfeError", "lib", "Camera", "FramePost", "frame_postprocess",
           "frame_postprocess_device", "grid_size", "ORBVocabulary"]
