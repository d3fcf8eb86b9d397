"""Host mirror of the RANSAC scoring of TwoViewReconstruction (Frontend/TwoViewReconstruction.cpp:226-345) over the C-ABI;
the arithmetic runs in liborbfe.so on the GPU (csrc/orbfe_geom.cu)."""
import numpy as np

from . import _capi


def _prep(M, pts1, pts2):
    M = np.ascontiguousarray(M, np.float32).reshape(-1, 9)
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2); p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    return M, p1, p2


def check_homography(extractor, H21, H12, pts1, pts2, sigma=1.0):
    """CheckHomography for every hypothesis: -> (scores[n_hyp] float32, inliers[n_hyp, n_matches] bool)."""
    H21, p1, p2 = _prep(H21, pts1, pts2)
    H12 = np.ascontiguousarray(H12, np.float32).reshape(-1, 9)
    scores = np.zeros(max(len(H21), 1), np.float32); inl = np.zeros((max(len(H21), 1), max(len(p1), 1)), np.uint8)
    _capi.check(extractor._h, _capi.lib().orbfe_check_homography(extractor._h, _capi.ptr(H21), _capi.ptr(H12), len(H21), _capi.ptr(p1), _capi.ptr(p2), len(p1),
                                                                 float(sigma), _capi.ptr(scores), _capi.ptr(inl)))
    return scores[:len(H21)], inl[:len(H21), :len(p1)].astype(bool)


def check_fundamental(extractor, F21, pts1, pts2, sigma=1.0):
    """CheckFundamental for every hypothesis: -> (scores[n_hyp] float32, inliers[n_hyp, n_matches] bool)."""
    F21, p1, p2 = _prep(F21, pts1, pts2)
    scores = np.zeros(max(len(F21), 1), np.float32); inl = np.zeros((max(len(F21), 1), max(len(p1), 1)), np.uint8)
    _capi.check(extractor._h, _capi.lib().orbfe_check_fundamental(extractor._h, _capi.ptr(F21), len(F21), _capi.ptr(p1), _capi.ptr(p2), len(p1), float(sigma),
                                                                  _capi.ptr(scores), _capi.ptr(inl)))
    return scores[:len(F21)], inl[:len(F21), :len(p1)].astype(bool)
