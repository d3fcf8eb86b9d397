"""Pins the grid and window-query restatements (oracle/orb_oracle.c: orc_grid_build / orc_features_in_area; oracle/frame_post.py:
grid_dims / build_grid) to the reference's own BasicObject/Frame.cpp compiled verbatim (oracle/_ref/libref_frame.so, see
oracle/frame_harness.cpp): Frame::Frame's 40-px grid (Frame.cpp:32-51) and Frame::getFeaturesInArea (:97-127) on random key points —
fractional and out-of-image coordinates, image sizes that are and are not multiples of 40, windows that leave the image, level
filters on and off.  KeyFrame::getFeaturesInArea (KeyFrame.cpp:181-211, `<` for `<=`) is pinned the same way against the reference's
own KeyFrame.cpp (oracle/_ref/libref_keyframe.so)."""
import numpy as np
import pytest

from oracle import frame_post
from oracle import orb_oracle as orc
from oracle import ref_frame as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_frame.so not built (needs the reference sources)")


def _kps(rng, n, w, h):
    k = np.zeros(n, orc.KP_DTYPE)
    k["x"] = rng.uniform(-30, w + 30, n).astype(np.float32); k["y"] = rng.uniform(-30, h + 30, n).astype(np.float32)
    snap = rng.random(n) < 0.3                                        # integer coordinates, cell borders included
    k["x"][snap] = np.round(k["x"][snap] / 40) * 40; k["y"][snap] = np.round(k["y"][snap] / 40) * 40
    k["octave"] = rng.integers(0, 8, n)
    k["size"] = 31; k["angle"] = rng.uniform(0, 360, n)
    return k


@pytest.mark.parametrize("w,h,n,seed", [(752, 480, 1500, 1), (1241, 376, 2000, 2), (1920, 1080, 4000, 3), (640, 480, 300, 4), (41, 39, 50, 5), (80, 40, 0, 6)])
def test_grid_and_window_queries_match_the_reference_frame(w, h, n, seed):
    orc.build()
    rng = np.random.default_rng(seed)
    kps = _kps(rng, n, w, h)
    # ---- the constructor's grid
    roff, ridx, (cols, rows) = ref.grid(kps, w, h)
    assert (cols, rows) == frame_post.grid_dims(w, h)
    off, idx = frame_post.build_grid(kps, w, h)[:2]
    assert np.array_equal(off, roff) and np.array_equal(idx[:off[-1]], ridx)
    # ---- window queries
    nq = 600
    qx = rng.uniform(-120, w + 120, nq).astype(np.float32); qy = rng.uniform(-120, h + 120, nq).astype(np.float32)
    qr = rng.choice(np.array([0.0, 1.0, 7.5, 15.0, 40.0, 100.0, 250.0, 5000.0], np.float32), nq)
    centre = rng.integers(0, max(n, 1), nq)                           # queries centred on key points, radius equal to an exact distance: the <= edge
    if n:
        on_kp = rng.random(nq) < 0.4
        qx[on_kp] = kps["x"][centre[on_kp]]; qy[on_kp] = kps["y"][centre[on_kp]]
        other = rng.integers(0, n, nq)
        edge = on_kp & (rng.random(nq) < 0.5)
        qr[edge] = np.maximum(np.abs(kps["x"][other[edge]] - qx[edge]), np.abs(kps["y"][other[edge]] - qy[edge]))
    mode = rng.integers(0, 4, nq)
    qmin = np.where(mode == 0, -1, rng.integers(0, 8, nq)).astype(np.int32)
    qmax = np.where(mode == 0, -1, np.where(mode == 1, -1, qmin + rng.integers(0, 3, nq))).astype(np.int32)
    qmin[mode == 3] = 0                                               # minLevel 0 with maxLevel >= 0: filter on through maxLevel alone
    roff, ridx, _ = ref.features_in_area(kps, w, h, qx, qy, qr, qmin, qmax)
    koff, kidx = ref.keyframe_features_in_area(kps, w, h, qx, qy, qr, qmin, qmax) if ref.keyframe_available() else (None, None)
    n_nonempty = n_edge = 0
    for i in range(nq):
        got = orc.features_in_area(kps, w, h, float(qx[i]), float(qy[i]), float(qr[i]), int(qmin[i]), int(qmax[i]), strict=False)
        exp = ridx[roff[i]:roff[i + 1]]
        assert np.array_equal(got, exp), (i, qx[i], qy[i], qr[i], qmin[i], qmax[i])
        n_nonempty += len(exp) > 0
        if koff is not None:                                          # the strict variant against the reference's KeyFrame
            got_s = orc.features_in_area(kps, w, h, float(qx[i]), float(qy[i]), float(qr[i]), int(qmin[i]), int(qmax[i]), strict=True)
            exp_s = kidx[koff[i]:koff[i + 1]]
            assert np.array_equal(got_s, exp_s), ("strict", i, qx[i], qy[i], qr[i], qmin[i], qmax[i])
            n_edge += len(exp_s) < len(exp)
    assert n == 0 or n_nonempty > nq // 4
    assert koff is None or n < 300 or n_edge > 10                     # windows whose border passes through a key point: `<` and `<=` differ
