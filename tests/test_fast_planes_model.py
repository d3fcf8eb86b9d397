"""CPU model of k_fast_planes' stage A (monoorbslam3_b200/csrc/orbfe_kernels.cuh): the eight difference planes F_D(p) = [|I(p + D) - I(p)| > t]
with the displaced-centre storage of the directions with Dx < 0, and the pair terms F_D(p) | F_D(p - D) of phase 2, against the direct
form of the test (for all eight opposite ring pairs (k, k + 8) one of the two ring pixels differs from the centre by more than t) and
against the oracle's FAST corners, which must all pass it.  No GPU: this pins the algebra the kernel relies on."""
import numpy as np
import pytest

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def planes(img, t):
    """Rows r of the result = image rows r .. (all rows that have r + 3 inside), columns = image columns that have c + 3 inside.
    Plane d as the kernel stores it: directions 5, 6, 7 (Dx = -3, -2, -1) hold G_D(x') = F_D(x' + |Dx|), i.e. |Dx| bits to the left."""
    h, w = img.shape
    I = img.astype(np.int32)
    H, W = h - 3, w - 3
    c0 = I[:H, :W]
    def at(dy, dx): return I[dy:dy + H, dx:dx + W]
    P = np.zeros((8, H, W), bool)
    P[0] = np.abs(c0 - at(3, 0)) > t                 # ( 0, 3)
    P[1] = np.abs(c0 - at(3, 1)) > t                 # ( 1, 3)
    P[2] = np.abs(c0 - at(2, 2)) > t                 # ( 2, 2)
    P[3] = np.abs(c0 - at(1, 3)) > t                 # ( 3, 1)
    P[4] = np.abs(c0 - at(0, 3)) > t                 # ( 3, 0)
    P[5] = np.abs(at(0, 3) - at(1, 0)) > t           # (-3, 1): centre displaced by 3
    P[6] = np.abs(at(0, 2) - at(2, 0)) > t           # (-2, 2)
    P[7] = np.abs(at(0, 1) - at(3, 0)) > t           # (-1, 3)
    return P


def shl(rows, s):
    out = np.zeros_like(rows)
    out[:, s:] = rows[:, :rows.shape[1] - s]
    return out


def pass_bits(img, t):
    """Phase 2 for the pixels (y, x) with 3 <= y, x and y + 3 < h, x + 3 < w; returned array is indexed [y - 3][x] (columns < 3 unused)."""
    P = planes(img, t)
    H = P.shape[1]
    r = slice(3, H)                                   # plane row of pixel row y is y itself
    def up(d, dy): return P[d][3 - dy:H - dy]
    ps = P[0][r] | up(0, 3)
    ps &= P[1][r] | shl(up(1, 3), 1)
    ps &= P[2][r] | shl(up(2, 2), 2)
    ps &= P[3][r] | shl(up(3, 1), 3)
    ps &= P[4][r] | shl(P[4][r], 3)
    ps &= shl(P[5][r], 3) | up(5, 1)
    ps &= shl(P[6][r], 2) | up(6, 2)
    ps &= shl(P[7][r], 1) | up(7, 3)
    return ps


def direct(img, t):
    h, w = img.shape
    I = img.astype(np.int32)
    ok = np.ones((h - 6, w - 6), bool)
    c = I[3:h - 3, 3:w - 3]
    for k in range(8):
        (ax, ay), (bx, by) = RING[k], RING[k + 8]
        a = np.abs(I[3 + ay:h - 3 + ay, 3 + ax:w - 3 + ax] - c) > t
        b = np.abs(I[3 + by:h - 3 + by, 3 + bx:w - 3 + bx] - c) > t
        ok &= a | b
    return ok


@pytest.mark.parametrize("t", [7, 20, 60, 140])
def test_planes_equal_the_direct_pair_test(t):
    rng = np.random.default_rng(t)
    for kind in range(3):
        img = rng.integers(0, 256, (45, 97)).astype(np.uint8)
        if kind == 1: img = (img // 64 * 64).astype(np.uint8)                      # plateaus: many equal neighbours
        if kind == 2: img = np.where(rng.random(img.shape) < 0.5, 0, 255).astype(np.uint8)
        got = pass_bits(img, t)[:, 3:img.shape[1] - 3]
        assert np.array_equal(got, direct(img, t)), (t, kind)


def test_every_oracle_corner_passes(oracle):
    """cv::FAST corners without non-max suppression (the oracle's whole-image FAST-9/16) are a subset of the pixels the planes let
    through, and on the dense synthetic profile the planes let through noticeably more than the corners (gradient pixels)."""
    from monoorbslam3_b200 import synth
    img = synth.frame(120, 160, 77, "dense")
    for t in (7, 20):
        corners = oracle.fast9_16(img, t, nms=False)
        assert len(corners) > 100
        ps = pass_bits(img, t)
        for c in corners:
            assert ps[int(c["y"]) - 3, int(c["x"])], (t, int(c["x"]), int(c["y"]))
        assert ps[:, 3:img.shape[1] - 3].sum() >= len(corners)
