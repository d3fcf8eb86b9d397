"""GPU: RANSAC hypothesis scoring (csrc/orbfe_geom.cu) through the C-ABI against the float32 restatement of
TwoViewReconstruction::CheckHomography / CheckFundamental — scores and inlier flags bit-identical for 200 hypotheses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_check_homography_and_fundamental_bit_exact():
    from monoorbslam3_b200 import ORBExtractor
    from monoorbslam3_b200.geometry import check_homography, check_fundamental
    from oracle import two_view as tv
    make_case = tv.synthetic_case
    ex = ORBExtractor(1000, 1.2, 8, 20, 7)
    rng = np.random.default_rng(0)
    for n in (0, 1, 333, 5000):                                   # 5000 > one shared-memory round of 4096 matches
        H0, _, F0, p1, p2 = make_case(7, max(n, 1))
        p1, p2 = p1[:n], p2[:n]
        H21 = (H0[None] + rng.normal(0, 2e-3, (200, 3, 3))).astype(np.float32)
        H21[5] = 0                                                 # degenerate hypothesis: divisions by zero, NaN / inf terms
        H12 = np.stack([np.linalg.pinv(h.astype(np.float64)).astype(np.float32) for h in H21])
        F21 = (F0[None] * (1 + rng.normal(0, 0.2, (200, 3, 3)))).astype(np.float32)
        for sigma in (1.0, 1.5):
            s, inl = check_homography(ex, H21, H12, p1, p2, sigma)
            for j in range(0, 200, 7):
                os_, oinl = tv.check_homography(H21[j], H12[j], p1, p2, sigma)
                assert np.float32(s[j]).tobytes() == np.float32(os_).tobytes() or (np.isnan(s[j]) and np.isnan(os_)), (n, j)
                assert np.array_equal(inl[j], oinl)
            s, inl = check_fundamental(ex, F21, p1, p2, sigma)
            for j in range(0, 200, 7):
                os_, oinl = tv.check_fundamental(F21[j], p1, p2, sigma)
                assert np.float32(s[j]).tobytes() == np.float32(os_).tobytes() or (np.isnan(s[j]) and np.isnan(os_)), (n, j)
                assert np.array_equal(inl[j], oinl)
    ex.close()
