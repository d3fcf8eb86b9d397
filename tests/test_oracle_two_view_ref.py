"""Pins the RANSAC-scoring restatement (oracle/two_view.py) to the reference's own Frontend/TwoViewReconstruction.cpp compiled verbatim
(oracle/_ref/libref_twoview.so, see oracle/twoview_harness.cpp): CheckHomography / CheckFundamental scores (float32, bit for bit — the
reference sums in match order without FMA) and inlier flags, on noisy matches with outliers, several sigmas, degenerate hypotheses.
H12 is the inverse the compiled function worked with (H21.inverse() belongs to the hypothesis generation, outside the path)."""
import numpy as np
import pytest

from oracle import ref_twoview as ref
from oracle import two_view as tv

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_twoview.so not built (needs the reference sources)")


@pytest.mark.parametrize("seed,n", [(0, 80), (1, 300), (2, 1000), (3, 8), (4, 1)])
def test_scoring_matches_the_reference_two_view(seed, n):
    H21, _, Fm, p1, p2 = tv.synthetic_case(seed, n)
    rng = np.random.default_rng(seed)
    for sigma in (1.0, 2.0, 0.5):
        for k in range(6):                                   # perturbed hypotheses like successive RANSAC iterations
            Hk = (H21 + rng.normal(0, 2e-3 * k, (3, 3))).astype(np.float32)
            rs, rin, H12 = ref.check_homography(Hk, p1, p2, sigma)
            s, inl = tv.check_homography(Hk, H12, p1, p2, sigma)
            assert np.float32(s).tobytes() == rs.tobytes() and np.array_equal(inl, rin), (sigma, k)
            Fk = (Fm + rng.normal(0, 1e-5 * k, (3, 3))).astype(np.float32)
            rs, rin = ref.check_fundamental(Fk, p1, p2, sigma)
            s, inl = tv.check_fundamental(Fk, p1, p2, sigma)
            assert np.float32(s).tobytes() == rs.tobytes() and np.array_equal(inl, rin), (sigma, k)
    if n >= 80:
        _, rin, _ = ref.check_homography(H21, p1, p2, 1.0)
        assert rin.any() and not rin.all()
