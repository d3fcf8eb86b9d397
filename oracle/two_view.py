"""CPU oracle of the RANSAC hypothesis scoring (TEST INFRASTRUCTURE ONLY — see oracle/orb_oracle.py's header):
TwoViewReconstruction::CheckHomography / CheckFundamental (Frontend/TwoViewReconstruction.cpp:226-288, 290-345) restated with
numpy float32 arrays — every operation is a separate IEEE single operation in the reference's order (the reference is compiled
without FMA), and the score is the sequential float32 sum over the matches (np.cumsum accumulates in order).
PINNED to the reference's own TwoViewReconstruction.cpp compiled verbatim against a name-level Eigen stand-in (oracle/twoview_harness.cpp,
oracle/_ref/libref_twoview.so): tests/test_oracle_two_view_ref.py compares scores (bit for bit) and inlier flags."""
import numpy as np

F = np.float32


def _seq_sum(t1, t2):
    terms = np.empty(2 * len(t1), F); terms[0::2] = t1; terms[1::2] = t2
    return F(0) if len(terms) == 0 else np.cumsum(terms, dtype=F)[-1]


def check_homography(H21, H12, pts1, pts2, sigma):
    h = np.asarray(H21, F).reshape(3, 3); hi = np.asarray(H12, F).reshape(3, 3)
    u1, v1 = np.asarray(pts1, F)[:, 0], np.asarray(pts1, F)[:, 1]; u2, v2 = np.asarray(pts2, F)[:, 0], np.asarray(pts2, F)[:, 1]
    th = F(5.991); inv_sigma2 = F(1) / (F(sigma) * F(sigma))
    with np.errstate(all="ignore"):
        w2 = F(1) / (hi[2, 0] * u2 + hi[2, 1] * v2 + hi[2, 2])
        u2in1 = (hi[0, 0] * u2 + hi[0, 1] * v2 + hi[0, 2]) * w2; v2in1 = (hi[1, 0] * u2 + hi[1, 1] * v2 + hi[1, 2]) * w2
        chi1 = ((u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1)) * inv_sigma2
        w1 = F(1) / (h[2, 0] * u1 + h[2, 1] * v1 + h[2, 2])
        u1in2 = (h[0, 0] * u1 + h[0, 1] * v1 + h[0, 2]) * w1; v1in2 = (h[1, 0] * u1 + h[1, 1] * v1 + h[1, 2]) * w1
        chi2 = ((u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2)) * inv_sigma2
        in1, in2 = ~(chi1 > th), ~(chi2 > th)
        return _seq_sum(np.where(in1, th - chi1, F(0)), np.where(in2, th - chi2, F(0))), in1 & in2


def check_fundamental(F21, pts1, pts2, sigma):
    f = np.asarray(F21, F).reshape(3, 3)
    u1, v1 = np.asarray(pts1, F)[:, 0], np.asarray(pts1, F)[:, 1]; u2, v2 = np.asarray(pts2, F)[:, 0], np.asarray(pts2, F)[:, 1]
    th, th_score = F(3.841), F(5.991); inv_sigma2 = F(1) / (F(sigma) * F(sigma))
    with np.errstate(all="ignore"):
        a2 = f[0, 0] * u1 + f[0, 1] * v1 + f[0, 2]; b2 = f[1, 0] * u1 + f[1, 1] * v1 + f[1, 2]; c2 = f[2, 0] * u1 + f[2, 1] * v1 + f[2, 2]
        num2 = a2 * u2 + b2 * v2 + c2
        chi1 = num2 * num2 / (a2 * a2 + b2 * b2) * inv_sigma2
        a1 = f[0, 0] * u2 + f[1, 0] * v2 + f[2, 0]; b1 = f[0, 1] * u2 + f[1, 1] * v2 + f[2, 1]; c1 = f[0, 2] * u2 + f[1, 2] * v2 + f[2, 2]
        num1 = a1 * u1 + b1 * v1 + c1
        chi2 = num1 * num1 / (a1 * a1 + b1 * b1) * inv_sigma2
        in1, in2 = ~(chi1 > th), ~(chi2 > th)
        return _seq_sum(np.where(in1, th_score - chi1, F(0)), np.where(in2, th_score - chi2, F(0))), in1 & in2


def synthetic_case(seed, n=80):
    """A homography / fundamental matrix close to the (7, 3) px shift of the synthetic frame pairs, matches with noise and outliers."""
    rng = np.random.default_rng(seed)
    p1 = rng.uniform(20, 700, (n, 2)).astype(F)
    H = np.array([[1.01, 0.02, 7.0], [-0.015, 0.99, 3.0], [1e-5, -2e-5, 1.0]]) + rng.normal(0, 1e-3, (3, 3))
    q = np.c_[p1, np.ones(n)] @ H.T; p2 = (q[:, :2] / q[:, 2:] + rng.normal(0, 1.2, (n, 2))).astype(F)
    p2[::9] += F(25)                                                                  # outliers
    t = np.array([0.3, 0.05, 0.02]); tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Fm = tx @ (np.eye(3) + rng.normal(0, 0.01, (3, 3))) * 1e-3
    return H.astype(F), np.linalg.inv(H).astype(F), Fm.astype(F), p1, p2
