"""RANSAC scoring oracle (oracle/two_view.py, vectorised float32) against a scalar line-by-line transcription of
TwoViewReconstruction::CheckHomography / CheckFundamental (Frontend/TwoViewReconstruction.cpp:226-345) with np.float32 scalars."""
import numpy as np

from oracle import two_view as tv

F = np.float32


def scalar_homography(H21, H12, p1, p2, sigma):
    h = [[F(x) for x in r] for r in np.asarray(H21, F).reshape(3, 3)]; hi = [[F(x) for x in r] for r in np.asarray(H12, F).reshape(3, 3)]
    score = F(0); th = F(5.991); inv = F(1) / (F(sigma) * F(sigma)); inl = []
    for (u1, v1), (u2, v2) in zip(np.asarray(p1, F), np.asarray(p2, F)):
        ok = True
        w = F(1) / (hi[2][0] * u2 + hi[2][1] * v2 + hi[2][2])
        a = (hi[0][0] * u2 + hi[0][1] * v2 + hi[0][2]) * w; b = (hi[1][0] * u2 + hi[1][1] * v2 + hi[1][2]) * w
        chi = ((u1 - a) * (u1 - a) + (v1 - b) * (v1 - b)) * inv
        if chi > th: ok = False
        else: score = F(score + (th - chi))
        w = F(1) / (h[2][0] * u1 + h[2][1] * v1 + h[2][2])
        a = (h[0][0] * u1 + h[0][1] * v1 + h[0][2]) * w; b = (h[1][0] * u1 + h[1][1] * v1 + h[1][2]) * w
        chi = ((u2 - a) * (u2 - a) + (v2 - b) * (v2 - b)) * inv
        if chi > th: ok = False
        else: score = F(score + (th - chi))
        inl.append(ok)
    return score, np.array(inl)


def scalar_fundamental(F21, p1, p2, sigma):
    f = [[F(x) for x in r] for r in np.asarray(F21, F).reshape(3, 3)]
    score = F(0); th = F(3.841); ths = F(5.991); inv = F(1) / (F(sigma) * F(sigma)); inl = []
    for (u1, v1), (u2, v2) in zip(np.asarray(p1, F), np.asarray(p2, F)):
        ok = True
        a2 = f[0][0] * u1 + f[0][1] * v1 + f[0][2]; b2 = f[1][0] * u1 + f[1][1] * v1 + f[1][2]; c2 = f[2][0] * u1 + f[2][1] * v1 + f[2][2]
        num = a2 * u2 + b2 * v2 + c2
        chi = num * num / (a2 * a2 + b2 * b2) * inv
        if chi > th: ok = False
        else: score = F(score + (ths - chi))
        a1 = f[0][0] * u2 + f[1][0] * v2 + f[2][0]; b1 = f[0][1] * u2 + f[1][1] * v2 + f[2][1]; c1 = f[0][2] * u2 + f[1][2] * v2 + f[2][2]
        num = a1 * u1 + b1 * v1 + c1
        chi = num * num / (a1 * a1 + b1 * b1) * inv
        if chi > th: ok = False
        else: score = F(score + (ths - chi))
        inl.append(ok)
    return score, np.array(inl)


make_case = tv.synthetic_case


def test_vectorised_equals_scalar_transcription():
    for seed in range(4):
        H21, H12, Fm, p1, p2 = make_case(seed)
        for sigma in (1.0, 2.0):
            s, inl = tv.check_homography(H21, H12, p1, p2, sigma); ss, sinl = scalar_homography(H21, H12, p1, p2, sigma)
            assert s.tobytes() == ss.tobytes() and np.array_equal(inl, sinl) and inl.any() and not inl.all()
            s, inl = tv.check_fundamental(Fm, p1, p2, sigma); ss, sinl = scalar_fundamental(Fm, p1, p2, sigma)
            assert s.tobytes() == ss.tobytes() and np.array_equal(inl, sinl)
