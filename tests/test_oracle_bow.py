"""SearchByBow oracle (oracle/orb_oracle.c: orc_search_by_bow) against a line-by-line Python transcription of
ORBMatcher.cpp:118-201 on small random inputs (descriptor ties, occupied slots, skipped nodes, rotation histogram)."""
import numpy as np
import pytest

from oracle import orb_oracle as orc


def py_search_by_bow(d1, a1, valid1, fv1, d2, a2, occ2, fv2, ratio, orient):
    ids1, off1, idx1 = fv1; ids2, off2, idx2 = fv2
    taken = occ2.astype(bool).copy(); asg = np.full(len(d2), -1, np.int32); n = 0
    hist = [[] for _ in range(30)]
    i = j = 0
    while i < len(ids1) and j < len(ids2):
        if ids1[i] == ids2[j]:
            for q in idx1[off1[i]:off1[i + 1]]:
                if not valid1[q]:
                    continue
                best, second, bi = 256, 256, -1
                for c in idx2[off2[j]:off2[j + 1]]:
                    if taken[c]:
                        continue
                    dist = int(np.unpackbits(d1[q] ^ d2[c]).sum())
                    if dist < best:
                        second, best, bi = best, dist, c
                    elif dist < second:
                        second = dist
                if best <= 50 and np.float32(best) < np.float32(ratio) * np.float32(second):
                    taken[bi] = True; asg[bi] = q; n += 1
                    if orient:
                        rot = np.float32(a1[q]) - np.float32(a2[bi])
                        if rot < 0:
                            rot = np.float32(rot + np.float32(360))
                        b = int(np.rint(np.float32(rot * np.float32(1.0 / 30))))
                        hist[0 if b == 30 else b].append(bi)
            i += 1; j += 1
        elif ids1[i] < ids2[j]:
            while i < len(ids1) and ids1[i] < ids2[j]:
                i += 1
        else:
            while j < len(ids2) and ids2[j] < ids1[i]:
                j += 1
    if orient:
        i1, i2, i3 = orc.compute_three_maxima([len(h) for h in hist])
        for b in range(30):
            if b in (i1, i2, i3):
                continue
            for s in hist[b]:
                asg[s] = -1; n -= 1
    return n, asg


def fv_of(desc, bits):
    node = desc[:, 0].astype(np.int32) >> (8 - bits)
    ids = np.unique(node); off = [0]; idx = []
    for v in ids:
        idx.extend(np.nonzero(node == v)[0].tolist()); off.append(len(idx))
    return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


@pytest.mark.parametrize("seed,ratio,orient", [(1, 0.7, True), (2, 0.9, False), (3, 0.6, True), (4, 1.0, True)])
def test_search_by_bow_oracle_matches_transcription(seed, ratio, orient):
    rng = np.random.default_rng(seed)
    n1, n2 = 300, 280
    base = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    d1 = base[rng.integers(0, 200, n1)].copy(); d2 = base[rng.integers(0, 200, n2)].copy()
    flip = rng.integers(0, 256, (n2, 32), dtype=np.uint8) & rng.integers(0, 256, (n2, 32), dtype=np.uint8) & rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    d2 ^= (flip & 0x11).astype(np.uint8); d2[:, 0] = (d2[:, 0] & 0x1f) | (d1[rng.integers(0, n1, n2), 0] & 0xe0)       # a few flipped bits, shared nodes
    a1 = rng.uniform(0, 360, n1).astype(np.float32); a2 = (a1[rng.integers(0, n1, n2)] + rng.normal(0, 3, n2)).astype(np.float32) % np.float32(360)
    valid1 = (rng.random(n1) < 0.8).astype(np.uint8); occ2 = (rng.random(n2) < 0.2).astype(np.uint8)
    fv1, fv2 = fv_of(d1, 3), fv_of(d2, 3)
    n, asg = orc.search_by_bow(d1, a1, valid1, fv1, d2, a2, occ2, fv2, ratio, orient)
    pn, pasg = py_search_by_bow(d1, a1, valid1, fv1, d2, a2, occ2, fv2, ratio, orient)
    assert n == pn and np.array_equal(asg, pasg)
    assert n > 5
