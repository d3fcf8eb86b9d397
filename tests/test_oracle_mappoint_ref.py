"""Pins the restatement of MapPoint::computeDescriptor (oracle/orb_oracle.c: orc_compute_descriptors) to the reference's own
BasicObject/MapPoint.cpp compiled verbatim (oracle/_ref/libref_mappoint.so, see oracle/mappoint_harness.cpp).  The reference walks a
point's observations in std::map order over key-frame ADDRESSES; the harness reports that order, the restatement is given the rows in
it (what the adapter of INTEGRATION.md does), and must pick the descriptor the reference picked — group sizes 1..70, exact duplicates
(median ties: the first row in that order wins), bad key frames."""
import numpy as np
import pytest

from oracle import orb_oracle as orc
from oracle import ref_mappoint as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_mappoint.so not built (needs the reference sources)")


@pytest.mark.parametrize("seed,with_bad", [(1, False), (2, False), (3, True)])
def test_compute_descriptor_matches_the_reference_mappoint(seed, with_bad):
    orc.build()
    rng = np.random.default_rng(seed)
    sizes = [1, 2, 3, 7, 32, 33, 70, 5, 2, 2, 40] + rng.integers(1, 25, 200).tolist()
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.repeat(base, sizes, axis=0).copy()
    desc ^= np.packbits(rng.random((len(desc), 32, 8)) < 0.06, axis=2).reshape(len(desc), 32)
    desc[off[3]:off[3] + 3] = desc[off[3]]                           # exact duplicates
    desc[off[4] + 5] = desc[off[4] + 20]
    bad = (rng.random(len(desc)) < 0.15).astype(np.uint8) if with_bad else np.zeros(len(desc), np.uint8)
    bad[off[:-1]] = 0                                                 # keep every group non-empty
    order, n_order, chosen = ref.compute_descriptors(desc, off, bad)
    # the restatement on the rows in the reference's iteration order
    rows = np.concatenate([order[off[g]:off[g] + n_order[g]] for g in range(len(sizes))])
    roff = np.concatenate([[0], np.cumsum(n_order)]).astype(np.int32)
    best = orc.compute_descriptors(desc[rows], roff)
    assert (n_order == np.array(sizes) - np.add.reduceat(bad, off[:-1])).all()
    for g in range(len(sizes)):
        assert best[g] >= 0
        assert np.array_equal(desc[rows[roff[g] + best[g]]], chosen[g]), g
