"""CPU model of the dead-column mask of k_allpairs_tc's epilogue (monoorbslam3_b200/csrc/orbfe_allpairs_tc.cu): for a 128-column train tile
the kernel builds, per query row, a 128-bit mask from at most four intervals with shifts — columns beyond the table, the row's own
key-frame range, the padding rows of the slab block the tile starts in, the padding rows of the next block — instead of testing every
column.  The model checks the shift construction against the column-by-column definition, and the two whole-tile skip rules against it.
No GPU: this pins the interval algebra; the kernel itself is compared with numpy / the oracle in tests/test_match_gpu.py."""
import numpy as np

TILE = 128
BIG = 0x7fffffff


def ge(t, w):
    """bits b of 32-bit word w with 32 w + b >= t (the kernel's shift of 0xffffffff by a clamped amount)"""
    s = min(max(t - 32 * w, 0), 32)
    return (0xffffffff << s) & 0xffffffff


def mask_by_shifts(col0, nt, ex, n0, f_split, n1):
    r_nt, r_e0, r_e1, r_n0, r_sp, r_n1 = nt - col0, ex[0] - col0, ex[1] - col0, n0 - col0, f_split - col0, n1 - col0
    return [ge(r_nt, w) | (ge(r_e0, w) & ~ge(r_e1, w) & 0xffffffff) | (ge(r_n0, w) & ~ge(r_sp, w) & 0xffffffff) | (ge(r_sp, w) & ge(r_n1, w)) for w in range(4)]


def dead(cc, nt, ex, n0, f_split, n1):
    return cc >= nt or (ex[0] <= cc < ex[1]) or (cc >= n0 if cc < f_split else cc >= n1)


def slab_terms(col0, nt, cap, counts):
    """f_split, n0, n1 as the kernel derives them for a tile of a slab window (blocks of `cap` rows, counts[f] key points each)"""
    f0 = col0 // cap
    f_split = (f0 + 1) * cap
    n0 = f0 * cap + counts[f0]
    n1 = f_split + counts[f0 + 1] if (f_split < col0 + TILE and f_split < nt) else BIG
    return f_split, n0, n1


def test_mask_equals_the_per_column_definition():
    rng = np.random.default_rng(3)
    for trial in range(4000):
        slab = trial % 2 == 0
        if slab:
            cap = int(rng.choice([96, 128, 130, 257, 384, 700]))
            n_frames = int(rng.integers(1, 6))
            counts = [int(rng.integers(1, cap + 1)) for _ in range(n_frames)]
            nt = n_frames * cap
            col0 = int(rng.integers(0, (nt + TILE - 1) // TILE)) * TILE
            row = int(rng.integers(0, nt))
            f = row // cap
            ex = (f * cap, (f + 1) * cap)
            f_split, n0, n1 = slab_terms(col0, nt, cap, counts)
        else:
            nt = int(rng.integers(1, 3000))
            col0 = int(rng.integers(0, (nt + TILE - 1) // TILE)) * TILE
            a = int(rng.integers(0, nt + 1)); b = int(rng.integers(0, nt + 1))
            ex = (min(a, b), max(a, b)) if rng.random() < 0.8 else (7, 7)
            f_split, n0, n1 = BIG, BIG, BIG
        m = mask_by_shifts(col0, nt, ex, n0, f_split, n1)
        for c in range(TILE):
            assert bool((m[c >> 5] >> (c & 31)) & 1) == dead(col0 + c, nt, ex, n0, f_split, n1), (trial, c)


def test_whole_tile_skips_only_drop_dead_tiles():
    """A tile inside the own range of every row of a warp, or inside the padding of its block, is dead in every column for those rows."""
    rng = np.random.default_rng(4)
    for trial in range(2000):
        cap = int(rng.choice([128, 256, 384, 640, 4096]))
        n_frames = int(rng.integers(2, 5))
        counts = [int(rng.integers(1, cap + 1)) for _ in range(n_frames)]
        nt = n_frames * cap
        col0 = int(rng.integers(0, nt // TILE)) * TILE
        f_split, n0, n1 = slab_terms(col0, nt, cap, counts)
        rows = int(rng.integers(0, nt - 31)) + np.arange(32)
        exs = [((r // cap) * cap, (r // cap + 1) * cap) for r in rows]
        c_lo, c_hi = max(e[0] for e in exs), min(e[1] for e in exs)
        skip = (col0 >= c_lo and col0 + TILE <= c_hi) or (col0 >= n0 and col0 + TILE <= f_split)
        if skip:
            for ex in exs:
                assert all(dead(col0 + c, nt, ex, n0, f_split, n1) for c in range(TILE)), trial
