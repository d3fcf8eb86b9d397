"""The GPU quadtree (k_octree) does not move key-point vectors: a node is a path prefix, the list order is recovered from the
creation order of 4-child slots, and the careful phase commits a sorted prefix.  This is a line-by-line Python model of that
formulation, checked against the oracle's list-based DistributeOctree (ORBExtractor.cpp:640-830) on random inputs."""
import math
import numpy as np
import pytest


def gpu_formulation(xs, ys, sc, W, H, max_x, nF):
    n = len(xs)
    n_ini = int(math.ceil(float(np.float32(W) / np.float32(H))))
    h_x = int(math.ceil(float(np.float32(W) / np.float32(n_ini))))
    nb, ncnt, nchild = {}, {}, {}
    for r in range(n_ini):
        nb[r] = (h_x * r, max_x if r == n_ini - 1 else h_x * (r + 1), 0, H); ncnt[r] = 0; nchild[r] = 0
    cur = xs // h_x
    for i in range(n):
        ncnt[cur[i]] += 1
    S = [r for r in range(n_ini) if ncnt[r] > 1]
    length = sum(1 for r in range(n_ini) if ncnt[r] > 0)
    pool_top, careful, E, ns = n_ini, False, list(S), len(S)
    while True:
        prev = length
        if careful:
            S = [E[j] for j in sorted(range(len(E)), key=lambda j: (ncnt[E[j]], j))]
            ns = len(S)
        if ns == 0:
            break
        for j, nd in enumerate(S):
            x0, x1, y0, y1 = nb[nd]
            base = pool_top + 4 * j
            nchild[nd] = base
            mx, my = x0 + (x1 - x0) // 2, y0 + (y1 - y0) // 2
            nb[base], nb[base + 1], nb[base + 2], nb[base + 3] = (x0, mx, y0, my), (mx, x1, y0, my), (x0, mx, my, y1), (mx, x1, my, y1)
            for q in range(4):
                ncnt[base + q] = 0; nchild[base + q] = 0
        tent = cur.copy()
        for i in range(n):
            base = nchild[cur[i]]
            if base > 0:
                b = nb[base]
                ch = base + (0 if xs[i] < b[1] else 1) + (0 if ys[i] < b[3] else 2)
                tent[i] = ch; ncnt[ch] += 1
        n_commit = ns
        if careful:
            run, found = length, ns
            for j in range(ns):
                b = pool_top + 4 * j
                run += sum(1 for q in range(4) if ncnt[b + q] > 0) - 1
                if run >= nF:
                    found = j; break
            n_commit = found + 1 if found < ns else ns
            for j in range(n_commit, ns):
                nchild[S[j]] = 0
            for i in range(n):
                if nchild[cur[i]] > 0:
                    cur[i] = tent[i]
        else:
            cur = tent
        newE, added = [], 0
        for s in range(4 * n_commit):
            c = ncnt[pool_top + s]
            if c > 1: newE.append(pool_top + s)
            if c > 0: added += 1
        length = length - n_commit + added
        pool_top += 4 * n_commit
        E = newE
        if careful:
            if length >= nF or length == prev: break
        else:
            if length > nF or length == prev: break
            if length + 3 * len(E) > nF: careful = True
            else:
                S = E[::-1]; ns = len(S)
    leaves = []
    for t in range(pool_top):
        slot = pool_top - 1 - t if t < pool_top - n_ini else t - (pool_top - n_ini)
        if ncnt[slot] > 0 and nchild[slot] == 0: leaves.append(slot)
    pos = {s: k for k, s in enumerate(leaves)}
    best = [-1] * len(leaves)
    for i in range(n):
        p = pos[cur[i]]
        if best[p] < 0 or sc[i] > sc[best[p]]: best[p] = i
    return np.array(best, np.int32), pool_top


def node_pool_bound(W, H, nF):
    """Size of the kernel's global node pool (configure() in csrc/orbfe_extract.cu): every pass that does not end the loop splits all
    expandable nodes, so there are at most ceil(log2(max(W, H))) + 1 passes of at most max(quota, roots) splits, 4 slots each."""
    n_ini = math.ceil(W / H)
    depth = 1
    while (1 << depth) < max(W, H):
        depth += 1
    return 4 * max(nF + 4, 4 * n_ini + 4) * (depth + 2) + 5 * n_ini + 64


def tight_pairs(W, H, gx, gy, cluster, gap=2):
    """Sparse pairs of key points `gap` px apart on a gx x gy lattice plus a dense little cluster: every pair keeps splitting (four
    slots per split, three of them empty) until the pair separates, which is what exhausts an empirically sized pool."""
    pts = []
    for j in range(gy):
        for i in range(gx):
            x, y = int((i + 0.5) * W / gx), int((j + 0.5) * H / gy)
            pts += [(x, y), (x + gap, y)]
    pts += [(5 + 4 * i, 5 + 4 * j) for j in range(cluster) for i in range(cluster)]      # 4 px apart: no dot on another's 16-ring
    pts = np.unique(np.array(pts), axis=0)
    return pts[np.lexsort((pts[:, 0], pts[:, 1], pts[:, 0] // 30, pts[:, 1] // 30))]


@pytest.mark.parametrize("gx,gy,cluster", [(26, 16, 8), (26, 16, 4), (29, 15, 8), (20, 20, 8)])
def test_sparse_tight_pairs_stay_within_the_proven_pool(oracle, gx, gy, cluster):
    """Level 0 of a 1920x1080 / 4000-feature frame (W=1882, H=1042, quota 868): these inputs need 7290-7890 node slots, more than
    the 7050 of the formula the kernel used in round 1 (8 * (quota + 4) + 5 * roots + 64); the proven bound holds them, and the
    slot formulation still equals the list formulation."""
    W, H, nF = 1882, 1042, 868
    pts = tight_pairs(W, H, gx, gy, cluster)
    c = np.zeros(len(pts), oracle.CORNER_DTYPE)
    c["x"], c["y"], c["score"] = pts[:, 0], pts[:, 1], 50
    ref = oracle.distribute_octree(c, 19, W + 19, 19, H + 19, nF)
    got, top = gpu_formulation(c["x"].astype(int), c["y"].astype(int), c["score"].astype(int), W, H, W + 19, nF)
    assert np.array_equal(ref, got)
    assert 8 * (nF + 4) + 5 * 2 + 64 < top <= node_pool_bound(W, H, nF)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_slot_formulation_equals_list_formulation(oracle, seed):
    rng = np.random.default_rng(seed)
    for trial in range(120):
        W, H = int(rng.integers(40, 900)), int(rng.integers(40, 500))
        n = int(rng.integers(1, 2500))
        kind = rng.integers(0, 4)
        if kind == 0: pts = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1)
        elif kind == 1:
            cx, cy = rng.integers(0, W), rng.integers(0, H)
            pts = np.stack([np.clip(rng.normal(cx, 6, n), 0, W - 1), np.clip(rng.normal(cy, 6, n), 0, H - 1)], 1).astype(int)
        elif kind == 2: pts = np.stack([rng.integers(0, min(W, 12), n), rng.integers(0, H, n)], 1)
        else:
            pts = np.stack([np.minimum(rng.integers(0, W, n) // 3 * 3, W - 1), np.minimum(rng.integers(0, H, n) // 7 * 7, H - 1)], 1)
        pts = np.unique(pts, axis=0)
        pts = pts[np.lexsort((pts[:, 0], pts[:, 1], pts[:, 0] // 30, pts[:, 1] // 30))]      # (cellRow, cellCol, y, x)
        c = np.zeros(len(pts), oracle.CORNER_DTYPE)
        c["x"], c["y"], c["score"] = pts[:, 0], pts[:, 1], rng.integers(7, 60, len(pts))
        nF = int(rng.choice([1, 5, 26, 75, 156, 323, 646, 1292]))
        ref = oracle.distribute_octree(c, 19, W + 19, 19, H + 19, nF)
        got, top = gpu_formulation(c["x"].astype(int), c["y"].astype(int), c["score"].astype(int), W, H, W + 19, nF)
        assert np.array_equal(ref, got), (trial, W, H, len(pts), nF)
        assert top <= node_pool_bound(W, H, nF)
