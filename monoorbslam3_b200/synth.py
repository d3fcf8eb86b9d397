"""Deterministic synthetic frames: numpy only, so the same frames exist here and on the GPU box.

dense   : smooth noise field on an (h/8, w/8) grid upsampled bilinearly, 150 alpha-0.5 rectangles, N(0,3) pixel noise
natural : (h/32, w/32) grid with values in [28,228), 250*(w*h)/(752*480) rectangles, N(0,1) noise — exercises the
          FAST threshold fallback and sparse quadtree exits."""
import numpy as np


def _upsample_bilinear(g, h, w):
    gh, gw = g.shape
    ys = (np.arange(h) + 0.5) * (gh - 1) / h
    xs = (np.arange(w) + 0.5) * (gw - 1) / w
    y0 = np.clip(np.floor(ys).astype(int), 0, gh - 2); x0 = np.clip(np.floor(xs).astype(int), 0, gw - 2)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx


def frame(h, w, seed, profile="dense"):
    rng = np.random.default_rng(seed)
    if profile == "dense":
        g = rng.uniform(0, 256, (h // 8 + 2, w // 8 + 2)); n_rect = 150; sigma = 3.0
    elif profile == "natural":
        g = rng.uniform(28, 228, (h // 32 + 2, w // 32 + 2)); n_rect = int(250 * (w * h) / (752 * 480)); sigma = 1.0
    else:
        raise ValueError(profile)
    img = _upsample_bilinear(g, h, w)
    for _ in range(n_rect):
        x0 = int(rng.integers(0, w)); y0 = int(rng.integers(0, h))
        rw = int(rng.integers(5, 80)); rh = int(rng.integers(5, 80)); v = float(rng.uniform(0, 256))
        img[y0:y0 + rh, x0:x0 + rw] = 0.5 * img[y0:y0 + rh, x0:x0 + rw] + 0.5 * v
    img += rng.normal(0, sigma, (h, w))
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def frames(n, h, w, seed0=1000, profile="dense"):
    return np.ascontiguousarray(np.stack([frame(h, w, seed0 + k, profile) for k in range(n)]))


def shifted_pair(h, w, seed, dx=7, dy=3, profile="dense"):
    """Frame k and the same scene shifted by (dx, dy) with fresh noise — the SearchForInitialization workload."""
    big = frame(h + 2 * abs(dy) + 8, w + 2 * abs(dx) + 8, seed, profile)
    a = big[4:4 + h, 4:4 + w].astype(np.int16)
    b = big[4 + dy:4 + dy + h, 4 + dx:4 + dx + w].astype(np.int16)
    rng = np.random.default_rng(seed + 7919)
    b = b + np.rint(rng.normal(0, 1.5, b.shape)).astype(np.int16)
    return np.ascontiguousarray(np.clip(a, 0, 255).astype(np.uint8)), np.ascontiguousarray(np.clip(b, 0, 255).astype(np.uint8))
