"""Frame post-processing oracle (oracle/frame_post.py) pinned to cv2: the committed golden vectors of cv2.undistortPoints
(tests/golden/frame_post.npz, tools/gen_golden_frame.py) and, when cv2 is importable, a live comparison; plus the grid against a
line-by-line model of Frame.cpp:31-51 / getFeaturesInArea's enumeration order."""
import os
import numpy as np
import pytest

from oracle import frame_post as fp
from oracle import orb_oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame_post.npz"))
CAMS = ["euroc", "kitti", "phone", "ntu", "strong"]


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    same = a.view(np.uint32) == b.view(np.uint32)
    return bool((same | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("name", CAMS)
def test_undistort_matches_cv2_golden(name):
    cam = GOLD["cam_" + name]
    out = fp.undistort_points(GOLD["pts_" + name], cam[2], cam[3], cam[4], cam[5], cam[6:])
    assert bits_equal(out, GOLD["und_" + name])                      # bit-exact, 4006 points per camera


def test_undistort_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for name in CAMS:
        cam = GOLD["cam_" + name]
        w, h = int(cam[0]), int(cam[1])
        pts = np.stack([rng.uniform(0, w, 5000), rng.uniform(0, h, 5000)], 1).astype(np.float32)
        K = np.array([[cam[2], 0, cam[4]], [0, cam[3], cam[5]], [0, 0, 1]], np.float32)
        D = np.array(cam[6:], np.float32).reshape(-1, 1)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
        assert bits_equal(fp.undistort_points(pts, cam[2], cam[3], cam[4], cam[5], cam[6:]), ref)


def make_kps(n, w, h, seed):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, orc.KP_DTYPE)
    k["x"] = rng.uniform(-30, w + 30, n).astype(np.float32); k["y"] = rng.uniform(-30, h + 30, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n); k["size"] = (np.float32(1.2) ** k["octave"]).astype(np.float32)
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32); k["response"] = rng.integers(7, 200, n); k["class_id"] = -1
    return k


def test_grid_is_frame_cpp_order():
    w, h = 752, 480
    k = make_kps(3000, w, h, 1)
    off, idx = fp.build_grid(k, w, h)
    cols, rows = fp.grid_dims(w, h)
    assert (cols, rows) == (19, 12) and len(off) == cols * rows + 1
    grid = [[[] for _ in range(rows)] for _ in range(cols)]                       # Frame.cpp:43-51
    for i in range(len(k)):
        x, y = int(np.floor(k["x"][i])), int(np.floor(k["y"][i]))
        if x < 0 or x >= w or y < 0 or y >= h:
            continue
        grid[x // 40][y // 40].append(i)
    flat = [i for cx in range(cols) for cy in range(rows) for i in grid[cx][cy]]
    assert idx.tolist() == flat
    for cx in range(cols):
        for cy in range(rows):
            c = cx * rows + cy
            assert idx[off[c]:off[c + 1]].tolist() == grid[cx][cy]


def test_postprocess_models():
    w, h = 752, 480
    k = make_kps(500, w, h, 2)
    cam = GOLD["cam_euroc"]
    raw, un, off, idx = fp.frame_postprocess(k, fp.PINHOLE, cam[2], cam[3], cam[4], cam[5], cam[6:], w, h)
    assert raw.tobytes() == k.tobytes()                                             # Pinhole::uncertainty == 1.f
    assert not np.array_equal(un["x"], k["x"]) and np.array_equal(un["size"], k["size"])
    raw0, un0, _, _ = fp.frame_postprocess(k, fp.PINHOLE, cam[2], cam[3], cam[4], cam[5], [0, 0.1, 0, 0], w, h)
    assert un0.tobytes() == k.tobytes()                                             # dist[0] == 0: undistortion skipped (Pinhole.cpp:62)
    m = np.random.default_rng(3).uniform(0.5, 2.0, (h, w)).astype(np.float32)
    kin = k[(k["x"] >= 0) & (k["x"] < w) & (k["y"] >= 0) & (k["y"] < h)]
    raw1, un1, _, _ = fp.frame_postprocess(kin, fp.FISHEYE, cam[2], cam[3], cam[4], cam[5], cam[6:], w, h, m)
    assert np.array_equal(raw1["size"], kin["size"] * m[kin["y"].astype(int), kin["x"].astype(int)])
    assert un1.tobytes() == raw1.tobytes()                                          # Fisheye::undistortKeyPoints copies
