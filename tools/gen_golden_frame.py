#!/usr/bin/env python3
"""Golden vectors of cv2.undistortPoints (OpenCV 4.13.0 in this image) for the Frame post-processing oracle and kernel:
writes tests/golden/frame_post.npz.  Cameras are the reference's own settings files (settings/*.yaml); points are key-point
shaped (integer level coordinates times float32 scale factors) plus uniform samples, image corners and out-of-image points."""
import os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAMS = {   # name: (w, h, fx, fy, cx, cy, dist)
    "euroc": (752, 480, 458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
    "kitti": (1392, 512, 9.786977e+02, 9.717435e+02, 6.900000e+02, 2.497222e+02, [-3.792567e-01, 2.121203e-01, 9.182571e-04, 1.911304e-03, -7.605535e-02]),
    "phone": (1920, 1080, 880.3842060257779, 880.0, 939.1481015502462, 540.0, [-0.04727872906456901, 0.04543545853401388, 0.0009606739301976519, -0.0008318890478998227]),
    "ntu": (752, 480, 4.250258563372763e+02, 4.267976260903337e+02, 3.860151866550880e+02, 2.419130336743440e+02, [-0.288105327549552, 0.074578284234601, 7.784489598138802e-04, -2.277853975035461e-04]),
    "strong": (640, 480, 300.0, 300.0, 320.0, 240.0, [-0.9, 0.5, 0.01, -0.01, -0.3, 0.1, 0.05, 0.01]),      # 8 coefficients, icdist changes sign far out
}
out = {}
rng = np.random.default_rng(42)
scales = [np.float32(1.0)]
for _ in range(7):
    scales.append(np.float32(scales[-1] * np.float32(1.2)))
for name, (w, h, fx, fy, cx, cy, dist) in CAMS.items():
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32).reshape(-1, 1)
    lv = rng.integers(0, 8, 3000)
    kx = (rng.integers(19, w - 19, 3000) / np.array(scales)[lv]).astype(np.int32).astype(np.float32) * np.array(scales, np.float32)[lv]
    ky = (rng.integers(19, h - 19, 3000) / np.array(scales)[lv]).astype(np.int32).astype(np.float32) * np.array(scales, np.float32)[lv]
    uni = np.stack([rng.uniform(-50, w + 50, 1000), rng.uniform(-50, h + 50, 1000)], 1).astype(np.float32)
    corners = np.array([[0, 0], [w, 0], [0, h], [w, h], [cx, cy], [w - 1, h - 1]], np.float32)
    pts = np.concatenate([np.stack([kx, ky], 1), uni, corners]).astype(np.float32)
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
    out["pts_" + name] = pts; out["und_" + name] = und
    out["cam_" + name] = np.array([w, h, fx, fy, cx, cy] + list(dist), np.float64)
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frame_post.npz"), **out)
print("wrote tests/golden/frame_post.npz with", {k: v.shape for k, v in out.items() if k.startswith("pts_")})
