"""The C++ host adapters (monoorbslam3_b200/host: the reference's class names and signatures over the C-ABI) give the same
results as the Python mirror, which the other GPU suites pin to the oracle."""
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cpp_adapters_match_python_mirror(tmp_path):
    from monoorbslam3_b200 import ORBExtractor, ORBMatcher, FrameView, KP_DTYPE, synth, build
    exe = build.build_cpp_adapter_test()
    a, b = synth.shifted_pair(480, 752, 1000)
    pa, pb, out = str(tmp_path / "a.raw"), str(tmp_path / "b.raw"), str(tmp_path / "out.bin")
    a.tofile(pa); b.tofile(pb)
    r = subprocess.run([exe, "752", "480", pa, pb, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(out, "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        arr = np.frombuffer(buf, dtype, n, off).copy()
        off += arr.nbytes
        return arr

    frames = []
    for _ in range(2):
        n = int(take(np.int32, 1)[0])
        frames.append((take(KP_DTYPE, n), take(np.uint8, n * 32).reshape(n, 32)))
    nm = int(take(np.int32, 1)[0])
    m12 = take(np.int32, len(frames[0][0]))
    pre = take(np.float32, 2 * len(frames[0][0])).reshape(-1, 2)
    n3 = int(take(np.int32, 1)[0])
    k3 = take(KP_DTYPE, n3)

    ex2 = ORBExtractor(2000, 1.2, 8, 20, 7)
    for img, (k, d) in zip((a, b), frames):
        pk, pd = ex2(img)
        assert pk.tobytes() == k.tobytes() and np.array_equal(pd, d)
    pk3, _ = ORBExtractor(1000, 1.2, 8, 20, 7)(a)
    assert pk3.tobytes() == k3.tobytes()
    f1 = FrameView(frames[0][0], frames[0][1], 752, 480); f2 = FrameView(frames[1][0], frames[1][1], 752, 480)
    ppre = np.stack([frames[0][0]["x"], frames[0][0]["y"]], 1).astype(np.float32)
    pn, pm12 = ORBMatcher(0.9, True).SearchForInitialization(f1, f2, ppre, 100)
    assert pn == nm and np.array_equal(pm12, m12) and np.array_equal(ppre, pre)
    # Frame post-processing through the C++ helper == the Python mirror (which the frame suite pins to cv2 and the oracle)
    from monoorbslam3_b200 import Camera, frame_postprocess
    un = take(KP_DTYPE, n3)
    gc, gr = (int(v) for v in take(np.int32, 2))
    post = frame_postprocess(ex2, k3, Camera(458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]), 752, 480)
    assert un.tobytes() == post.key_points.tobytes() and (gc, gr) == (post.cols, post.rows)
    for cx in range(gc):
        for cy in range(gr):
            m = int(take(np.int32, 1)[0])
            assert take(np.int32, m).tolist() == post.cell(cx, cy).tolist()
