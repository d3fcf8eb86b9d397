"""GPU: Frame post-processing kernels (csrc/orbfe_frame.cu) through the C-ABI against the oracle — undistorted coordinates
bit-exact (also against the committed cv2 vectors directly), sizes, grid CSR identical; host and device entry points."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame_post.npz"))


@pytest.fixture(scope="module")
def env():
    from monoorbslam3_b200 import ORBExtractor, Camera, frame_postprocess, frame_postprocess_device, synth, KP_DTYPE
    from oracle import frame_post as fp
    ex = ORBExtractor(1000, 1.2, 8, 20, 7)
    yield ex, Camera, frame_postprocess, frame_postprocess_device, synth, KP_DTYPE, fp
    ex.close()


def kps_from_points(pts, KP):
    k = np.zeros(len(pts), KP)
    k["x"], k["y"] = pts[:, 0], pts[:, 1]
    k["size"] = 1.0; k["octave"] = 0; k["class_id"] = -1
    return k


@pytest.mark.parametrize("name", ["euroc", "kitti", "phone", "ntu", "strong"])
def test_undistort_equals_cv2_golden(env, name):
    ex, Camera, post, _, _, KP, fp = env
    cam = GOLD["cam_" + name]
    w, h = int(cam[0]), int(cam[1])
    k = kps_from_points(GOLD["pts_" + name], KP)
    out = post(ex, k, Camera(cam[2], cam[3], cam[4], cam[5], cam[6:], "radtan"), w, h)
    ref = GOLD["und_" + name]
    got = np.stack([out.key_points["x"], out.key_points["y"]], 1)
    same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
    assert same.all()                                                                    # bit-exact vs cv2.undistortPoints
    o_raw, o_un, o_off, o_idx = fp.frame_postprocess(k, fp.PINHOLE, cam[2], cam[3], cam[4], cam[5], cam[6:], w, h)
    assert np.array_equal(out.grid_off, o_off) and np.array_equal(out.grid_idx, o_idx)
    assert out.raw_key_points.tobytes() == o_raw.tobytes()


def test_extractor_output_all_models(env):
    ex, Camera, post, _, synth, KP, fp = env
    w, h = 752, 480
    kps, _ = ex(synth.frame(h, w, 1000, "dense"))
    cam = GOLD["cam_euroc"]
    for model, dist, umap in [("radtan", cam[6:], None), ("radtan", [0.0, 0.1, 0.0, 0.0], None), ("equidistant", cam[6:], None),
                              ("equidistant", cam[6:], np.random.default_rng(3).uniform(0.5, 2.0, (h, w)).astype(np.float32))]:
        out = post(ex, kps, Camera(cam[2], cam[3], cam[4], cam[5], dist, model, umap), w, h)
        o_raw, o_un, o_off, o_idx = fp.frame_postprocess(kps, fp.PINHOLE if model == "radtan" else fp.FISHEYE, cam[2], cam[3], cam[4], cam[5], dist, w, h, umap)
        assert out.raw_key_points.tobytes() == o_raw.tobytes(), model
        assert out.key_points.tobytes() == o_un.tobytes(), model
        assert np.array_equal(out.grid_off, o_off) and np.array_equal(out.grid_idx, o_idx), model
    empty = post(ex, np.zeros(0, KP), Camera(cam[2], cam[3], cam[4], cam[5], cam[6:]), w, h)
    assert len(empty.key_points) == 0 and not empty.grid_off.any()


def test_device_batch_chained_after_extractor(env):
    torch = pytest.importorskip("torch")
    ex0, Camera, post, post_dev, synth, KP, fp = env
    from monoorbslam3_b200 import ORBExtractor, grid_size
    w, h, B, cap = 752, 480, 5, 1100
    frames = synth.frames(B, h, w, 40, "dense")
    frames[2] = synth.frame(h, w, 9, "natural")
    ex = ORBExtractor(1000, 1.2, 8, 20, 7, max_batch=B)
    d_fr = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); d_un = torch.zeros_like(d_kps)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda"); d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    cols, rows = grid_size(w, h)
    d_off = torch.zeros((B, cols * rows + 1), dtype=torch.int32, device="cuda"); d_idx = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    d_nin = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(d_fr, B, h, w, d_kps, d_desc, cap, d_n, sync=False)
    cam = GOLD["cam_euroc"]
    camera = Camera(cam[2], cam[3], cam[4], cam[5], cam[6:])
    post_dev(ex, camera, d_kps, d_un, d_n, B, cap, w, h, d_off, d_idx, d_nin, sync=True)
    n = d_n.cpu().numpy()
    for b in range(B):
        raw = d_kps[b, :n[b]].cpu().numpy().view(KP).reshape(-1)
        o_raw, o_un, o_off, o_idx = fp.frame_postprocess(raw, fp.PINHOLE, cam[2], cam[3], cam[4], cam[5], cam[6:], w, h)
        assert d_un[b, :n[b]].cpu().numpy().view(KP).reshape(-1).tobytes() == o_un.tobytes()
        assert np.array_equal(d_off[b].cpu().numpy(), o_off)
        assert int(d_nin[b]) == len(o_idx) and np.array_equal(d_idx[b, :len(o_idx)].cpu().numpy(), o_idx)
    ex.close()
