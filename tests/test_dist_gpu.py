"""Multi-GPU behind the C-ABI (include/orbfe_dist.h, liborbfe_dist.so: one host thread and one NCCL communicator per GPU, NCCL called
from C++): sharded extraction (host and device-resident with a gather to either end of the group) and sharded all-pairs equal the
single-GPU C-ABI bitwise.  Runs on every GPU count the box offers from 2 up (skipped on a single-GPU box); world size 1 is run
everywhere — it exercises the same code without the collectives."""
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_c_abi_equals_single_gpu(tmp_path, world):
    import torch
    from monoorbslam3_b200 import synth, build
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    exe = build.build_cpp_dist_test()
    B, H, W = 13, 480, 752                                     # 13 frames: uneven shards (and empty ones at world 8 would need B < 8)
    frames = synth.frames(B, H, W, 7100, "dense")
    frames[5] = synth.frame(H, W, 9, "natural")
    p = str(tmp_path / "frames.raw")
    frames.tofile(p)
    r = subprocess.run([exe, str(world), p, str(B), str(W), str(H)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dist_test ok" in r.stdout


def test_python_mirror_of_the_dist_group(oracle):
    """monoorbslam3_b200.dist.DistGroup over every GPU of the box: the same results as the single-GPU mirrors (and the oracle)."""
    import torch
    from monoorbslam3_b200 import ORBExtractor, ORBMatcher, synth
    from monoorbslam3_b200.dist import DistGroup
    g = DistGroup(torch.cuda.device_count(), 1000, 1.2, 8, 20, 7)
    frames = synth.frames(5, 480, 752, 7300, "dense")
    n, kps, desc = g.extract_batch(frames, cap=1100)
    ex = ORBExtractor(1000, 1.2, 8, 20, 7)
    for b in range(5):
        k, d = ex(frames[b])
        assert n[b] == len(k) and kps[b, :n[b]].tobytes() == k.tobytes() and np.array_equal(desc[b, :n[b]], d)
    table = np.concatenate([desc[b, :n[b]] for b in range(5)])
    got = g.hamming_allpairs(table[:3000], table)
    exp = oracle.hamming_allpairs(table[:3000], table)
    assert all(np.array_equal(a, b) for a, b in zip(got, exp))
    g.close()
