"""T4 (CPU side): the sharded workloads return exactly the single-process result.  world_size 2 over gloo; the per-block compute
is the oracle here (the GPU suite checks the CUDA path itself) — what is under test is the partitioning and the gather."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAP = 400


def _oracle_extract_block(frames):
    from oracle import orb_oracle as orc
    ex = orc.Extractor(300, 1.2, 8, 20, 7)
    B = frames.shape[0]
    n = torch.zeros(B, dtype=torch.int32); kps = torch.zeros((B, CAP, 7), dtype=torch.float32); desc = torch.zeros((B, CAP, 32), dtype=torch.uint8)
    for b in range(B):
        k, d = ex(frames[b].numpy())
        n[b] = len(k)
        kps[b, :len(k)] = torch.from_numpy(k.view(np.float32).reshape(-1, 7).copy())
        desc[b, :len(k)] = torch.from_numpy(d)
    return n, kps, desc


def _oracle_match(q, t):
    from oracle import orb_oracle as orc
    bi, bd, sd = orc.hamming_allpairs(q.numpy(), t.numpy())
    return torch.from_numpy(bi), torch.from_numpy(bd), torch.from_numpy(sd)


def _worker(rank, world, port, frames, descs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monoorbslam3_b200 import sharding
    n, kps, desc = sharding.extract_sharded(_oracle_extract_block, frames, CAP)
    lo, hi = sharding.shard_range(descs.shape[0], rank, world)
    bi, bd, sd = sharding.allpairs_sharded(_oracle_match, descs[lo:hi])
    if rank == 0:
        torch.save({"n": n, "kps": kps, "desc": desc, "bi": bi, "bd": bd, "sd": sd}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from monoorbslam3_b200.sharding import shard_range
    for n in (0, 1, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_equals_single_process(tmp_path, oracle):
    from monoorbslam3_b200 import synth
    frames = torch.from_numpy(synth.frames(5, 160, 240, 50, "dense"))          # 5 frames over 2 ranks: uneven shards
    rng = np.random.default_rng(0)
    descs = torch.from_numpy(rng.integers(0, 256, (301, 32), dtype=np.uint8))
    out = str(tmp_path / "res.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, frames, descs, out), nprocs=2, join=True)
    got = torch.load(out)
    n, kps, desc = _oracle_extract_block(frames)
    assert torch.equal(got["n"], n) and torch.equal(got["kps"].view(torch.int32), kps.view(torch.int32)) and torch.equal(got["desc"], desc)
    bi, bd, sd = _oracle_match(descs, descs)
    assert torch.equal(got["bi"], bi) and torch.equal(got["bd"], bd) and torch.equal(got["sd"], sd)
