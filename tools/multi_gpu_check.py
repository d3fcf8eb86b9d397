#!/usr/bin/env python3
"""T4 on real GPUs (run with torchrun --nproc-per-node N): frames sharded by rank + NCCL gather == single-GPU result, bitwise;
key-frame-window all-pairs sharded by query rows == single-GPU result.  Prints one line per check on rank 0."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, synth, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W, NF, CAP, B = 480, 752, 1000, 1064, 37            # 37 frames: uneven shards
frames = torch.from_numpy(synth.frames(B, H, W, 7000, "dense")).to(dev)
ex = ORBExtractor(NF, 1.2, 8, 20, 7, device=local, max_batch=64)
stream = torch.cuda.current_stream().cuda_stream


def extract(block):
    b = block.shape[0]
    kps = torch.zeros((b, CAP, 7), dtype=torch.float32, device=dev); desc = torch.zeros((b, CAP, 32), dtype=torch.uint8, device=dev)
    n = torch.zeros(b, dtype=torch.int32, device=dev)
    if b:
        torch.cuda.synchronize()
        ex.extract_batch_device(block.contiguous(), b, H, W, kps, desc, CAP, n, sync=True)
    return n, kps, desc


n, kps, desc = sharding.extract_sharded(extract, frames, CAP)
n1, kps1, desc1 = extract(frames)
ok = torch.equal(n, n1)
for b in range(B):
    k = int(n1[b])
    ok = ok and torch.equal(kps[b, :k].view(torch.int32), kps1[b, :k].view(torch.int32)) and torch.equal(desc[b, :k], desc1[b, :k])
# key-frame window: 20 "key frames" x their descriptors
table = torch.cat([desc1[b, :int(n1[b])] for b in range(20)], 0).contiguous()
mt = ORBMatcher(handle=ex._h)


def match(q, t):
    nq = q.shape[0]
    bi = torch.zeros(nq, dtype=torch.int32, device=dev); bd = torch.zeros_like(bi); sd = torch.zeros_like(bi)
    if nq:
        torch.cuda.synchronize()
        mt.hamming_allpairs_device(q.contiguous(), nq, t, t.shape[0], bi, bd, sd, sync=True)
    return bi, bd, sd


lo, hi = sharding.shard_range(table.shape[0], rank, world)
bi, bd, sd = sharding.allpairs_sharded(match, table[lo:hi].contiguous())
bi1, bd1, sd1 = match(table, table)
ok2 = torch.equal(bi, bi1) and torch.equal(bd, bd1) and torch.equal(sd, sd1)
flags = torch.tensor([int(ok), int(ok2)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: sharded extraction == single GPU (bitwise): %s; sharded all-pairs (%d descriptors) == single GPU: %s" %
          (world, bool(flags[0]), table.shape[0], bool(flags[1])))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flags.min()) else 1)
