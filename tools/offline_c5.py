#!/usr/bin/env python3
"""BASELINE.json config 5, the offline batch workload: F synthetic 1920x1080 frames (phone.yaml shape, 4000 features) extracted
and all-pairs matched inside non-overlapping windows of 20 key frames, sharded across the GPUs of one node.
Run alone (1 GPU) or under torchrun (--nproc-per-node N).  Windows are the sharding unit — a rank owns a contiguous block of
windows, so extraction and matching need no data-path collective; the per-descriptor match results (best index, best and
second-best distance) are gathered with one NCCL all_gather per array at the end (fixed-capacity slabs).  STRONG scaling: the
total work is fixed.  Prints one JSON line on rank 0; timing = CUDA events on the work stream, max over ranks."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoorbslam3_b200 import ORBExtractor, ORBMatcher, synth, sharding

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, W, NF, WIN, PASS = 1080, 1920, 4000, 20, 128
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n_win = (F + WIN - 1) // WIN
w_lo, w_hi = sharding.shard_range(n_win, rank, world)
f_lo, f_hi = w_lo * WIN, min(w_hi * WIN, F)
nb = f_hi - f_lo
base = synth.frames(8, H, W, 1000 + 100 * rank, "dense")                     # 8 distinct scenes per rank, repeated
frames = torch.from_numpy(base).to(dev)[torch.arange(nb, device=dev) % 8].contiguous() if nb else torch.zeros((0, H, W), dtype=torch.uint8, device=dev)
cap = NF + 128
ex = ORBExtractor(NF, 1.2, 8, 20, 7, device=local, max_batch=PASS)
mt = ORBMatcher(handle=ex._h)
kps = torch.zeros((max(nb, 1), cap, 7), dtype=torch.float32, device=dev); desc = torch.zeros((max(nb, 1), cap, 32), dtype=torch.uint8, device=dev)
n = torch.zeros(max(nb, 1), dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
wcap = WIN * cap
best_i = torch.full((max(w_hi - w_lo, 1), wcap), -1, dtype=torch.int32, device=dev); best_d = torch.zeros_like(best_i); second_d = torch.zeros_like(best_i)


def run():
    matches = 0
    if nb:
        ex.extract_batch_device(frames, nb, H, W, kps, desc, cap, n, stream=s.cuda_stream, sync=False)
    counts = n.cpu().numpy() if nb else np.zeros(0, np.int32)               # per-frame key-point counts size the window tables
    for wi in range(w_hi - w_lo):
        lo = wi * WIN; hi = min(lo + WIN, nb)
        table = torch.cat([desc[b, :int(counts[b])] for b in range(lo, hi)], 0).contiguous()
        m = table.shape[0]
        mt.hamming_allpairs_device(table, m, table, m, best_i[wi], best_d[wi], second_d[wi], stream=s.cuda_stream, sync=False)
        matches += m * m
    return matches


for _ in range(1):
    run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
matches = run()
e1.record()
if world > 1:                                                               # results back in window order on every rank
    per = [sharding.shard_range(n_win, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in per)
    for t in (best_i, best_d, second_d):
        pad = torch.zeros((mx, wcap), dtype=t.dtype, device=dev); pad[:t.shape[0]] = t
        buf = torch.empty((world, mx, wcap), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(buf, pad)
e2.record()
torch.cuda.synchronize()
ms_work, ms_total = e0.elapsed_time(e1), e0.elapsed_time(e2)
stat = torch.tensor([ms_work, ms_total, float(matches), float(nb), float(n[:max(nb, 1)].float().sum())], dtype=torch.float64, device=dev)
if world > 1:
    mxs = stat.clone(); dist.all_reduce(mxs, op=dist.ReduceOp.MAX)
    sums = stat.clone(); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
else:
    mxs = sums = stat
if rank == 0:
    tot_ms = float(mxs[1])
    print(json.dumps({"workload": "C5: %d frames 1920x1080 / 4000 features, all-pairs inside %d-key-frame windows" % (F, WIN), "n_gpus": world,
                      "scaling": "strong", "ms_total": tot_ms, "ms_work_max_rank": float(mxs[0]), "frames_per_s": float(sums[3]) / tot_ms * 1e3,
                      "gmatch_per_s": float(sums[2]) / tot_ms / 1e6, "matches": float(sums[2]), "mean_keypoints_per_frame": float(sums[4]) / max(float(sums[3]), 1),
                      "sharding": "contiguous blocks of key-frame windows per rank; NCCL all_gather of the match results at the end"}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
