"""Multi-GPU plumbing for the two workloads that shard: batched extraction (frames are independent units)
and key-frame-window all-pairs matching (query rows are independent).  One process per GPU; torch.distributed carries the
only exchange steps (NCCL on the GPUs, gloo in the CPU tests): an all-gather of fixed-capacity result slabs, and an
all-gather of the descriptor table before matching.  The single-sequence tracking loop does not shard (replicas only)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block of units for `rank`: [lo, hi) with sizes differing by at most one, in rank order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def _all_gather(t, world, group=None):
    """All-gather equally-shaped tensors; one collective on NCCL (all_gather_into_tensor), the list form elsewhere (gloo)."""
    if t.is_cuda:
        buf = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        return [buf[r] for r in range(world)]
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t.contiguous(), group=group)
    return bufs


def gather_slabs(n_local, kps_local, desc_local, total_frames, group=None):
    """All-gather per-frame result slabs (n[b], kps[b, cap, 7] float32 view, desc[b, cap, 32]) from contiguous frame shards and
    return them in global frame order on every rank.  Shards may differ by one frame: they are padded to the largest."""
    rank, world = _world()
    if world == 1:
        return n_local, kps_local, desc_local
    per = [shard_range(total_frames, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in per)

    def pad(t):
        if t.shape[0] == mx:
            return t.contiguous()
        out = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        return out

    outs = []
    for t in (n_local, kps_local, desc_local):
        buf = _all_gather(pad(t), world, group)
        outs.append(torch.cat([buf[r][:per[r][1] - per[r][0]] for r in range(world)], 0))
    return tuple(outs)


def extract_sharded(extract_fn, frames, cap, group=None):
    """frames: the full [B, H, W] batch (every rank holds it or a view of its own block is enough).  `extract_fn(block)` returns
    (n, kps, desc) torch tensors for a block of frames — on the GPU this is ORBExtractor.extract_batch_device."""
    rank, world = _world()
    lo, hi = shard_range(frames.shape[0], rank, world)
    n, kps, desc = extract_fn(frames[lo:hi])
    return gather_slabs(n, kps, desc, frames.shape[0], group)


def allpairs_sharded(match_fn, desc_local, group=None):
    """Key-frame-window all-pairs: every rank contributes its descriptors, the table is all-gathered, each rank matches its own
    query rows against the whole table (`match_fn(q, table)` -> best_idx, best_dist, second_dist) and the rows are gathered back
    in order.  Shards are padded to the largest with zero rows that are dropped again."""
    rank, world = _world()
    if world == 1:
        return match_fn(desc_local, desc_local)
    counts = torch.zeros(world, dtype=torch.int64, device=desc_local.device)
    counts[rank] = desc_local.shape[0]
    dist.all_reduce(counts, group=group)
    counts = [int(c) for c in counts.tolist()]
    mx = max(counts)
    padded = torch.zeros((mx, 32), dtype=torch.uint8, device=desc_local.device)
    padded[:desc_local.shape[0]] = desc_local
    table = _all_gather(padded, world, group)
    table = torch.cat([table[r][:counts[r]] for r in range(world)], 0).contiguous()
    res = match_fn(desc_local, table)
    outs = []
    for t in res:
        p = torch.zeros(mx, dtype=t.dtype, device=t.device)
        p[:t.shape[0]] = t
        buf = _all_gather(p, world, group)
        outs.append(torch.cat([buf[r][:counts[r]] for r in range(world)], 0))
    return tuple(outs)
