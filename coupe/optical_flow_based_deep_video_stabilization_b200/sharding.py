"""Clip / frame-pair sharding across the GPUs of one box.

The reference is single-process (one tf.Session, main_dl.py:489) and treats every video independently
(main_dl.py:470-476).  Frame pairs -- one [384,512,27] network input plus one full-resolution frame -- are the
independent units of the hot path, so the N-GPU form is N replicas: one process per GPU, weights replicated,
pair index ranges split contiguously, NO collective on the inference path.  The only exchange is the optional
gather of the stabilised frames when one clip's pair range was split over ranks.
"""
from __future__ import annotations

import torch


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [lo, hi) of `n_items` units owned by `rank`: the first n_items % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request: n_items={n_items} rank={rank} world={world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_output(local: torch.Tensor, n_items: int, group=None):
    """All-gathers the per-rank output frames [n_local, H, W, C] of a split clip into [n_items, H, W, C] in pair
    order on every rank (NCCL for CUDA tensors, gloo for CPU tensors).  Ranks may own different counts
    (shard_range); shards are padded to the largest count for the collective and trimmed afterwards."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_items:
            raise ValueError("single process: the local shard must be the whole clip")
        return local
    world = dist.get_world_size(group)
    counts = [shard_range(n_items, r, world) for r in range(world)]
    rank = dist.get_rank(group)
    lo, hi = counts[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns pairs [{lo},{hi}) but holds {local.shape[0]} frames")
    nmax = max(h - l for l, h in counts)
    padded = local
    if local.shape[0] < nmax:
        padded = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[: h - l] for p, (l, h) in zip(parts, counts)], 0)
