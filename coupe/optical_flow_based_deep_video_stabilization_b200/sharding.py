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


def gather_output(local: torch.Tensor, n_items: int, group=None, dst=None, out=None):
    """Collects the per-rank output frames [n_local, H, W, C] of a split clip (or of a set of clips laid end to end) into
    [n_items, H, W, C] in pair order: on every rank (`dst=None`, all_gather) or on rank `dst` only (the other ranks
    return None).  NCCL for CUDA tensors, gloo for CPU tensors; any dtype -- send the np.uint8 frames the video writer
    consumes (2.76 MB per 720p frame) rather than float32 (11.06 MB).  Ranks may own different counts (shard_range);
    shards are padded to the largest count for the collective and trimmed afterwards.  `out`: optional preallocated
    receive buffer [world * max_count, H, W, C] on the receiving rank(s) (a 1024-frame 720p clip set is tens of GB: allocate
    it once, outside the loop that calls this)."""
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
    if dst is not None and not (0 <= dst < world):
        raise ValueError(f"dst {dst} outside [0, {world})")
    nmax = max(h - l for l, h in counts)
    padded = local
    if local.shape[0] < nmax:
        padded = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    padded = padded.contiguous()
    even = all(h - l == nmax for l, h in counts)
    full_shape = (world * nmax,) + tuple(local.shape[1:])
    if out is not None and (tuple(out.shape) != full_shape or out.dtype != local.dtype or out.device != local.device
                            or not out.is_contiguous()):
        raise ValueError(f"out must be a contiguous {local.dtype} tensor of shape {full_shape} on {local.device}")
    if dst is None:
        if out is None:
            out = local.new_empty(full_shape)
        dist.all_gather_into_tensor(out, padded, group=group)      # one flat receive buffer: no per-rank list, no concat
    else:
        if rank != dst:
            out = None
        elif out is None:
            out = local.new_empty(full_shape)
        parts = list(out.view((world, nmax) + tuple(local.shape[1:])).unbind(0)) if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
    if even:
        return out
    return torch.cat([out[r * nmax: r * nmax + (h - l)] for r, (l, h) in enumerate(counts)], 0)
