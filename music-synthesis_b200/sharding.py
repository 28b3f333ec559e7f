"""Clip sharding for multi-GPU runs.

Clips are independent, so inference shards the batch contiguously across ranks with NO
data-path collective: rank r of W takes clips [r*ceil(B/W), ...).  `gather_clips` is only for
callers that want the whole batch back on every rank (evaluation); the hot path never calls it.
"""
import torch


def clip_shard(num_clips, rank, world_size):
    """Contiguous, balanced split: the first (num_clips % world_size) ranks get one extra clip.
    Returns (lo, hi)."""
    if world_size < 1 or not (0 <= rank < world_size) or num_clips < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(num_clips, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_clips(local, num_clips, group=None):
    """All-gather ragged per-rank shards (B_r, ...) back into (num_clips, ...), in clip order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [clip_shard(num_clips, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    lo, hi = sizes[rank]
    assert hi - lo == local.shape[0], "local shard does not match clip_shard()"
    return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)], dim=0)
