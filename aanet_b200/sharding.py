"""Batch sharding of stereo pairs across ranks (one process per GPU, no data-path collective).

The hot path is per-image: every rank owns a contiguous slice of the batch and replicates
the 3.9-8.4 M parameters once.  This replaces the reference's per-forward nn.DataParallel
scatter/replicate/gather (inference.py:131-133).  The only collective is the optional result gather /
timing reduction, which works over any torch.distributed backend (NCCL on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) slice of `n_items` for `rank`; the first n % world ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every tensor of a (nested) list along dim 0 to this rank's share."""
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(shard_batch(t, rank, world_size) for t in tensors)
    lo, hi = shard_range(tensors.shape[0], rank, world_size)
    return tensors[lo:hi]


def gather_batch(local, n_items, group=None):
    """All-gather per-rank result slices (possibly of unequal length) back into batch order."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def max_over_ranks(value, device, group=None):
    """Max of a python float over all ranks (used for max-over-ranks timing)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
