"""Series sharding across GPUs.

The unit of work is a series (symbol x timeframe); series share nothing, so ranks own disjoint
series and there is no collective on the data path.  torch.distributed is used only for the
timing barrier, the max-over-ranks reduction and (optionally) a host gather of small results.
"""
from __future__ import annotations


def weak_shard(rank: int, series_per_rank: int):
    """Global series indices owned by `rank` when every rank gets the same count (bench.py)."""
    first = rank * series_per_rank
    return range(first, first + series_per_rank)


def strong_shard(total_series: int, rank: int, world: int):
    """Contiguous balanced blocks of a fixed total (WaveCyclesBatchFetcher-style sweep)."""
    base, rem = divmod(total_series, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return range(first, first + count)


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local, total_series: int, world: int, rank: int):
    """Host gather of per-series results (numpy arrays keyed by global series index)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return dict(local)
    parts = [None] * world
    dist.all_gather_object(parts, dict(local))
    out = {}
    for p in parts:
        out.update(p)
    assert len(out) == total_series, "series lost or duplicated by the sharding"
    return out
