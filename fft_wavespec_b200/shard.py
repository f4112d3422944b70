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


def bar_range_shard(series_len: int, window_len: int, hop: int, rank: int, world: int):
    """One series split over ranks by bar range (fewer series than GPUs, or one dominant series).

    Returns (first_window, n_windows, first_sample, n_samples): rank `rank` owns the windows
    [first_window, first_window + n_windows) of the series and must be handed the samples
    [first_sample, first_sample + n_samples) — its windows plus a halo of window_len - hop samples
    shared with the next rank.  Only the stateless stages (spectra, selection, rows, waves) may be
    split this way; the per-series recursions (Kalman4D, weight-Kalman, tracker pool) stay on one
    rank per series.  Ranks beyond the window count get n_windows = 0."""
    if series_len < window_len:
        return 0, 0, 0, 0
    nwin = 1 + (series_len - window_len) // hop
    base, rem = divmod(nwin, world)
    w0 = rank * base + min(rank, rem)
    cnt = base + (1 if rank < rem else 0)
    if cnt == 0:
        return w0, 0, 0, 0
    a0 = w0 * hop
    return w0, cnt, a0, (cnt - 1) * hop + window_len


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
