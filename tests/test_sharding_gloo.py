"""world_size-2 gloo test of the multi-GPU host logic (series sharding, max-over-ranks timing,
host gather).  The data path has no collective; each rank here stands in for one GPU and runs
the oracle on its shard, and the gathered result must equal the unsharded run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fft_wavespec_b200 import shard, synth
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = orc.default_cfg(256, top_k=4, min_period=9.0, max_period=100.0)
    local = {}
    for s in shard.strong_shard(total, rank, world):
        local[s] = orc.pipeline_series(synth.random_walk(s, 400), cfg, orc.OUT_BINS)["bins"]
    merged = shard.gather_rows(local, total, world, rank)
    tmax = shard.max_over_ranks(1.0 + rank)
    dist.barrier()
    if rank == 0:
        q.put((sorted(merged.keys()), {k: v.tolist() for k, v in merged.items()}, tmax))
    dist.destroy_process_group()


def test_two_rank_series_sharding_matches_single_rank():
    total, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    keys, merged, tmax = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert keys == list(range(total))
    assert tmax == 2.0                                  # max over ranks, not rank 0's own time
    sys.path.insert(0, ROOT)
    from fft_wavespec_b200 import synth
    from oracle import oracle as orc
    cfg = orc.default_cfg(256, top_k=4, min_period=9.0, max_period=100.0)
    for s in range(total):
        ref = orc.pipeline_series(synth.random_walk(s, 400), cfg, orc.OUT_BINS)["bins"]
        assert np.array_equal(np.array(merged[s]), ref)


def test_shard_helpers_cover_every_series_once():
    from fft_wavespec_b200 import shard
    for total in (1, 7, 64, 4000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += list(shard.strong_shard(total, r, world))
            assert seen == list(range(total))
    assert list(shard.weak_shard(3, 64))[:2] == [192, 193]


def test_bar_range_shard_covers_every_window_once():
    from fft_wavespec_b200 import shard
    for series_len, n, hop in ((5000, 1024, 1), (5000, 512, 7), (1024, 1024, 1), (3000, 2048, 3), (100, 256, 1)):
        nwin = 0 if series_len < n else 1 + (series_len - n) // hop
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                w0, cnt, a0, na = shard.bar_range_shard(series_len, n, hop, r, world)
                assert w0 == nxt or cnt == 0
                nxt += cnt
                if cnt:
                    assert a0 == w0 * hop and a0 + na <= series_len
                    assert na == (cnt - 1) * hop + n              # own windows + halo, nothing more
            assert nxt == nwin


def _range_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fft_wavespec_b200 import shard, synth
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = synth.random_walk(11, 1500)
    cfg = orc.default_cfg(256, top_k=4, min_period=9.0, max_period=100.0, hop=3)
    w0, cnt, a0, na = shard.bar_range_shard(x.size, 256, 3, rank, world)
    part = orc.pipeline_series(x[a0:a0 + na], cfg, orc.OUT_BINS | orc.OUT_SPECTRA)
    merged = shard.gather_rows({rank: (w0, part["bins"], part["spectra"])}, world, world, rank)
    dist.barrier()
    if rank == 0:
        q.put({k: (v[0], v[1].tolist(), v[2].tolist()) for k, v in merged.items()})
    dist.destroy_process_group()


def test_two_rank_bar_range_split_of_one_series_matches_whole_series():
    """One series, two ranks: each runs its window range (with the N - hop halo) and the
    concatenation equals the unsplit run bit for bit — windows are independent of their tile."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_range_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from fft_wavespec_b200 import synth
    from oracle import oracle as orc
    x = synth.random_walk(11, 1500)
    cfg = orc.default_cfg(256, top_k=4, min_period=9.0, max_period=100.0, hop=3)
    ref = orc.pipeline_series(x, cfg, orc.OUT_BINS | orc.OUT_SPECTRA)
    bins = np.concatenate([np.array(merged[r][1]) for r in range(world)])
    spec = np.concatenate([np.array(merged[r][2]) for r in range(world)])
    assert [merged[r][0] for r in range(world)] == [0, (ref["bins"].shape[0] + 1) // 2]
    assert np.array_equal(bins, ref["bins"])
    assert np.array_equal(spec, ref["spectra"])
