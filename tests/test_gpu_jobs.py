"""GPU tests of the job table behind the imports.mqh submit / try_get / free calls
(Include/imports.mqh:12-19): chunked non-blocking delivery into pageable and page-locked caller
buffers, the cycle-cache record as a job product (WaveSpecZZ_1.1.0-gpuopt.mq5:1067-1102, :294-324),
and jobs spread over several devices."""
import os
import time

import numpy as np
import pytest

from fft_wavespec_b200 import synth

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True)
def small_chunks(monkeypatch):
    monkeypatch.setenv("WAVESPEC_JOB_CHUNK", "1000")      # several chunks per job at test sizes


@pytest.fixture(scope="module")
def br():
    from fft_wavespec_b200 import bridge
    st = bridge.gpu_init(0, 8)
    assert st == bridge.OK, bridge.last_error()
    yield bridge
    bridge.gpu_shutdown()


def pinned(n):
    import torch
    return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()


def poll(br, getter, jid, out, tries=20000):
    for _ in range(tries):
        st, n, ready = getter(jid, out)
        if st == br.OK and ready == 1:
            return n
        assert st == br.OK and ready == 0, (st, br.last_error())       # the Fetcher's "still running"
        time.sleep(0.0005)
    raise AssertionError("job did not finish")


@pytest.mark.parametrize("page_locked", [False, True])
@pytest.mark.parametrize("n,k,bars", [(1024, 8, 6000), (256, 2, 3456), (4096, 4, 4096 + 2500)])
def test_batch_rows_chunked_delivery(br, oracle, n, k, bars, page_locked):
    s = synth.random_walk(300 + n, bars)
    nwin = bars - n + 1
    out = pinned(nwin * k * 15) if page_locked else np.empty(nwin * k * 15)
    out[:] = -7.0
    st, jid = br.gpu_submit_extract_cycles_batch(s, n, 1, k, 9.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK and jid != 0, br.last_error()
    assert poll(br, br.gpu_try_get_cycles_batch, jid, out) == nwin * k
    # polling again with the same buffer is idempotent
    st, cnt, ready = br.gpu_try_get_cycles_batch(jid, out)
    assert (st, cnt, ready) == (br.OK, nwin * k, 1)
    assert br.gpu_free_job(jid) == br.OK
    cfg = oracle.default_cfg(n, top_k=k, min_period=9.0, max_period=200.0, sample_rate_seconds=60.0)
    ref = oracle.pipeline_series(s, cfg, oracle.OUT_ROWS | oracle.OUT_BINS)
    rows = out.reshape(nwin, k, 15)
    assert np.array_equal(np.rint(n / rows[..., 2]).astype(int), ref["bins"])      # period = N / bin is exact
    for f in (0, 1, 2, 6):
        assert np.abs(rows[..., f] - ref["rows"][..., f]).max() <= 1e-9 * np.abs(ref["rows"][..., f]).max()


def test_batch_capacity_smaller_than_the_result_gives_whole_windows(br, oracle):
    n, k, bars = 1024, 8, 5000
    s = synth.random_walk(17, bars)
    nwin = bars - n + 1
    st, jid = br.gpu_submit_extract_cycles_batch(s, n, 1, k, 18.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK
    cap_rows = 2500 * k + 3                                # not a whole number of windows
    out = np.full(cap_rows * 15, -7.0)
    got = poll(br, br.gpu_try_get_cycles_batch, jid, out)
    assert got == 2500 * k
    assert np.all(out[got * 15:] == -7.0)                  # nothing written past the whole windows
    full = np.empty(nwin * k * 15)                         # a second, larger buffer re-arms the delivery
    assert poll(br, br.gpu_try_get_cycles_batch, jid, full) == nwin * k
    assert np.array_equal(full[:got * 15], out[:got * 15])
    assert br.gpu_free_job(jid) == br.OK


@pytest.mark.parametrize("page_locked", [False, True])
@pytest.mark.parametrize("hop,top_k,music_only,weights", [(1, 2, False, False), (1, 8, False, False),
                                                           (3, 4, False, False), (1, 2, True, False),
                                                           (1, 4, False, True)])
def test_cycle_cache_record_as_job_product(br, oracle, hop, top_k, music_only, weights, page_locked):
    n, bars = 512, 5300
    s = synth.random_walk(40 + hop + top_k, bars)
    kw = dict(music_only=music_only, use_music_weights=weights)
    st, jid = br.submit_cycle_cache_batch(s, n, hop, top_k, 9.0, 200.0, 60.0, 0, 10, **kw)
    assert st == br.OK and jid != 0, br.last_error()
    out = pinned(bars * 20) if page_locked else np.empty(bars * 20)
    assert poll(br, br.try_get_cycle_cache, jid, out) == bars
    assert br.gpu_free_job(jid) == br.OK
    cfg = oracle.default_cfg(n, hop=hop, top_k=top_k, min_period=9.0, max_period=200.0, sample_rate_seconds=60.0)
    rows = oracle.pipeline_series(s, cfg, oracle.OUT_ROWS)["rows"]
    ref = oracle.cycle_cache(rows, top_k, n, hop, bars, 60.0, **kw)
    rec = out.reshape(bars, 20)
    empty = ref == np.finfo(np.float64).max
    assert np.array_equal(rec == np.finfo(np.float64).max, empty)
    scale = np.abs(np.where(empty, 0.0, ref)).max(axis=0)
    err = np.abs(np.where(empty, 0.0, rec - ref)).max(axis=0)
    # wave, period, eta, energy and the MUSIC fields: 1e-9 of the column scale; theta (cols 6, 7) is an
    # angle that back-propagates over up to N bars: compared on the circle
    for col in range(20):
        if col in (6, 7):
            d = np.angle(np.exp(1j * np.where(empty[:, col], 0.0, rec[:, col] - ref[:, col])))
            assert np.abs(d).max() < 1e-7
        else:
            assert err[col] <= 1e-9 * max(scale[col], 1e-300), col


def test_cycle_cache_job_matches_the_rows_job_decoded_on_the_device(br):
    """Same series through both products: decoding the delivered rows gives the delivered record."""
    n, bars, k = 1024, 7000, 2
    s = synth.random_walk(99, bars)
    st, j1 = br.gpu_submit_extract_cycles_batch(s, n, 1, k, 18.0, 200.0, 60.0, 0, 10, 15)
    st2, j2 = br.submit_cycle_cache_batch(s, n, 1, k, 18.0, 200.0, 60.0, 0, 10)
    assert st == br.OK and st2 == br.OK
    rows = np.empty((bars - n + 1) * k * 15)
    rec = np.empty(bars * 20)
    poll(br, br.gpu_try_get_cycles_batch, j1, rows)
    poll(br, br.try_get_cycle_cache, j2, rec)
    br.gpu_free_job(j1); br.gpu_free_job(j2)
    ref = br.cycle_cache_host(rows.reshape(-1, k, 15), k, n, 1, bars)
    assert np.array_equal(rec.reshape(bars, 20), ref)


def test_window_jobs_answer_not_ready_and_deliver_from_the_job(br, oracle):
    cfg = oracle.default_cfg(1024, top_k=2, min_period=9.0, max_period=200.0)
    jobs = []
    for i in range(64):                                    # InpAsyncDepth = 64 (1.1.0 :62)
        st, jid = br.gpu_submit_extract_cycles(synth.random_walk(500 + i, 1024), 2, 9.0, 200.0, 60.0, 1, 10)
        assert st == br.OK and jid != 0
        jobs.append(jid)
    buf = np.zeros((2, 15))
    for i, jid in enumerate(jobs):
        for _ in range(100000):
            st, n, ready = br.gpu_try_get_cycles(jid, buf, 15, 2)
            if st != br.NOT_READY:                         # the only "running" answer :1342-1374 accepts
                break
            assert ready == 0
        assert (st, n, ready) == (br.OK, 2, 1)
        ref = oracle.pipeline_series(synth.random_walk(500 + i, 1024), cfg, oracle.OUT_BINS)
        assert np.array_equal(np.rint(1024 / buf[:, 2]).astype(int), ref["bins"][0])
        assert br.gpu_free_job(jid) == br.OK


def test_wrong_kind_and_unknown_ids(br):
    s = synth.random_walk(3, 3000)
    st, jid = br.gpu_submit_extract_cycles_batch(s, 1024, 1, 4, 9.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK
    st, n, ready = br.try_get_cycle_cache(jid, np.empty(3000 * 20))
    assert st == br.BAD_ARGS and ready == 0
    st, n, ready = br.gpu_try_get_cycles(jid, np.empty((4, 15)), 15, 4)
    assert st == br.BAD_ARGS
    assert br.gpu_free_job(jid) == br.OK
    assert br.gpu_free_job(jid) == br.BAD_ARGS
    assert br.job_device(jid) == -1


def test_jobs_spread_over_every_device(oracle):
    """gpu_init(-1): one session per device, jobs bound round robin, results identical on each."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    from fft_wavespec_b200 import bridge as br
    br.gpu_shutdown()
    assert br.gpu_init(-1, 8) == br.OK, br.last_error()
    try:
        ndev = br.device_count()
        assert ndev == torch.cuda.device_count()
        n, k, bars = 1024, 4, 20000
        series = [synth.random_walk(700 + i, bars) for i in range(2 * ndev)]
        jobs, outs = [], []
        for s in series:
            st, jid = br.gpu_submit_extract_cycles_batch(s, n, 1, k, 9.0, 200.0, 60.0, 0, 10, 15)
            assert st == br.OK, br.last_error()
            jobs.append(jid)
            outs.append(np.empty((bars - n + 1) * k * 15))
        assert sorted(br.job_device(j) for j in jobs) == sorted(list(range(ndev)) * 2)
        cfg = oracle.default_cfg(n, top_k=k, min_period=9.0, max_period=200.0)
        for s, jid, out in zip(series, jobs, outs):
            poll(br, br.gpu_try_get_cycles_batch, jid, out)
            assert br.gpu_free_job(jid) == br.OK
            ref = oracle.pipeline_series(s, cfg, oracle.OUT_BINS)
            rows = out.reshape(-1, k, 15)
            assert np.array_equal(np.rint(n / rows[..., 2]).astype(int), ref["bins"])
        # synchronous calls run on the first opened device
        x = synth.random_walk(1, 1024)
        assert np.abs(br.gpu_fft_real_forward(x) - oracle.fft_interleaved(x)).max() < 1e-6
    finally:
        br.gpu_shutdown()
        assert br.gpu_init(0, 8) == br.OK


def test_free_and_shutdown_with_jobs_in_flight(br, oracle):
    """OnDeinit frees whatever is still queued and shuts the session down (WaveSpecZZ_1.1.0-gpuopt.mq5:
    698-720): neither may wait for, or trip over, work that is still running; a new session works."""
    s = synth.random_walk(61, 400_000)
    jobs = []
    for _ in range(6):
        st, jid = br.gpu_submit_extract_cycles_batch(s, 1024, 1, 8, 18.0, 200.0, 60.0, 0, 10, 15)
        assert st == br.OK
        jobs.append(jid)
    out = pinned((400_000 - 1023) * 8 * 15)
    br.gpu_try_get_cycles_batch(jobs[0], out)              # arm one of them: copies in flight into `out`
    for jid in jobs[:3]:
        assert br.gpu_free_job(jid) == br.OK               # freed while running (job 0 with an armed buffer)
    br.gpu_shutdown()                                      # three jobs still in the table
    st, n, ready = br.gpu_try_get_cycles_batch(jobs[4], out)
    assert st == br.BAD_ARGS                               # ids of the old session are gone (and the
    assert br.gpu_init(0, 8) == br.OK                      # library says so instead of crashing)
    x = synth.random_walk(62, 3000)
    st, jid = br.gpu_submit_extract_cycles_batch(x, 1024, 1, 4, 9.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK
    res = np.empty((3000 - 1023) * 4 * 15)
    assert poll(br, br.gpu_try_get_cycles_batch, jid, res) == (3000 - 1023) * 4
    assert br.gpu_free_job(jid) == br.OK
    cfg = oracle.default_cfg(1024, top_k=4, min_period=9.0, max_period=200.0)
    ref = oracle.pipeline_series(x, cfg, oracle.OUT_BINS)
    assert np.array_equal(np.rint(1024 / res.reshape(-1, 4, 15)[..., 2]).astype(int), ref["bins"])


def test_two_host_threads_share_the_session(br, oracle):
    """Several indicator instances and the Fetcher can live in one process (threading):
    two host threads submit, poll and free their own jobs at the same time."""
    import threading
    errors = []

    def client(seed):
        try:
            cfg = oracle.default_cfg(512, top_k=4, min_period=9.0, max_period=100.0)
            for r in range(6):
                s = synth.random_walk(seed + r, 512 + 2500)
                st, jid = br.gpu_submit_extract_cycles_batch(s, 512, 1, 4, 9.0, 100.0, 60.0, 0, 10, 15)
                assert st == br.OK, br.last_error()
                out = np.empty(2501 * 4 * 15)
                assert poll(br, br.gpu_try_get_cycles_batch, jid, out) == 2501 * 4
                assert br.gpu_free_job(jid) == br.OK
                ref = oracle.pipeline_series(s, cfg, oracle.OUT_BINS)
                assert np.array_equal(np.rint(512 / out.reshape(-1, 4, 15)[..., 2]).astype(int), ref["bins"])
                w = br.gpu_fft_real_forward(s[:512])           # synchronous calls interleave with the jobs
                assert np.abs(w - oracle.fft_interleaved(s[:512])).max() <= 1e-9 * np.abs(w).max()
        except Exception as e:                                 # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=client, args=(8000 + 100 * i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
