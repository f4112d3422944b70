"""GPU parity at size: BASELINE configs 2, 4 and 5 on multi-series batches large enough that every
tile shape, launch group and window-range chunk of the production path is exercised, with EVERY
window compared against the oracle run on all host threads (oracle.pipeline_batch_mt), plus the
batched inverse FFT / top-K wave reconstruction (A8d) against oracle_fft_inverse.

References: Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-568 (selection +
reconstruction), Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1183-1263 (phase chain),
Legacy/WaveSpecZZ_1.0.4-kalman.mq5:182-192 (A8b), WaveCyclesBatchFetcher.mq5:28-36 (config 5)."""
import ctypes as C
import json
import os
import threading
import time

import numpy as np
import pytest

from fft_wavespec_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9
REPORT = {}
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def br():
    from fft_wavespec_b200 import bridge
    st = bridge.gpu_init(0, 8)
    assert st == bridge.OK, bridge.last_error()
    yield bridge
    bridge.gpu_shutdown()
    print("\nFULL-SIZE PARITY REPORT " + json.dumps(REPORT, sort_keys=True))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report_fullsize.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)


def ocfg_from(oracle, cfg):
    o = oracle.PipelineCfg()
    C.memmove(C.byref(o), C.byref(cfg), C.sizeof(o))
    return o


def near_ties_sampled(series, n, lo, hi, k, step=97, gap=1e-11):
    """near ties between the k-th and (k+1)-th in-band power on every `step`-th window (numpy rfft:
    a report, not a parity check)."""
    idx = np.arange(0, series.size - n + 1, step)
    w = np.lib.stride_tricks.sliding_window_view(series, n)[idx]
    p = np.abs(np.fft.rfft(w, axis=1)[:, lo:hi + 1]) ** 2
    p = np.sort(p, axis=1)[:, ::-1]
    return int(np.sum(np.abs(p[:, k - 1] - p[:, k]) <= gap * p[:, k - 1])), int(idx.size)


def test_config2_4x200k_every_window(br, oracle):
    """Config 2 (N=1024, plain hop-1, top-8, band 18-200) on 4 series x 200k bars: every one of the
    795 908 windows, bins exact, amplitude / energy ratio / last-sample waves within 1e-9."""
    n, k, ns, bars = 1024, 8, 4, 200_000
    s = synth.random_walk_batch(2000, ns, bars)
    cfg = br.default_cfg(n, top_k=k, min_period=18.0, max_period=200.0)
    got = br.pipeline_host(s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES | br.OUT_SPECTRA)
    assert br.last_kernel() in ("sliding_overlap", "sliding_staged")   # the kernel bench.py times
    done, ref = oracle.pipeline_batch_mt(s, ocfg_from(oracle, cfg), THREADS, want=("bins", "rows", "waves"))
    assert done == ns * (bars - n + 1)
    assert np.array_equal(got["bins"], ref["bins"]), "selected cycle bins differ"
    for f in (0, 1, 2, 6):
        assert np.abs(got["rows"][..., f] - ref["rows"][..., f]).max() <= REL_TOL * np.abs(ref["rows"][..., f]).max()
    assert np.abs(got["waves"] - ref["waves"]).max() <= REL_TOL * np.abs(ref["waves"]).max()
    # spectra plane: every 499th window of every series against the oracle's FFT
    worst = 0.0
    for i in range(ns):
        for w in range(0, bars - n + 1, 499):
            r = oracle.fft_interleaved(s[i, w:w + n])
            worst = max(worst, float(np.abs(got["spectra"][i, w] - r).max() / np.abs(r).max()))
    assert worst < REL_TOL
    ties = [near_ties_sampled(s[i], n, 6, 56, k) for i in range(ns)]
    REPORT["config2_windows"] = int(done)
    REPORT["config2_near_ties_in_sample"] = [sum(t[0] for t in ties), sum(t[1] for t in ties)]
    REPORT["config2_max_rel_err_spectra_sampled"] = worst


def test_config4_8x100k_phase_top8_reconstruction(br, oracle):
    """Config 4 = config 2 + phase chain (A6) + A8a / A8b for the top-8 cycles: 8 series x 100k bars.
    Bins, rows, waves and the A8b contribution plane on every window; the three phase planes on two
    whole series (the oracle's atan2 chain is the slow part: one host thread per series)."""
    n, k, ns, bars = 1024, 8, 8, 100_000
    nwin = bars - n + 1
    s = synth.random_walk_batch(2100, ns, bars)
    cfg = br.default_cfg(n, top_k=k, min_period=18.0, max_period=200.0)
    got = br.pipeline_host(s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES | br.OUT_CONTRIB)
    done, ref = oracle.pipeline_batch_mt(s, ocfg_from(oracle, cfg), THREADS, want=("bins", "rows", "waves"))
    assert done == ns * nwin
    assert np.array_equal(got["bins"], ref["bins"])
    for f in (0, 1, 2, 6):
        assert np.abs(got["rows"][..., f] - ref["rows"][..., f]).max() <= REL_TOL * np.abs(ref["rows"][..., f]).max()
    wscale = np.abs(ref["waves"]).max()
    assert np.abs(got["waves"] - ref["waves"]).max() <= REL_TOL * wscale
    # A8b is algebraically 2 x A8a (one-sided vs two-sided amplitude): every window against that, and
    # oracle_contribution itself on a strided sample
    assert np.abs(got["contrib"] - 2.0 * ref["waves"]).max() <= REL_TOL * 2.0 * wscale
    worst = 0.0
    for i in range(ns):
        for w in range(0, nwin, 997):
            re, im = oracle.fft_forward(s[i, w:w + n])
            for j in range(k):
                c = oracle.contribution(re, im, int(got["bins"][i, w, j]))
                worst = max(worst, abs(got["contrib"][i, w, j] - c))
    assert worst <= REL_TOL * 2.0 * wscale
    REPORT["config4_windows"] = int(done)
    REPORT["config4_max_abs_err_contrib_vs_oracle_contribution"] = worst
    del got, ref

    # phase planes: two whole series, 99k windows each, GPU in one call, oracle one thread per series
    ps = s[:2]
    pcfg = br.default_cfg(n, top_k=k, min_period=18.0, max_period=200.0)
    gp = br.pipeline_host(ps, pcfg, br.OUT_PHASE)["phase"]
    refs = [None, None]

    def run(i):
        refs[i] = oracle.pipeline_series(ps[i], ocfg_from(oracle, pcfg), oracle.OUT_PHASE | oracle.OUT_SPECTRA)
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    on_decision = total = 0
    worst_ph = 0.0
    for i in range(2):
        rp, sp = refs[i]["phase"], refs[i]["spectra"]
        mag = np.hypot(sp[:, 0::2], sp[:, 1::2])
        dph = np.abs(np.angle(np.exp(1j * (gp[i][:, 0] - rp[:, 0]))))
        tol = REL_TOL * mag.max(axis=1, keepdims=True) / np.where(mag > 0, mag, np.inf)
        tol = np.where(mag > 0, np.maximum(tol, 1e-15), np.pi)
        assert np.all(dph <= tol), float((dph - tol).max())
        worst_ph = max(worst_ph, float(dph.max()))
        d = np.abs(np.abs(np.diff(rp[:, 0], axis=1)) - np.pi)
        safe = d.min(axis=1) > 1e-6
        on_decision += int((~safe).sum()); total += int(safe.size)
        assert np.abs(gp[i][safe, 1] - rp[safe, 1]).max() < 1e-9 * max(1.0, np.abs(rp[safe, 1]).max())
        assert np.abs(gp[i][safe, 2] - rp[safe, 2]).max() < 1e-7
        if (~safe).any():                                    # the other branch is 2 pi away, bin by bin
            dd = np.angle(np.exp(1j * (gp[i][~safe, 1] - rp[~safe, 1])))
            assert np.abs(dd).max() < 1e-6
    assert on_decision < 0.1 * total
    REPORT["config4_phase_windows"] = total
    REPORT["config4_phase_windows_on_a_pi_decision"] = on_decision
    REPORT["config4_max_phase_err"] = worst_ph


def test_config5_4x100k_n4096_through_the_batch_api(br, oracle):
    """Config 5 (WaveCyclesBatchFetcher.mq5:28-36: N=4096, K=4, band 9-200, stride 15) on 4 series x
    100k bars through gpu_submit_extract_cycles_batch / try_get / free, every window's bins exact."""
    n, k, ns, bars = 4096, 4, 4, 100_000
    nwin = bars - n + 1
    s = synth.random_walk_batch(2200, ns, bars)
    jobs, outs = [], []
    for i in range(ns):
        st, jid = br.gpu_submit_extract_cycles_batch(s[i], n, 1, k, 9.0, 200.0, 60.0, 0, 10, 15)
        assert st == br.OK and jid != 0, br.last_error()
        jobs.append(jid)
        outs.append(np.empty(nwin * k * 15))
    for jid, out in zip(jobs, outs):
        for _ in range(4000):                                # the Fetcher's poll budget (:127-132)
            st, cnt, ready = br.gpu_try_get_cycles_batch(jid, out)
            if st == br.OK and ready == 1:
                break
            assert st == br.OK and ready == 0
            time.sleep(0.005)
        assert ready == 1 and cnt == nwin * k
        assert br.gpu_free_job(jid) == br.OK
    cfg = oracle.default_cfg(n, top_k=k, min_period=9.0, max_period=200.0, sample_rate_seconds=60.0)
    done, ref = oracle.pipeline_batch_mt(s, cfg, THREADS, want=("bins", "rows"))
    assert done == ns * nwin
    rows = np.stack([o.reshape(nwin, k, 15) for o in outs])
    assert np.array_equal(np.rint(n / rows[..., 2]).astype(np.int32), ref["bins"])      # period = N / bin is exact
    for f in (0, 1, 2, 6):
        assert np.abs(rows[..., f] - ref["rows"][..., f]).max() <= REL_TOL * np.abs(ref["rows"][..., f]).max()
    assert np.all(rows[..., 14] == 0.0)
    ties = [near_ties_sampled(s[i], n, 21, 455, k, step=397) for i in range(ns)]
    REPORT["config5_windows"] = int(done)
    REPORT["config5_near_ties_in_sample"] = [sum(t[0] for t in ties), sum(t[1] for t in ties)]


# ---- A8d: batched inverse real FFT and top-K wave reconstruction ---------------------------------
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_inverse_fft_matches_oracle(br, oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    spec = br.gpu_fft_real_forward(x)
    back = br.gpu_fft_real_inverse(spec)                     # the Legacy single-window symbol
    ref = oracle.fft_inverse(oracle.fft_interleaved(x))
    assert np.abs(back - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    nyq = np.sum(x * (-1.0) ** np.arange(n))                 # the forward contract drops the Nyquist bin
    assert np.abs(back - (x - nyq * (-1.0) ** np.arange(n) / n)).max() < 1e-12


@pytest.mark.parametrize("n,nwin", [(256, 3001), (512, 2000), (1024, 1501), (2048, 700), (4096, 333)])
def test_inverse_fft_batch_on_price_spectra(br, oracle, n, nwin):
    """The spectra plane of the forward pipeline fed back through the batched inverse: every window
    against oracle_fft_inverse of the oracle's own spectrum, and against the window itself."""
    s = synth.random_walk(3000 + n, n + nwin - 1)
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
    spec = br.pipeline_host(s, cfg, br.OUT_SPECTRA)["spectra"]
    back = br.fft_real_inverse_batch(spec, n)
    assert back.shape == (nwin, n)
    win = np.lib.stride_tricks.sliding_window_view(s, n)
    sign = (-1.0) ** np.arange(n)
    nyq = (win * sign).sum(axis=1, keepdims=True)
    assert np.abs(back - (win - nyq * sign / n)).max() < 1e-9 * np.abs(win).max()
    for w in range(0, nwin, max(1, nwin // 40)):
        ref = oracle.fft_inverse(oracle.fft_interleaved(s[w:w + n]))
        assert np.abs(back[w] - ref).max() <= REL_TOL * np.abs(ref).max()


@pytest.mark.parametrize("n,k", [(1024, 8), (512, 5), (4096, 4), (256, 2)])
def test_topk_wave_reconstruction(br, oracle, n, k):
    """Inverse FFT of the spectrum masked to the selected cycles (north star: "inverse-FFT wave
    reconstruction of the top-8 cycles"): whole-window waves against oracle_fft_inverse of the masked
    oracle spectrum; their last sample is the sum of the A8b contributions."""
    nwin = 700
    s = synth.random_walk(3100 + n, n + nwin - 1)
    cfg = br.default_cfg(n, top_k=k, min_period=9.0, max_period=200.0)
    got = br.pipeline_host(s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_CONTRIB)
    waves = br.reconstruct_topk(got["spectra"], got["bins"])
    assert waves.shape == (nwin, n)
    assert np.abs(waves[:, -1] - got["contrib"].sum(axis=1)).max() <= 1e-9 * np.abs(got["contrib"]).sum(axis=1).max()
    ocfg = oracle.default_cfg(n, top_k=k, min_period=9.0, max_period=200.0)
    for w in range(0, nwin, 23):
        ref = oracle.pipeline_series(s[w:w + n], ocfg, oracle.OUT_SPECTRA | oracle.OUT_BINS)
        assert np.array_equal(ref["bins"][0], got["bins"][w])
        masked = np.zeros(n)
        for b in ref["bins"][0]:
            masked[2 * b:2 * b + 2] = ref["spectra"][0][2 * b:2 * b + 2]
        r = oracle.fft_inverse(masked)
        assert np.abs(waves[w] - r).max() <= REL_TOL * np.abs(r).max()


def test_inverse_bad_args(br):
    with pytest.raises(br.WaveSpecError) as e:
        br.gpu_fft_real_inverse(np.zeros(1000))
    assert e.value.status == br.BAD_ARGS
    with pytest.raises(br.WaveSpecError):
        br.gpu_fft_real_inverse(np.zeros(2))
