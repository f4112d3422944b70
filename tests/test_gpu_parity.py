"""GPU parity tests: libwavespec.so (through the C ABI, host buffers) against the CPU oracle on
the same seeded inputs.

Bars (north_star): selected cycle bins and PLA pivot indices bit-exact; spectra and
reconstructed waves within 1e-9 relative (max |dX| / max |X| per window)."""
import ctypes as C

import os

import numpy as np
import pytest

from fft_wavespec_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9

# Counts the assertions cannot express (windows excluded by a decision-boundary filter, near ties,
# observed maxima); printed at the end of the module and written to gpurun_out/parity_report.json.
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def parity_report():
    yield
    import json
    print("\nPARITY REPORT " + json.dumps(REPORT, sort_keys=True))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)


def count_near_ties(name, spectra, lo, hi, k):
    """near ties (k-th vs (k+1)-th in-band power closer than 1e-11 relative) are reported, never hidden
    (SURVEY.md 7.3 item 3): a flipped selection there would be a rounding artefact, not a bug."""
    c = near_ties(spectra.reshape(-1, spectra.shape[-1]), lo, hi, k)
    REPORT["near_ties_" + name] = REPORT.get("near_ties_" + name, 0) + c
    REPORT["near_tie_windows_checked_" + name] = REPORT.get("near_tie_windows_checked_" + name, 0) + int(np.prod(spectra.shape[:-1]))
    return c


@pytest.fixture(scope="module")
def br():
    from fft_wavespec_b200 import bridge
    st = bridge.gpu_init(0, 8)
    assert st == bridge.OK, bridge.last_error()
    yield bridge
    bridge.gpu_shutdown()


def rel_err(got, ref):
    """max |got-ref| / max |ref| per window (last axis)."""
    den = np.abs(ref).max(axis=-1)
    den = np.where(den == 0, 1.0, den)
    return (np.abs(got - ref).max(axis=-1) / den).max()


def near_ties(spectra, lo, hi, k, gap=1e-11):
    """count windows whose k-th/(k+1)-th in-band powers are closer than `gap` relative."""
    p = spectra[:, 0::2] ** 2 + spectra[:, 1::2] ** 2
    band = np.sort(p[:, lo:hi + 1], axis=1)[:, ::-1]
    if band.shape[1] <= k:
        return 0
    a, b = band[:, k - 1], band[:, k]
    return int(np.sum(np.abs(a - b) <= gap * np.maximum(a, 1e-300)))


def poll_batch(br, jid, out, tries=4000, sleep_s=0.005, getter=None):
    """The poll loop of WaveCyclesBatchFetcher.mq5:127-132: a running batch job must answer OK with
    ready == 0 (the Fetcher sleeps only on that answer; anything else ends its loop)."""
    import time
    getter = getter or br.gpu_try_get_cycles_batch
    for _ in range(tries):
        st, n, ready = getter(jid, out)
        if st == br.OK and ready == 1:
            return st, n, ready
        assert st == br.OK and ready == 0, (st, br.last_error())
        time.sleep(sleep_s)
    raise AssertionError("batch job did not finish within the Fetcher's poll budget")


def ocfg_from(oracle, cfg):
    o = oracle.PipelineCfg()
    C.memmove(C.byref(o), C.byref(cfg), C.sizeof(o))
    return o


# ---- A4: forward FFT behind gpu_fft_real_forward (imports.mqh:8) --------------------------------
@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 256, 512, 1024, 2048, 4096, 8192])
def test_fft_real_forward_matches_oracle(br, oracle, n):
    x = synth.random_walk(100 + n, n)
    out = br.gpu_fft_real_forward(x)
    ref = oracle.fft_interleaved(x)
    assert np.abs(out - ref).max() / np.abs(ref).max() < REL_TOL
    # detrended-like input (no large DC): per-bin accuracy is visible here
    y = np.random.default_rng(n).standard_normal(n)
    out = br.gpu_fft_real_forward(y)
    ref = oracle.fft_interleaved(y)
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-13


def test_fft_known_answers(br):
    n, k = 1024, 37
    t = np.arange(n)
    out = br.gpu_fft_real_forward(2.5 * np.cos(2 * np.pi * k * t / n + 0.3))
    p = out[0::2] ** 2 + out[1::2] ** 2
    assert int(np.argmax(p)) == k and p[k] == pytest.approx((2.5 * n / 2) ** 2, rel=1e-12)
    flat = br.gpu_fft_real_forward(np.full(n, 1.23456))
    assert np.all(flat[2:] == 0.0)                      # exact zeros off DC, like the CPU statement
    imp = np.zeros(n); imp[0] = 3.0
    assert np.allclose(br.gpu_fft_real_forward(imp)[0::2], 3.0)


def test_fft_bad_args(br):
    with pytest.raises(br.WaveSpecError) as e:
        br.gpu_fft_real_forward(np.zeros(1000))
    assert e.value.status == br.BAD_ARGS
    with pytest.raises(br.WaveSpecError):
        br.gpu_fft_real_forward(np.zeros(1))
    assert "power of two" in br.last_error()


def test_fft_batch_and_sliding(br, oracle):
    n, nwin = 256, 7
    x = synth.random_walk(5, n * nwin)
    out = br.gpu_fft_real_forward_batch(x, n, nwin)
    for w in range(nwin):
        ref = oracle.fft_interleaved(x[w * n:(w + 1) * n])
        assert np.abs(out[w] - ref).max() / np.abs(ref).max() < REL_TOL
    for hop in (1, 3, 64):
        sl = br.fft_real_forward_sliding(x[:900], n, hop)
        assert sl.shape[0] == 1 + (900 - n) // hop
        for w in (0, sl.shape[0] // 2, sl.shape[0] - 1):
            ref = oracle.fft_interleaved(x[w * hop:w * hop + n])
            assert np.abs(sl[w] - ref).max() / np.abs(ref).max() < REL_TOL


def test_fft_inverse_roundtrip(br):
    for n in (4, 64, 1024, 4096):
        x = np.random.default_rng(n).standard_normal(n)
        spec = br.gpu_fft_real_forward(x)
        back = br.gpu_fft_real_inverse(spec)
        nyq = np.sum(x * (-1.0) ** np.arange(n))          # the dropped Nyquist bin
        expect = x - nyq * (-1.0) ** np.arange(n) / n
        assert np.abs(back - expect).max() < 1e-12


# ---- the five BASELINE configs at oracle-sized lengths ------------------------------------------
def run_both(br, oracle, series, cfg, outputs):
    got = br.pipeline_host(series, cfg, outputs)
    ref = oracle.pipeline_series(series, ocfg_from(oracle, cfg), outputs)
    return got, ref


def check_planes(br, got, ref, cfg):
    n = cfg.window_len
    if "spectra" in ref:
        assert rel_err(got["spectra"], ref["spectra"]) < REL_TOL
    if "bins" in ref:
        assert np.array_equal(got["bins"], ref["bins"]), "selected cycle bins differ"
    if "waves" in ref:
        scale = np.abs(ref["waves"]).max()
        assert np.abs(got["waves"] - ref["waves"]).max() <= REL_TOL * scale
    if "rows" in ref:
        g, r = got["rows"], ref["rows"]
        for f in (0, 1, 2, 6, 14):                          # amplitude, freq, period, energy, method
            if f >= g.shape[-1]:                            # older row layouts: a prefix of the 15 fields
                continue
            assert np.abs(g[..., f] - r[..., f]).max() <= REL_TOL * max(1e-300, np.abs(r[..., f]).max())
        # Phase and eta are functions of atan2(im, re) of the selected bin, so the 1e-9 bar on the
        # spectrum (max |dX| / max |X| per window) propagates as |d phase| <= 1e-9 * max|X| / |X_k|:
        # that is the tolerance when the spectra plane is there to give max|X| (the in-band
        # bins of a raw price window sit 3-5 decades below its DC term); 1e-7 absolute otherwise.
        if g.shape[-1] > 3:
            dph = np.abs(np.angle(np.exp(1j * (g[..., 3] - r[..., 3]))))
            if "spectra" in ref:
                xmax = np.abs(ref["spectra"]).max(axis=-1)[..., None]
                xk = r[..., 0] * n / 2.0
                tol = REL_TOL * xmax / np.where(xk > 0, xk, np.inf)
                tol = np.where(xk > 0, np.maximum(tol, 4e-16), 0.0)
            else:
                tol = np.full(dph.shape, 1e-7)
            assert np.all(dph <= tol), float((dph - tol).max())
            REPORT.setdefault("max_row_phase_err", 0.0)
            REPORT["max_row_phase_err"] = max(REPORT["max_row_phase_err"], float(dph.max()))
            m = min(cfg.row_stride, 15)
            if m > 4:
                # eta_bars = d / (2 pi f), d in [0, pi): error = phase error * period / 2 pi, except
                # where d sits on its wrap (0 <-> pi), i.e. within the phase error of a half-period
                # jump: those rows are compared modulo half a period and COUNTED
                period = np.where(r[..., 2] > 0, r[..., 2], 1.0)
                etol = tol * period / (2 * np.pi) + 1e-15 * period
                d = np.abs(g[..., 4] - r[..., 4])
                wrapped = d > etol
                d = np.where(wrapped, np.abs(0.5 * period - d), d)
                assert np.all(d <= etol), float((d - etol).max())
                REPORT["eta_rows_on_the_half_period_wrap"] = REPORT.get("eta_rows_on_the_half_period_wrap", 0) + int(wrapped.sum())
                REPORT["eta_rows_compared"] = REPORT.get("eta_rows_compared", 0) + int(wrapped.size)
                if m > 5:
                    # eta_seconds = eta_bars * sample_rate_seconds, same rows, same rule
                    rate = cfg.sample_rate_seconds
                    ds = np.abs(g[..., 5] - r[..., 5])
                    ds = np.where(wrapped, np.abs(0.5 * period * rate - ds), ds)
                    assert np.all(ds <= etol * rate * (1 + 1e-12)), float((ds - etol * rate).max())
    if "kalman" in ref:
        assert np.array_equal(got["kalman"], ref["kalman"]), "Kalman4D must be bit-identical"
    if "wkalman" in ref:
        assert np.abs(got["wkalman"] - ref["wkalman"]).max() <= 1e-9 * max(1.0, np.abs(ref["wkalman"]).max())
    if "phase" in ref:
        gp, rp = got["phase"], ref["phase"]
        dph = np.angle(np.exp(1j * (gp[:, 0] - rp[:, 0])))
        assert np.abs(dph).max() < 1e-6


def check_unwrap_and_group_delay(name, gp, rp, tol=1e-6, utol=None, gtol=None):
    """Unwrapped phase and group delay (A6).  The unwrap adds +-2 pi where a phase step crosses +-pi:
    a window with a step within `tol` of that decision can legitimately take the other branch, so
    its planes are compared modulo the branch (unwrapped phase modulo 2 pi) and the window is COUNTED;
    every other window is compared directly."""
    ph = rp[:, 0]
    d = np.abs(np.abs(np.diff(ph, axis=1)) - np.pi)
    safe = d.min(axis=1) > tol
    REPORT["unwrap_windows_" + name] = REPORT.get("unwrap_windows_" + name, 0) + int(safe.size)
    REPORT["unwrap_windows_on_a_pi_decision_" + name] = REPORT.get("unwrap_windows_on_a_pi_decision_" + name, 0) + int((~safe).sum())
    assert safe.sum() > 0.9 * safe.size
    assert np.abs(gp[safe, 1] - rp[safe, 1]).max() < (tol if utol is None else utol)
    assert np.abs(gp[safe, 2] - rp[safe, 2]).max() < (tol if gtol is None else gtol)
    if (~safe).any():
        dd = np.angle(np.exp(1j * (gp[~safe, 1] - rp[~safe, 1])))
        assert np.abs(dd).max() < tol


def test_config1_mean_hann_top5(br, oracle):
    """C1: N=512, mean removal + Hann (gpu_wip form), top-5, band 9-200."""
    s = synth.random_walk(0, 4000)
    cfg = br.default_cfg(512, top_k=5, min_period=9.0, max_period=200.0, detrend=br.DETREND_MEAN,
                         window_type=br.WINDOW_HANN_WIP)
    out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES
    got, ref = run_both(br, oracle, s, cfg, out)
    check_planes(br, got, ref, cfg)
    assert count_near_ties("config1", ref["spectra"], 3, 56, 5) == 0


def test_config2_plain_top8(br, oracle):
    """C2: N=1024, no detrend / window, top-8, band 18-200 (nodetrend.mq5 path)."""
    s = synth.random_walk_batch(0, 3, 3500)
    cfg = br.default_cfg(1024, top_k=8, min_period=18.0, max_period=200.0)
    out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES
    got = br.pipeline_host(s, cfg, out)
    for i in range(3):
        ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), out)
        check_planes(br, {k: v[i] for k, v in got.items()}, ref, cfg)
        # per-bin check off DC: relative to the in-band magnitudes, not the DC term
        gb = got["spectra"][i][:, 12:114]; rb = ref["spectra"][:, 12:114]
        assert rel_err(gb, rb) < REL_TOL
        count_near_ties("config2", ref["spectra"], 6, 56, 8)


def test_config3_iir_blackman_kalman(br, oracle):
    """C3: N=2048, trend IIR T=1024, Blackman, Kalman4D, band 18-52."""
    s = synth.random_walk(2, 2048 + 700)
    cfg = br.default_cfg(2048, top_k=8, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR,
                         trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
    out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_KALMAN | br.OUT_WAVES
    got, ref = run_both(br, oracle, s, cfg, out)
    check_planes(br, got, ref, cfg)
    gb = got["spectra"][:, 80:228]; rb = ref["spectra"][:, 80:228]     # the band itself
    assert rel_err(gb, rb) < REL_TOL
    count_near_ties("config3", ref["spectra"], 40, 113, 8)


def test_config4_phase_and_top8_reconstruction(br, oracle):
    """C4: N=1024, phase chain + A8a/A8b outputs for the top-8 cycles; Hann + selection sort +
    weight-Kalman as in WaveSpecZZ_1.0.4-kalman.mq5."""
    s = synth.random_walk(3, 1024 + 500)
    cfg = br.default_cfg(1024, top_k=8, min_period=12.0, max_period=256.0, window_type=br.WINDOW_HANN,
                         select=br.SELECT_SORT)
    out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_WAVES | br.OUT_ROWS | br.OUT_WKALMAN | br.OUT_PHASE
    got, ref = run_both(br, oracle, s, cfg, out)
    check_planes(br, got, ref, cfg)
    # unwrapped phase / group delay only where no bin sits within 1e-6 of a +-pi jump decision
    count_near_ties("config4", ref["spectra"], 4, 85, 8)
    check_unwrap_and_group_delay("config4", got["phase"], ref["phase"])


def test_config5_n4096_batch_api(br, oracle):
    """C5: N=4096, K=4, band 9-200, stride 15 through gpu_submit_extract_cycles_batch."""
    s = synth.random_walk(4, 4096 + 300)
    st, jid = br.gpu_submit_extract_cycles_batch(s, 4096, 1, 4, 9.0, 200.0, 60.0, 1, 10, 15)
    assert st == br.OK and jid != 0, br.last_error()
    nwin = 301
    out = np.zeros(nwin * 4 * 15)
    st, n, ready = poll_batch(br, jid, out)
    assert ready == 1 and n == nwin * 4
    assert br.gpu_free_job(jid) == br.OK
    cfg = oracle.default_cfg(4096, top_k=4, min_period=9.0, max_period=200.0, sample_rate_seconds=60.0)
    ref = oracle.pipeline_series(s, cfg, oracle.OUT_ROWS | oracle.OUT_BINS)
    rows = out.reshape(nwin, 4, 15)
    # period = N / bin is exact: selected bins are bit-exact through the public API
    assert np.array_equal(np.rint(4096 / rows[..., 2]).astype(int), ref["bins"])
    assert np.abs(rows[..., 0] - ref["rows"][..., 0]).max() <= REL_TOL * ref["rows"][..., 0].max()
    assert np.all(rows[..., 14] == 0.0)


# ---- shared-butterfly sliding kernel (plain hop-1 path) ------------------------------------------
@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_sliding_shared_kernel_all_lengths_and_ragged_tiles(br, oracle, n):
    # series lengths chosen so the last tile is partial and the first tile is the only full one
    tiles = {256: 128, 512: 64, 1024: 32, 2048: 16, 4096: 8}[n]
    for extra in (0, 1, tiles - 1, tiles, 2 * tiles + 3):
        s = synth.random_walk_batch(200 + n, 2, n + extra)
        cfg = br.default_cfg(n, top_k=8, min_period=9.0, max_period=200.0)
        out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES
        got = br.pipeline_host(s, cfg, out)
        if extra > 0:                                # 2 series x 1 window: the per-window kernel takes it
            assert br.last_kernel() in ("sliding_shared", "sliding_overlap", "sliding_staged")
        for i in range(2):
            ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), out)
            check_planes(br, {k: v[i] for k, v in got.items()}, ref, cfg)
            # off-DC bins: error relative to the largest non-DC bin of the window
            assert rel_err(got["spectra"][i][:, 2:], ref["spectra"][:, 2:]) < 1e-12


def test_sliding_shared_rows_only_and_sort_rule(br, oracle):
    s = synth.random_walk(210, 1024 + 777)
    for sel in (br.SELECT_INSERTION, br.SELECT_SORT):
        cfg = br.default_cfg(1024, top_k=8, min_period=12.0, max_period=256.0, select=sel)
        out = br.OUT_BINS | br.OUT_ROWS | br.OUT_WKALMAN
        got, ref = run_both(br, oracle, s, cfg, out)
        assert br.last_kernel() in ("sliding_shared", "sliding_overlap", "sliding_staged")
        check_planes(br, got, ref, cfg)


def test_sliding_shared_long_series_spot_checks(br, oracle):
    """200k bars through the batch API; windows spot-checked against the oracle, every selected
    bin compared on a strided subset."""
    n = 1024
    s = synth.random_walk(220, 200000)
    cfg = br.default_cfg(n, top_k=8, outputs=br.OUT_BINS | br.OUT_SPECTRA)
    got = br.pipeline_host(s, cfg)
    assert br.last_kernel() in ("sliding_shared", "sliding_overlap", "sliding_staged")
    nw = got["bins"].shape[0]
    ocfg = ocfg_from(oracle, cfg)
    for w0 in (0, 31, 32, 12345, nw - 40):
        seg = s[w0:w0 + n + 39]
        ref = oracle.pipeline_series(seg, ocfg, oracle.OUT_BINS | oracle.OUT_SPECTRA)
        assert np.array_equal(got["bins"][w0:w0 + 40], ref["bins"])
        assert rel_err(got["spectra"][w0:w0 + 40], ref["spectra"]) < REL_TOL
    # linearity property at full length: spectrum(a*x) == a*spectrum(x) exactly for a power of two
    got2 = br.pipeline_host(4.0 * s, cfg, br.OUT_SPECTRA)
    assert np.array_equal(got2["spectra"], 4.0 * got["spectra"])


# ---- prologue variants (A2, A3) -----------------------------------------------------------------
@pytest.mark.parametrize("wtype", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("detrend", [0, 1, 2])
def test_prologue_matrix(br, oracle, wtype, detrend):
    s = synth.random_walk(30 + wtype, 256 + 200)
    cfg = br.default_cfg(256, top_k=6, min_period=4.0, max_period=100.0, detrend=detrend,
                         trend_period=64.0, window_type=wtype)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS)
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("hop", [1, 5])
def test_window_lengths_and_hops(br, oracle, n, hop):
    s = synth.random_walk(40 + n, n + 333)
    cfg = br.default_cfg(n, hop=hop, top_k=4, min_period=9.0, max_period=200.0)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_WAVES)
    check_planes(br, got, ref, cfg)


# ---- edge cases ---------------------------------------------------------------------------------
def test_flat_market_ties_resolve_by_bin_order(br, oracle):
    s = np.full(1500, 1.2345)
    cfg = br.default_cfg(1024, top_k=8)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_SPECTRA)
    assert np.array_equal(got["bins"], ref["bins"])
    assert np.all(got["bins"] == np.arange(6, 14))


def test_quantised_prices_with_long_flat_runs(br, oracle):
    rng = np.random.default_rng(9)
    s = np.round(1.1 + 1e-5 * np.cumsum(rng.integers(-1, 2, 3000) * (rng.random(3000) < 0.05)), 5)
    cfg = br.default_cfg(512, top_k=8, min_period=9.0, max_period=200.0)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_SPECTRA)
    check_planes(br, got, ref, cfg)


def test_single_window_and_short_series(br, oracle):
    s = synth.random_walk(50, 1024)
    cfg = br.default_cfg(1024)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_SPECTRA)
    assert got["bins"].shape == (1, 8)
    check_planes(br, got, ref, cfg)
    with pytest.raises(br.WaveSpecError) as e:
        br.pipeline_host(s[:1000], cfg)
    assert e.value.status == br.BAD_ARGS


def test_top_k_larger_than_band(br, oracle):
    s = synth.random_walk(51, 700)
    cfg = br.default_cfg(256, top_k=12, min_period=40.0, max_period=64.0)     # band = bins 4..6
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES)
    assert np.array_equal(got["bins"], ref["bins"])
    assert np.all(got["bins"][:, 3:] == -1) and np.all(got["rows"][:, 3:, :] == 0.0)
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("stride", [1, 4, 8, 12, 15, 20])
def test_row_strides(br, oracle, stride):
    s = synth.random_walk(52, 900)
    cfg = br.default_cfg(512, top_k=4, row_stride=stride, min_period=9.0, max_period=200.0)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_ROWS)
    assert got["rows"].shape[-1] == stride
    m = min(stride, 15)
    assert np.abs(got["rows"][..., 0] - ref["rows"][..., 0]).max() <= REL_TOL * ref["rows"][..., 0].max()
    if stride > 15:
        assert np.all(got["rows"][..., 15:] == 0.0)
    if m >= 3:
        assert np.array_equal(got["rows"][..., 2], ref["rows"][..., 2])


def test_selection_sort_rule_ties_and_k_ge_1(br, oracle):
    s = np.full(900, 1.5)                                   # all in-band powers equal (zero)
    cfg = br.default_cfg(256, top_k=6, min_period=2.0, max_period=1e9, select=br.SELECT_SORT)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS)
    assert np.array_equal(got["bins"], ref["bins"])
    assert got["bins"].min() >= 1


# ---- A9 / A10 sequential recursions -------------------------------------------------------------
def test_kalman4d_bit_exact_long(br, oracle):
    s = synth.random_walk_batch(60, 5, 64 + 20000)
    cfg = br.default_cfg(64, outputs=br.OUT_KALMAN)
    got = br.pipeline_host(s, cfg, br.OUT_KALMAN)
    for i in range(5):
        ref = oracle.kalman4d_series(s[i][63:])
        assert np.array_equal(got["kalman"][i], ref)


def test_kalman4d_parameter_variants(br, oracle):
    s = synth.random_walk(61, 64 + 3000)
    for over in ({"ema_blend_period": 5.0}, {"adapt_gain": 0.0}, {"clip_std": 0.0},
                 {"clip_std": 0.5, "meas_noise": 1e-6}, {"follow_strength": 0.01}):
        cfg = br.default_cfg(64)
        for k, v in over.items():
            setattr(cfg.kalman, k, v)
        got = br.pipeline_host(s, cfg, br.OUT_KALMAN)
        ocfg = ocfg_from(oracle, cfg)
        ref = oracle.kalman4d_series(s[63:], ocfg.kalman)
        assert np.array_equal(got["kalman"], ref), over


# ---- A11 PLA ------------------------------------------------------------------------------------
def test_pla_pivots_bit_exact(br, oracle):
    n = 512
    s = synth.random_walk(70, n + 150)
    lines, bounds, counts = br.pla_windows_host(s, n, 1, 32, 0.0005)
    for w in range(0, 151, 7):
        line, st, en, _, _ = oracle.pla_build(s[w:w + n], 32, 0.0005)
        assert counts[w] == st.size
        assert np.array_equal(bounds[w, :st.size, 0], st) and np.array_equal(bounds[w, :st.size, 1], en)
        assert np.array_equal(lines[w], line), "PLA line must be bit-identical (same sums, same order)"


def test_pla_feed_pipeline(br, oracle):
    n = 256
    s = synth.random_walk(71, n + 120)
    cfg = br.default_cfg(n, top_k=4, min_period=9.0, max_period=100.0, feed=br.FEED_PLA,
                         detrend=br.DETREND_IIR, trend_period=64.0, window_type=br.WINDOW_BLACKMAN,
                         pla_max_segments=16, pla_max_error=0.0003)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_KALMAN)
    check_planes(br, got, ref, cfg)


# ---- imports.mqh job API ------------------------------------------------------------------------
def test_extract_cycles_sync_and_async(br, oracle):
    s = synth.random_walk(80, 1024)
    rows = br.gpu_extract_cycles(s, 4, 9.0, 200.0, 60.0, method=1, out_stride=15)
    cfg = oracle.default_cfg(1024, top_k=4, min_period=9.0, max_period=200.0)
    ref = oracle.pipeline_series(s, cfg, oracle.OUT_ROWS | oracle.OUT_BINS)
    assert rows.shape == (4, 15)
    assert np.array_equal(np.rint(1024 / rows[:, 2]).astype(int), ref["bins"][0])
    assert np.all(rows[:, 14] == 0.0)                       # FFT ridge rows are tagged method 0

    jobs = []
    for i in range(16):                                     # out-of-order polling of many jobs
        st, jid = br.gpu_submit_extract_cycles(synth.random_walk(81 + i, 1024), 4, 9.0, 200.0, 60.0, 0, 10)
        assert st == br.OK and jid != 0
        jobs.append(jid)
    buf = np.zeros((4, 15))
    for i in reversed(range(16)):
        for _ in range(100000):
            st, n, ready = br.gpu_try_get_cycles(jobs[i], buf, 15, 4)
            if st != br.NOT_READY:
                break
        assert st == br.OK and ready == 1 and n == 4
        r = oracle.pipeline_series(synth.random_walk(81 + i, 1024), cfg, oracle.OUT_BINS)
        assert np.array_equal(np.rint(1024 / buf[:, 2]).astype(int), r["bins"][0])
        assert br.gpu_free_job(jobs[i]) == br.OK
    assert br.gpu_free_job(jobs[0]) == br.BAD_ARGS          # already freed
    st, n, ready = br.gpu_try_get_cycles(123456789, buf, 15, 4)
    assert st == br.BAD_ARGS and ready == 0


def test_free_in_flight_job_and_reinit(br):
    s = synth.random_walk(90, 200000)
    st, jid = br.gpu_submit_extract_cycles_batch(s, 1024, 1, 8, 18.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK
    assert br.gpu_free_job(jid) == br.OK                    # tolerated while still running
    assert br.gpu_init(0, 64) == br.OK                      # idempotent
    st, jid = br.gpu_submit_extract_cycles_batch(s[:100], 1024, 1, 8, 18.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.BAD_ARGS and jid == 0
    st, jid = br.gpu_submit_extract_cycles_batch(s, 8196, 1, 8, 18.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.BAD_ARGS                                # non power of two (1.0.4-old default)


def test_last_error_utf16_contract(br):
    buf = (C.c_uint16 * 8)()
    with pytest.raises(br.WaveSpecError):
        br.gpu_fft_real_forward(np.zeros(3))
    n = br.lib().gpu_get_last_error_w(buf, 8)
    assert n == 8 and buf[7] == 0                           # truncated, terminator counted
    assert br.lib().gpu_get_last_error_w(buf, 0) == 0


def test_split_rows_kernel_path_matches_oracle():
    """WAVESPEC_SPLIT=1: sliding kernel hands the in-band bins to the separate rows kernel
    (ws_rows.cu).  Run in a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import ctypes as C, numpy as np, sys
sys.path.insert(0, ".")
from fft_wavespec_b200 import bridge as br, synth
from oracle import oracle as orc
assert br.gpu_init(0, 2) == 0
for n, extra in ((1024, 777), (512, 64), (4096, 9)):
    s = synth.random_walk_batch(500 + n, 2, n + extra)
    cfg = br.default_cfg(n, top_k=8, min_period=9.0, max_period=200.0)
    out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES
    got = br.pipeline_host(s, cfg, out)
    o = orc.PipelineCfg(); C.memmove(C.byref(o), C.byref(cfg), C.sizeof(o))
    for i in range(2):
        ref = orc.pipeline_series(s[i], o, out)
        assert np.array_equal(got["bins"][i], ref["bins"])
        assert np.abs(got["rows"][i][..., 0] - ref["rows"][..., 0]).max() <= 1e-9 * ref["rows"][..., 0].max()
        assert np.abs(got["waves"][i] - ref["waves"]).max() <= 1e-9 * np.abs(ref["waves"]).max()
        e = np.abs(got["spectra"][i] - ref["spectra"]).max(axis=1) / np.abs(ref["spectra"]).max(axis=1)
        assert e.max() < 1e-9
    rows_only = br.pipeline_host(s, cfg, br.OUT_BINS)
    assert np.array_equal(rows_only["bins"], got["bins"])
print("split ok", br.launch_count())
'''
    env = dict(os.environ, WAVESPEC_SPLIT="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "split ok" in r.stdout


# ---- A12 ZigZag pivot -> feed expansion ---------------------------------------------------------
def _zigzag_buffers(n, seed, density=0.06):
    """Synthetic stock-ZigZag buffers: sparse alternating highs/lows in `main`, high/low maps."""
    rng = np.random.default_rng(seed)
    close = synth.random_walk(seed, n)
    hi, lo = synth.high_low(seed, close)
    main = np.zeros(n); zh = np.zeros(n); zl = np.zeros(n)
    idx = np.flatnonzero(rng.random(n) < density)
    for c, a in enumerate(idx):
        if c % 2 == 0:
            main[a] = hi[a]; zh[a] = hi[a]
        else:
            main[a] = lo[a]; zl[a] = lo[a]
    return main, zh, zl, hi, lo


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_zigzag_feed_110_bit_exact(br, oracle, mode):
    n, N = 900, 256
    main, zh, zl, hi, lo = _zigzag_buffers(n, 700 + mode)
    main[:40] = 0.0                                        # first windows start before any pivot
    fb = (hi[0] + lo[0]) * 0.5
    lines, valid = br.zigzag_feed_host(main, hi, lo, N, 1, pivot_rule=0, mode=mode, fallback=fb)
    for w in range(0, n - N + 1, 13):
        ref = oracle.zigzag_feed_110(main[w:w + N], hi[w:w + N], lo[w:w + N], mode, hi[0], lo[0])
        assert np.array_equal(lines[w], ref), (mode, w)
    # a window with no pivot at all falls back to (high[0]+low[0])/2
    empty = np.zeros(n)
    lines, _ = br.zigzag_feed_host(empty, hi, lo, N, 7, pivot_rule=0, mode=0, fallback=fb)
    assert np.all(lines == fb)


@pytest.mark.parametrize("mode", [0, 1])
def test_zigzag_series_legacy_bit_exact(br, oracle, mode):
    n, N = 700, 128
    main, zh, zl, hi, lo = _zigzag_buffers(n, 710 + mode, density=0.03)
    main[100:140] = 0.0; zh[120] = hi[120]                 # pivot visible only through the high map
    main[300] = np.inf                                     # non-finite main value is ignored
    lines, valid = br.zigzag_feed_host(main, zh, zl, N, 1, pivot_rule=1, mode=1 - mode, fallback=0.0,
                                       min_pivots=2)
    for w in range(0, n - N + 1, 9):
        ok, ref = oracle.zigzag_series_legacy(main[w:w + N], zh[w:w + N], zl[w:w + N], mode)
        assert bool(valid[w]) == ok, w
        if ok:
            assert np.array_equal(lines[w], ref), (mode, w)


# ---- A13 tracker pool + stable slots ------------------------------------------------------------
@pytest.mark.parametrize("case", ["c3", "plain", "plain_rows", "wide"])
def test_tracker_pool_slots_match_oracle(br, oracle, case):
    if case == "c3":
        n, over = 2048, dict(min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR, trend_period=1024.0,
                             window_type=br.WINDOW_BLACKMAN)
        out = br.OUT_TRACKER | br.OUT_BINS
    elif case == "plain":
        n, over = 1024, dict(min_period=18.0, max_period=200.0)
        out = br.OUT_TRACKER
    elif case == "plain_rows":
        n, over = 1024, dict(min_period=18.0, max_period=200.0)
        out = br.OUT_TRACKER | br.OUT_ROWS | br.OUT_BINS | br.OUT_SPECTRA
    else:
        n, over = 256, dict(min_period=4.0, max_period=64.0, window_type=br.WINDOW_HANN, tracker_tolerance=1.0,
                            tracker_max_inactive=2)
        out = br.OUT_TRACKER
    s = synth.random_walk_batch(900, 2, n + 260)
    cfg = br.default_cfg(n, **over)
    got = br.pipeline_host(s, cfg, out)
    for i in range(2):
        ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), out)
        assert np.array_equal(got["trk_index"][i], ref["trk_index"]), case
        assert np.array_equal(got["trk_period"][i], ref["trk_period"]), case
        if "bins" in ref:
            assert np.array_equal(got["bins"][i], ref["bins"])
        if "rows" in ref:
            assert np.abs(got["rows"][i][..., 0] - ref["rows"][..., 0]).max() <= REL_TOL * ref["rows"][..., 0].max()


# ---- 8f rank 2: device-side decode of the batch result into the cycle-cache record --------------
@pytest.mark.parametrize("hop,top_k,music_only,weights", [(1, 4, False, False), (1, 2, False, True),
                                                           (3, 4, False, True), (1, 1, False, False),
                                                           (1, 4, True, True), (7, 3, True, False)])
def test_cycle_cache_record_matches_sequential_decode(br, oracle, hop, top_k, music_only, weights):
    N, bars = 128, 700
    rng = np.random.default_rng(hop * 10 + top_k)
    nwin = 1 + (bars - N) // hop
    rows = np.zeros((nwin * top_k, 15))
    rows[:, 0] = rng.random(rows.shape[0]) * 1e-3            # amplitude
    rows[:, 1] = rng.random(rows.shape[0]) * 0.2 + 0.005     # freq
    rows[:, 2] = 1.0 / rows[:, 1]
    rows[:, 3] = rng.uniform(-np.pi, np.pi, rows.shape[0])
    rows[:, 5] = rng.random(rows.shape[0]) * 3000
    rows[:, 6:14] = rng.random((rows.shape[0], 8))
    rows[:, 8] = rng.uniform(-60, 30, rows.shape[0])         # snr_db
    rows[:, 14] = (rng.random(rows.shape[0]) < 0.7).astype(float)   # mixed methods: exercises the skip search
    kw = dict(period_seconds=60.0, music_only=music_only, use_music_weights=weights)
    got = br.cycle_cache_host(rows, top_k, N, hop, bars, **kw)
    ref = oracle.cycle_cache(rows, top_k, N, hop, bars, **kw)
    empty = ref == np.finfo(float).max
    assert np.array_equal(got == np.finfo(float).max, empty)
    den = np.maximum(1e-300, np.abs(ref[~empty]))
    assert (np.abs(got[~empty] - ref[~empty]) / den).max() < 1e-9 or np.abs(got[~empty] - ref[~empty]).max() < 1e-15


def test_cycle_cache_from_library_rows(br, oracle):
    """End to end: batch API rows -> cache record; FFT-ridge rows carry method 0, so with
    InpMusicOnly the record is all EMPTY_VALUE, and with it off buffer 1 holds amp*sin(phase)."""
    s = synth.random_walk(950, 1500)
    cfg = br.default_cfg(256, top_k=4, min_period=9.0, max_period=100.0)
    rows = br.pipeline_host(s, cfg, br.OUT_ROWS)["rows"]
    rec = br.cycle_cache_host(rows, 4, 256, 1, 1500, music_only=True)
    assert np.all(rec == np.finfo(float).max)
    # the MUSIC quality fields of FFT-ridge rows are 0, and the decode zero-weights rows below
    # InpMinCoherence / InpMinScore (:1079) even with InpUseMusicWeights off: both go to 0
    kw = dict(music_only=False, use_music_weights=False, min_coherence=0.0, min_score=0.0)
    zero = br.cycle_cache_host(rows, 4, 256, 1, 1500, music_only=False, use_music_weights=False)
    assert np.all(zero[:, 0] == 0.0)
    rec = br.cycle_cache_host(rows, 4, 256, 1, 1500, **kw)
    ref = oracle.cycle_cache(rows, 4, 256, 1, 1500, **kw)
    nw = rows.shape[0]
    assert np.abs(rec[:nw, 0] - rows[:, 0, 0] * np.sin(rows[:, 0, 3])).max() < 1e-15   # k = 0 for every bar < nwin
    assert np.abs(rec - ref).max() / 1.0 < 1e-9 or np.array_equal(rec, ref)


# ---- BASELINE full sizes through size-independent properties -------------------------------------
def test_full_size_series_1m_bars_properties(br, oracle):
    """One config-2 series at its full length (1M bars, N=1024, 998 977 windows) through the batch
    API: row count, exact selected bins on a strided sample of windows against the oracle, exact
    linearity of the spectra plane under a power-of-two scale, monotone period ordering rule."""
    n, bars = 1024, 1000000
    s = synth.random_walk(7, bars)
    st, jid = br.gpu_submit_extract_cycles_batch(s, n, 1, 8, 18.0, 200.0, 60.0, 0, 10, 15)
    assert st == br.OK
    nwin = bars - n + 1
    out = np.empty(nwin * 8 * 15)
    st, cnt, ready = poll_batch(br, jid, out, sleep_s=0.001)
    assert st == br.OK and ready == 1 and cnt == nwin * 8
    assert br.gpu_free_job(jid) == br.OK
    rows = out.reshape(nwin, 8, 15)
    bins = np.rint(n / rows[..., 2]).astype(np.int32)
    assert bins.min() >= 6 and bins.max() <= 56
    amp = rows[..., 0]
    assert np.all(np.diff(amp, axis=1) <= 0)                 # strongest first in every window
    ocfg = oracle.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
    for w in range(0, nwin, 49999):
        ref = oracle.pipeline_series(s[w:w + n], ocfg, oracle.OUT_BINS | oracle.OUT_ROWS)
        assert np.array_equal(bins[w], ref["bins"][0]), w
        assert np.abs(rows[w, :, 0] - ref["rows"][0, :, 0]).max() <= REL_TOL * ref["rows"][0, :, 0].max()
    # checksum-of-checksums: per-window energy ratio of the top-8 never exceeds 1 and is positive
    er = rows[..., 6].sum(axis=1)
    assert er.max() <= 1.0 + 1e-12 and er.min() > 0.0


def test_pla_long_windows_direct_render_fallback(br, oracle):
    """N = 1024 random-walk windows produce more PLA segments than the shared staging holds
    (left spines are not bounded by max_segments): those windows render by themselves."""
    n = 1024
    s = synth.random_walk(71, n + 40)
    lines, bounds, counts = br.pla_windows_host(s, n, 1, 32, 0.0005)
    seen_big = False
    for w in range(0, 41, 5):
        line, st_, en_, _, _ = oracle.pla_build(s[w:w + n], 32, 0.0005)
        assert counts[w] == st_.size
        seen_big |= st_.size > 60
        assert np.array_equal(lines[w], line)
        m = min(st_.size, bounds.shape[1])
        assert np.array_equal(bounds[w, :m, 0], st_[:m]) and np.array_equal(bounds[w, :m, 1], en_[:m])


# ---- warp-per-window FFT kernel (ws_window_fft_warp.cu): N = 512 / 1024 / 2048 / 4096 ------------
@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("detrend,wtype", [(0, 3), (1, 3), (2, 1), (2, 0), (1, 0), (0, 5), (0, 2), (1, 2), (0, 4), (2, 4),
                                           (1, 4), (2, 2)])
def test_warp_kernel_prologues_match_oracle(br, oracle, n, detrend, wtype):
    """Every detrend / window combination through the one-warp-per-window kernel, on a window count
    that leaves a ragged last tile and warps without a window in it."""
    s = synth.random_walk(900 + n + 7 * detrend + wtype, n + 64 + 64 + 5)
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=52.0, detrend=detrend,
                         trend_period=float(n // 2), window_type=wtype)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES)
    assert br.last_kernel() == "window_fft_warp"
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("n,hop", [(512, 3), (1024, 7), (2048, 2), (512, 512), (1024, 1024), (1024, 300)])
def test_warp_kernel_hops_and_contiguous_batches(br, oracle, n, hop):
    nwin = 21
    s = synth.random_walk(950 + n + hop, n + (nwin - 1) * hop)
    cfg = br.default_cfg(n, hop=hop, top_k=4, min_period=9.0, max_period=200.0, window_type=br.WINDOW_HANN)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_WAVES)
    assert got["bins"].shape[0] == nwin
    assert br.last_kernel() in ("window_fft_warp", "window_fft")   # wide strides may fall back
    if hop <= 7:
        assert br.last_kernel() == "window_fft_warp"
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("select", [0, 1])
@pytest.mark.parametrize("band", [(2.0, 1.0e9), (18.0, 52.0), (4.0, 40.0), (3.0, 3.5)])
def test_warp_kernel_selection_rules_and_band_widths(br, oracle, select, band):
    """Wide bands take the shared-memory scan, narrow ones the register-resident rounds; the last
    band is narrower than top_k; both tie rules."""
    n = 1024
    s = np.round(synth.random_walk(990, n + 150), 3)          # coarse quantisation: power ties
    cfg = br.default_cfg(n, top_k=8, min_period=band[0], max_period=band[1], select=select,
                         detrend=br.DETREND_MEAN, window_type=br.WINDOW_HANN)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES | br.OUT_SPECTRA)
    assert br.last_kernel() == "window_fft_warp"
    check_planes(br, got, ref, cfg)


def test_warp_kernel_many_series_rows_only(br, oracle):
    """No spectra plane: the split only feeds the selection (and nothing is stored per bin)."""
    n = 512
    s = synth.random_walk_batch(1000, 5, n + 700)
    cfg = br.default_cfg(n, top_k=5, min_period=9.0, max_period=200.0, detrend=br.DETREND_MEAN,
                         window_type=br.WINDOW_HANN_WIP)
    got = br.pipeline_host(s, cfg, br.OUT_ROWS | br.OUT_BINS)
    assert br.last_kernel() == "window_fft_warp"
    for i in range(s.shape[0]):
        ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), br.OUT_ROWS | br.OUT_BINS)
        check_planes(br, {k: (v[i] if v is not None else None) for k, v in got.items()}, ref, cfg)


@pytest.mark.parametrize("top_k", [3, 4, 8, 12])
@pytest.mark.parametrize("n", [1024, 4096])
def test_sliding_kernel_wide_band_selection_with_ties(br, oracle, n, top_k):
    """Bands wider than the lanes' register candidates (config 5: 435 bins): top_k <= 8 takes the one-pass
    per-lane lists (ws_epilogue.cuh::warp_wide_topk), larger top_k the K scans; coarse quantisation and a
    flat stretch force equal powers, which must resolve to the lower bin."""
    s = np.round(synth.random_walk(1234 + n, n + 90), 3)
    s[n // 2:n // 2 + 400] = s[n // 2]
    # the widest bands whose capture still fits the sliding kernels' shared memory (bins 1..292 / 1..455)
    cfg = br.default_cfg(n, top_k=top_k, min_period=3.5 if n == 1024 else 9.0, max_period=1.0e9)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES)
    if top_k <= 8:                      # larger top_k may not fit the sliding kernels' shared memory at this band
        assert br.last_kernel().startswith("sliding")
    check_planes(br, got, ref, cfg)
    flat = np.full(n + 40, 1.2345)
    got, ref = run_both(br, oracle, flat, cfg, br.OUT_BINS)
    assert np.array_equal(got["bins"], ref["bins"])


# ---- producer / consumer sliding kernel vs the phase-ordered one --------------------------------
@pytest.mark.parametrize("n,tile", [(1024, 40), (512, 64), (256, 128)])
def test_overlap_kernel_ragged_tiles_against_oracle(br, oracle, n, tile):
    """The producer / consumer kernel (spectra + rows, band <= 64 bins) on window counts around its tile
    length: one window, one short of a tile, exactly one, one more, and two tiles plus a ragged third."""
    for nwin in (1, tile - 1, tile, tile + 1, 2 * tile + 7):
        s = synth.random_walk(1300 + n + nwin, n + nwin - 1)
        cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
        out = br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES
        got, ref = run_both(br, oracle, s, cfg, out)
        if nwin > 2:                                 # one or two windows go to the per-window kernel
            assert br.last_kernel() in ("sliding_overlap", "sliding_staged")
        check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("n", [512, 1024])
def test_overlap_kernel_rows_bit_identical_to_phase_kernel_and_repeatable(br, n):
    """Spectra + rows run the producer / consumer kernel (selection beside the chains, named
    barriers); rows alone run the kernel whose epilogue follows a block barrier.  Same arithmetic,
    different schedule: rows and bins must agree bit for bit over ~10^5 tiles' worth of hand-offs,
    and a second run must reproduce the first (a lost hand-off would show up as a mismatch)."""
    s = synth.random_walk_batch(1200 + n, 4, 120_000 + n)
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
    both = br.pipeline_host(s, cfg, br.OUT_SPECTRA | br.OUT_ROWS | br.OUT_BINS)
    assert br.last_kernel() in ("sliding_overlap", "sliding_staged")
    rows_only = br.pipeline_host(s, cfg, br.OUT_ROWS | br.OUT_BINS)
    assert br.last_kernel() == "sliding_shared"
    assert np.array_equal(both["bins"], rows_only["bins"])
    assert np.array_equal(both["rows"], rows_only["rows"])
    again = br.pipeline_host(s, cfg, br.OUT_SPECTRA | br.OUT_ROWS | br.OUT_BINS)
    assert np.array_equal(both["bins"], again["bins"])
    assert np.array_equal(both["rows"], again["rows"])
    assert np.array_equal(both["spectra"], again["spectra"])
    spec_only = br.pipeline_host(s, cfg, br.OUT_SPECTRA)
    assert np.array_equal(both["spectra"], spec_only["spectra"])


# ---- A1: applied price ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5, 6, 7])
def test_applied_price_bit_exact(br, oracle, mode):
    c = synth.random_walk(1300, 200_003)
    h, l = synth.high_low(1300, c)
    o = np.roll(c, 1); o[0] = c[0]
    got = br.applied_price_host(o, h, l, c, mode)
    assert br.last_kernel() == "applied_price"
    assert np.array_equal(got, oracle.applied_price(o, h, l, c, mode))
    # inputs a mode does not read may be omitted
    need = {1: (0, 0, 0, 1), 2: (1, 0, 0, 0), 3: (0, 1, 0, 0), 4: (0, 0, 1, 0), 5: (0, 1, 1, 0),
            6: (0, 1, 1, 1), 7: (0, 1, 1, 1)}[mode]
    args = [a if k else None for a, k in zip((o, h, l, c), need)]
    assert np.array_equal(br.applied_price_host(*args, mode), got)


def test_applied_price_bad_args(br):
    c = synth.random_walk(1301, 100)
    for mode, args in ((0, (c, c, c, c)), (8, (c, c, c, c)), (6, (None, c, c, None)), (5, (None, None, c, None))):
        with pytest.raises(br.WaveSpecError) as e:
            br.applied_price_host(*args, mode)
        assert e.value.status == br.BAD_ARGS


def test_median_price_feeds_the_pipeline(br, oracle):
    """The applied-price series is just another per-bar series for the window kernels."""
    c = synth.random_walk(1302, 512 + 400)
    h, l = synth.high_low(1302, c)
    med = br.applied_price_host(None, h, l, None, br.PRICE_MEDIAN)
    cfg = br.default_cfg(512, top_k=5, min_period=9.0, max_period=200.0, detrend=br.DETREND_MEAN,
                         window_type=br.WINDOW_HANN_WIP)
    got, ref = run_both(br, oracle, med, cfg, br.OUT_SPECTRA | br.OUT_BINS)
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("n", [256, 1024, 4096])
def test_sliding_kernels_odd_series_stride(br, oracle, n):
    """Series of odd length packed back to back: every second series starts on an odd sample, so its
    tiles take the shifted form of the 16-byte aligned bulk staging (and the split-at-the-boundary form
    of the 128-bit loads), and the last tile of each series runs into the next one's samples."""
    s = synth.random_walk_batch(4321 + n, 3, n + 151)
    cfg = br.default_cfg(n, top_k=8 if n < 4096 else 4, min_period=18.0, max_period=200.0)
    for outs in (br.OUT_SPECTRA | br.OUT_ROWS | br.OUT_BINS, br.OUT_ROWS | br.OUT_BINS, br.OUT_SPECTRA):
        got = br.pipeline_host(s, cfg, outs)
        assert br.last_kernel().startswith("sliding")
        for i in range(s.shape[0]):
            ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), outs)
            check_planes(br, {k: (v[i] if v is not None else None) for k, v in got.items()}, ref, cfg)


# ---- 8(e): one series split over GPUs by bar range ----------------------------------------------
@pytest.mark.parametrize("n,plain", [(1024, True), (512, False)])
def test_bar_range_split_reproduces_the_unsplit_series(br, n, plain):
    """A rank that owns windows [w0, w0 + cnt) of a series gets the samples of those windows plus the
    N - 1 halo; the stateless outputs of the pieces concatenate to the unsplit result.  Bitwise for
    the plain sliding path (a window's butterfly tree does not depend on its tile); the tile-level
    trend scan of the detrended path differs by rounding only."""
    from fft_wavespec_b200 import shard
    x = synth.random_walk(1400 + n, 50_000 + n)
    if plain:
        cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
    else:
        cfg = br.default_cfg(n, top_k=5, min_period=9.0, max_period=200.0, detrend=br.DETREND_IIR,
                             trend_period=256.0, window_type=br.WINDOW_HANN)
    outs = br.OUT_SPECTRA | br.OUT_ROWS | br.OUT_BINS
    whole = br.pipeline_host(x, cfg, outs)
    world = 3
    parts = []
    for r in range(world):
        w0, cnt, a0, na = shard.bar_range_shard(x.size, n, 1, r, world)
        parts.append(br.pipeline_host(x[a0:a0 + na], cfg, outs))
        assert parts[-1]["bins"].shape[0] == cnt
    for key in ("spectra", "rows", "bins"):
        cat = np.concatenate([p[key] for p in parts])
        if plain or key == "bins":
            assert np.array_equal(cat, whole[key]), key
        else:
            scale = np.abs(whole[key]).max()
            assert np.abs(cat - whole[key]).max() <= 1e-9 * scale, key


# ---- config 3 and config 1 at full length --------------------------------------------------------
def test_config3_full_size_1m_bars_properties(br, oracle):
    """One config-3 series at its full length (1M bars, N=2048, IIR T=1024 + Blackman, band 18-52):
    the warp-per-window kernel over 997 953 windows.  Selected bins exact and waves within 1e-9 on a
    strided sample of windows against the oracle (each window's trend filter restarts, so a window
    is checked from its own 2048 samples); Kalman4D bit-identical over the whole series; the
    selection is invariant under a power-of-two scale of the prices."""
    n, bars = 2048, 1_000_000
    s = synth.random_walk(8, bars)
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR,
                         trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
    got = br.pipeline_host(s, cfg, br.OUT_BINS | br.OUT_WAVES | br.OUT_KALMAN)
    assert br.last_kernel() == "window_fft_warp"
    nwin = bars - n + 1
    assert got["bins"].shape == (nwin, 8)
    assert got["bins"].min() >= 40 and got["bins"].max() <= 113
    ocfg = ocfg_from(oracle, cfg)
    for w in list(range(0, nwin, 83_161)) + [nwin - 1]:
        ref = oracle.pipeline_series(s[w:w + n], ocfg, oracle.OUT_BINS | oracle.OUT_WAVES)
        assert np.array_equal(got["bins"][w], ref["bins"][0]), w
        scale = np.abs(ref["waves"][0]).max()
        assert np.abs(got["waves"][w] - ref["waves"][0]).max() <= REL_TOL * scale, w
    refk = oracle.pipeline_series(s, ocfg, oracle.OUT_KALMAN)["kalman"]
    assert np.array_equal(got["kalman"], refk)
    got4 = br.pipeline_host(4.0 * s, cfg, br.OUT_BINS)
    assert np.array_equal(got4["bins"], got["bins"])


def test_config1_full_size_all_windows_against_oracle(br, oracle):
    """Config 1 is small enough for the oracle to run every window: 100k bars, N=512, mean + Hann,
    top-5, band 9-200 — all 99 489 windows, bins exact, spectra within 1e-9."""
    n, bars = 512, 100_000
    s = synth.random_walk(0, bars)
    cfg = br.default_cfg(n, top_k=5, min_period=9.0, max_period=200.0, detrend=br.DETREND_MEAN,
                         window_type=br.WINDOW_HANN_WIP)
    got = br.pipeline_host(s, cfg, br.OUT_BINS | br.OUT_SPECTRA)
    assert br.last_kernel() == "window_fft_warp"
    done, ref = oracle.pipeline_batch_mt(s.reshape(1, -1), ocfg_from(oracle, cfg), os.cpu_count() or 1,
                                         want=("bins", "spectra"))
    assert done == bars - n + 1
    assert np.array_equal(got["bins"], ref["bins"][0])
    err = np.abs(got["spectra"] - ref["spectra"][0]).max(axis=1) / np.abs(ref["spectra"][0]).max(axis=1)
    assert err.max() < REL_TOL


# ---- A6 behind the FFT kernels (ws_phase.cu) ----------------------------------------------------
def _check_phase_planes(got, ref, name="phase"):
    """phase = atan2(im, re) per bin: the 1e-9 bar on the spectrum propagates to
    |d phase_k| <= 1e-9 * max|X| / |X_k| (bins far below the window's largest are ill-conditioned by
    exactly that ratio); with no spectra plane at hand the flat 1e-6 of round 1 is kept.  Unwrapped
    phase and group delay: see check_unwrap_and_group_delay (windows on a +-pi decision are counted
    and compared modulo 2 pi)."""
    gp, rp = got["phase"], ref["phase"]
    dph = np.abs(np.angle(np.exp(1j * (gp[:, 0] - rp[:, 0]))))
    if "spectra" in ref:
        sp = ref["spectra"]
        mag = np.hypot(sp[:, 0::2], sp[:, 1::2])
        tol = REL_TOL * mag.max(axis=1, keepdims=True) / np.where(mag > 0, mag, np.inf)
        tol = np.where(mag > 0, np.maximum(tol, 1e-15), np.pi)      # a bin that is exactly 0 has no phase
        assert np.all(dph <= tol), float((dph - tol).max())
    else:
        assert dph.max() < 1e-6
    REPORT["max_phase_plane_err_" + name] = max(REPORT.get("max_phase_plane_err_" + name, 0.0), float(dph.max()))
    check_unwrap_and_group_delay(name, gp, rp, utol=1e-9 * max(1.0, np.abs(rp[:, 1]).max()), gtol=1e-7)


@pytest.mark.parametrize("with_spectra", [True, False])
def test_phase_chain_behind_the_sliding_kernel(br, oracle, with_spectra, monkeypatch):
    """Config-4 shape on plain hop-1 windows: the sliding kernel produces spectra + rows, the phase
    kernel follows on the plane (the caller's, or a scratch plane walked in window ranges)."""
    monkeypatch.setenv("WAVESPEC_PHASE_CHUNK", "37")          # several ragged window ranges
    s = synth.random_walk(1500, 1024 + 300)
    cfg = br.default_cfg(1024, top_k=8, min_period=18.0, max_period=200.0)
    outs = br.OUT_PHASE | br.OUT_BINS | br.OUT_WAVES | br.OUT_ROWS | (br.OUT_SPECTRA if with_spectra else 0)
    got, ref = run_both(br, oracle, s, cfg, outs)
    assert br.last_kernel() in ("sliding_shared", "sliding_overlap", "sliding_staged")
    check_planes(br, got, ref, cfg)
    _check_phase_planes(got, ref)


@pytest.mark.parametrize("n", [64, 512, 2048, 8192])
def test_phase_chain_behind_the_per_window_kernels(br, oracle, n, monkeypatch):
    monkeypatch.setenv("WAVESPEC_PHASE_CHUNK", "5")
    s = synth.random_walk(1501 + n, n + 23)
    cfg = br.default_cfg(n, top_k=4, min_period=9.0, max_period=200.0, detrend=br.DETREND_MEAN,
                         window_type=br.WINDOW_HANN)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_PHASE | br.OUT_BINS)
    check_planes(br, got, ref, cfg)
    _check_phase_planes(got, ref)


def test_phase_chain_two_series_batch(br, oracle):
    s = synth.random_walk_batch(1510, 2, 1024 + 90)
    cfg = br.default_cfg(1024, top_k=8, min_period=18.0, max_period=200.0)
    got = br.pipeline_host(s, cfg, br.OUT_PHASE | br.OUT_SPECTRA)
    for i in range(2):
        ref = oracle.pipeline_series(s[i], ocfg_from(oracle, cfg), br.OUT_PHASE | br.OUT_SPECTRA)
        _check_phase_planes({"phase": got["phase"][i]}, ref)
        assert rel_err(got["spectra"][i], ref["spectra"]) < REL_TOL


@pytest.mark.parametrize("top_k,stride", [(1, 15), (12, 4), (32, 20), (3, 8)])
def test_warp_kernel_top_k_and_row_strides(br, oracle, top_k, stride):
    """K from 1 to the maximum (32) and the older row layouts through the warp-per-window kernel."""
    s = synth.random_walk(1600 + top_k, 1024 + 77)
    cfg = br.default_cfg(1024, top_k=top_k, row_stride=stride, min_period=6.0, max_period=300.0,
                         detrend=br.DETREND_MEAN, window_type=br.WINDOW_HAMMING)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES)
    assert br.last_kernel() == "window_fft_warp"
    check_planes(br, got, ref, cfg)


@pytest.mark.parametrize("top_k,stride", [(1, 15), (4, 15), (12, 4), (32, 20)])
def test_sliding_kernels_top_k_and_row_strides(br, oracle, top_k, stride):
    """Same on the plain hop-1 path: K <= 8 with a narrow band takes the network (and the producer /
    consumer kernel), larger K the per-window scan."""
    s = synth.random_walk(1650 + top_k, 1024 + 130)
    cfg = br.default_cfg(1024, top_k=top_k, row_stride=stride, min_period=18.0, max_period=200.0)
    got, ref = run_both(br, oracle, s, cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_WAVES)
    assert br.last_kernel() in ("sliding_shared", "sliding_overlap", "sliding_staged")
    check_planes(br, got, ref, cfg)
