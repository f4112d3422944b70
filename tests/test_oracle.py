"""CPU tests of the oracle itself: analytic known answers and independent numpy/scipy
restatements.  The reference ships no vectors, so these are the pins.
Each block names the reference lines the oracle function follows."""
import numpy as np
import pytest

from fft_wavespec_b200 import synth


# ---- A4 FourierTransformManual, Legacy/WaveSpecZZ_1.0.2.mq5:938-975 ---------------------------
@pytest.mark.parametrize("n", [2, 4, 64, 256, 512, 1024, 2048, 4096])
def test_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    re, im = oracle.fft_forward(x)
    X = np.fft.fft(x)
    assert np.abs(re + 1j * im - X).max() / np.abs(X).max() < 1e-12


def test_fft_pure_sinusoid_known_answer(oracle):
    n, k, A = 1024, 37, 2.5
    t = np.arange(n)
    x = A * np.cos(2 * np.pi * k * t / n + 0.3)
    re, im = oracle.fft_forward(x)
    p = oracle.power(re, im)
    assert int(np.argmax(p)) == k
    assert p[k] == pytest.approx((A * n / 2) ** 2, rel=1e-12)
    assert np.arctan2(im[k], re[k]) == pytest.approx(0.3, abs=1e-12)
    others = np.delete(p, k)
    assert others.max() < 1e-18 * p[k]


def test_fft_impulse_and_parseval(oracle):
    n = 512
    x = np.zeros(n); x[0] = 3.0
    re, im = oracle.fft_forward(x)
    assert np.allclose(re, 3.0) and np.allclose(im, 0.0)
    y = np.random.default_rng(1).standard_normal(n)
    re, im = oracle.fft_forward(y)
    assert (re ** 2 + im ** 2).sum() / n == pytest.approx((y ** 2).sum(), rel=1e-12)


def test_fft_constant_gives_exact_zero_bins(oracle):
    # flat market: every non-DC bin is exactly 0 -> ties resolved by bin order downstream
    re, im = oracle.fft_forward(np.full(1024, 1.23456))
    assert np.all(re[1:] == 0.0) and np.all(im[1:] == 0.0)
    assert re[0] == pytest.approx(1024 * 1.23456, rel=1e-15)


def test_fft_interleaved_contract(oracle):
    # Legacy/WaveSpecZZ_1.0.4-new.mq5:3171-3194: out[2k]=re, out[2k+1]=im, k<n/2
    x = np.random.default_rng(2).standard_normal(256)
    out = oracle.fft_interleaved(x)
    re, im = oracle.fft_forward(x)
    assert np.array_equal(out[0::2], re[:128]) and np.array_equal(out[1::2], im[:128])


# ---- A3 windows, Legacy/...-kalman-fast.mq5:1126-1177 ------------------------------------------
@pytest.mark.parametrize("wtype,ref", [(1, np.hanning), (2, np.hamming), (3, np.blackman), (4, np.bartlett)])
def test_windows_match_numpy_symmetric_forms(oracle, wtype, ref):
    n = 512
    w = oracle.apply_window(np.ones(n), wtype)
    assert np.abs(w - ref(n)).max() < 1e-15 * 4
    assert w[0] == pytest.approx(w[-1], abs=1e-15)


def test_window_none_and_wip_hann(oracle):
    x = np.arange(16.0)
    assert np.array_equal(oracle.apply_window(x, 0), x)
    assert np.abs(oracle.apply_window(np.ones(64), 5) - np.hanning(64)).max() < 1e-15


# ---- A2a trend IIR, Legacy/...-kalman-fast.mq5:3367-3379 ---------------------------------------
def test_detrend_iir_matches_lfilter(oracle):
    from scipy.signal import lfilter
    x = synth.random_walk(3, 2048)
    T = 1024
    om = 2 * np.pi / T
    alpha = (1 - np.sin(om)) / np.cos(om)
    c = (1 - alpha) / 2
    # y[j] = c x[j] + c x[j-1] + alpha y[j-1], with x[-1]=x[0], y[-1]=0
    xx = np.concatenate([[x[0]], x])
    y = lfilter([c, c], [1, -alpha], xx)[1:] - (alpha ** np.arange(1, x.size + 1)) * (c * x[0])
    tr, d = oracle.detrend_iir(x, T)
    assert tr[0] == c * (x[0] + x[0])
    assert np.abs(tr - y).max() < 1e-12
    assert np.array_equal(d, x - tr)


# ---- A2b mean removal + Hann, Legacy/WaveSpecZZ_gpu_wip.mq5:935-957 ----------------------------
def test_mean_hann(oracle):
    x = synth.random_walk(4, 512)
    out = oracle.mean_hann(x)
    assert np.abs(out - (x - x.mean()) * np.hanning(512)).max() < 1e-15


# ---- A6 phase chain, Legacy/...-kalman-fast.mq5:1183-1263 --------------------------------------
def test_phase_chain_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    re = rng.standard_normal(256); im = rng.standard_normal(256)
    ph, un, gd = oracle.phase_chain(re, im, 256)
    assert np.abs(ph - np.arctan2(im, re)).max() < 1e-15
    assert np.abs(un - np.unwrap(ph)).max() < 1e-9
    g = -np.gradient(un)
    assert np.abs(gd - np.clip(g, -100, 100)).max() < 1e-12


def test_group_delay_of_pure_delay(oracle):
    n, d = 256, 5
    x = np.zeros(n); x[d] = 1.0
    re, im = oracle.fft_forward(x)
    _, _, gd = oracle.phase_chain(re, im, n // 2)
    # phase = -2 pi k d / n  -> -dphi/dk = 2 pi d / n
    assert np.allclose(gd[1:-1], 2 * np.pi * d / n, atol=1e-9)


# ---- A7a insertion top-K, Legacy/...-gpuopt-nodetrend.mq5:537-554 ------------------------------
def test_topk_insertion_band_and_order(oracle):
    n = 1024
    sp = np.zeros(n // 2)
    sp[[6, 10, 56, 57, 5, 30]] = [3.0, 9.0, 4.0, 100.0, 100.0, 1.0]
    b, p = oracle.topk_insertion(sp, n, 18, 200, 8)
    # band = [ceil(1024/200), floor(1024/18)] = [6, 56]; bins 5 and 57 are outside
    assert list(b[:4]) == [10, 56, 6, 30]
    assert list(p[:4]) == [9.0, 4.0, 3.0, 1.0]
    # remaining slots take the zero-power bins in ascending order (0 > -1 inserts, ties keep order)
    assert list(b[4:]) == [7, 8, 9, 11]


def test_topk_insertion_ties_prefer_lower_bin(oracle):
    n = 1024
    sp = np.zeros(n // 2); sp[[20, 12, 40]] = 7.0
    b, _ = oracle.topk_insertion(sp, n, 18, 200, 3)
    assert list(b) == [12, 20, 40]


def test_topk_band_clipped_to_half_spectrum(oracle):
    n = 64
    sp = np.arange(n // 2, dtype=float)
    b, _ = oracle.topk_insertion(sp, n, 1.0, 200, 4)    # floor(64/1)=64 -> clipped to 31
    assert list(b) == [31, 30, 29, 28]


# ---- A7b selection sort, Legacy/WaveSpecZZ_1.0.4-kalman.mq5:143-180 ----------------------------
def test_collect_sorted_descending_and_k_ge_1(oracle):
    n = 256
    rng = np.random.default_rng(7)
    re = rng.standard_normal(n); im = rng.standard_normal(n)
    idx, pw = oracle.collect_sorted(re, im, 2.0, 1e9)   # ceil(n/maxP)=1 -> k>=1 ; max floor(128)->127
    assert idx.min() == 1 and idx.max() == 127 and idx.size == 127
    assert np.all(np.diff(pw) <= 0)
    p = re[:128] ** 2 + im[:128] ** 2
    assert np.array_equal(pw, p[idx])


def test_collect_sorted_swap_order_on_ties(oracle):
    # swap-based selection sort is not stable: [a,b,c,c'] with c==c' largest
    n = 64
    re = np.zeros(n); im = np.zeros(n)
    re[4], re[5], re[6], re[7] = 1.0, 2.0, 3.0, 3.0
    idx, pw = oracle.collect_sorted(re, im, 64 / 7.0, 64 / 4.0)   # bins 4..7
    # pass 0: max=bin6 (first of the tie) swaps with bin4 -> [6,5,4,7]; pass 1: max among [5,4,7]
    # is bin7 -> swaps with bin5 -> [6,7,4,5]; pass 2: max among [4,5] = bin5 -> [6,7,5,4]
    assert list(idx) == [6, 7, 5, 4]


# ---- A8a / A8b reconstruction ------------------------------------------------------------------
def test_reconstruction_forms_agree(oracle):
    n = 1024
    x = synth.random_walk(8, n)
    re, im = oracle.fft_forward(x)
    p = oracle.power(re, im)
    for k in (6, 17, 56):
        wave, period = oracle.recon_last(re, im, k, p[k])
        contrib = oracle.contribution(re, im, k)
        assert period == n / k
        assert contrib == pytest.approx(2 * wave, rel=1e-9, abs=1e-18)
    # a pure in-bin cosine is reproduced at the last sample by the single-bin inverse DFT
    t = np.arange(n)
    y = 0.7 * np.cos(2 * np.pi * 9 * t / n + 1.1)
    re, im = oracle.fft_forward(y)
    assert oracle.contribution(re, im, 9) == pytest.approx(y[-1], abs=1e-12)


# ---- A9 Kalman4D, Legacy/...-kalman-fast.mq5:2015-2125 -----------------------------------------
def test_kalman4d_tracks_constant_and_ramp(oracle):
    out = oracle.kalman4d_series(np.full(400, 1.2345))
    assert abs(out[-1] - 1.2345) < 1e-9
    ramp = 1.0 + 0.001 * np.arange(3000)
    out = oracle.kalman4d_series(ramp)
    assert abs(out[-1] - ramp[-1]) < 1e-3


def test_kalman4d_against_matrix_form(oracle):
    """Independent restatement with F, H matrices; adaptive boost and clip included."""
    z = synth.random_walk(9, 300)
    p = oracle.kalman4d_defaults()
    F = np.array([[1, 1, .5, 1 / 6], [0, 1, 1, .5], [0, 0, 1, 1], [0, 0, 0, 1.0]])
    Q = np.diag([p.q_pos, p.q_vel, p.q_acc, p.q_jerk])
    x = np.array([z[0], 0, 0, 0.0]); P = np.diag([16.0, 9, 4, 1])
    ref = []
    for zi in z:
        x = F @ x
        P0 = P
        P = F @ P @ F.T + Q
        # reference quirk, reproduced: P11' is written with doubled P12/P22 terms and undamped
        # P13/P23 (Legacy/...-kalman-fast.mq5:2052), which is not (F P F^T)[1,1]
        P[1, 1] = (P0[1, 1] + 2 * P0[1, 2] + P0[1, 3] + P0[2, 1] + 2 * P0[2, 2] + P0[2, 3]
                   + 0.5 * P0[3, 1] + 0.5 * P0[3, 2] + 0.25 * P0[3, 3] + p.q_vel)
        # the reference mirrors the upper triangle into the lower one after prediction (:2057-2058)
        P = np.triu(P) + np.triu(P, 1).T
        y = zi - x[0]; S = P[0, 0] + p.meas_noise
        k = min(5.0, abs(y) / np.sqrt(S)) * p.adapt_gain
        P = P + k * Q; S = P[0, 0] + p.meas_noise
        lim = p.clip_std * np.sqrt(S); y = np.clip(y, -lim, lim)
        K = P[:, 0] / S
        x = x + K * y
        P = P - np.outer(K, P[0, :])
        for d in range(4):
            P[d, d] = max(P[d, d], 1e-12)
        ref.append(x[0])
    out = oracle.kalman4d_series(z, p)
    assert np.abs(out - np.array(ref)).max() < 1e-9


# ---- A10 weight Kalman, Legacy/WaveSpecZZ_1.0.4-kalman.mq5:194-231 -----------------------------
def test_weight_kalman_single_cycle_converges(oracle):
    import ctypes as C
    st = oracle.WKalmanState()
    oracle.lib().oracle_wkalman_reset(C.byref(st), 25.0)
    h = np.array([2.0])
    out = 0.0
    for _ in range(200):
        out = oracle.lib().oracle_wkalman_update(C.byref(st), h, 1, 6.0, 0.25, 9.0)
    assert out == pytest.approx(6.0, rel=1e-3)       # weight -> 3
    assert st.weights[0] == pytest.approx(3.0, rel=1e-3)


# ---- A11 PLA, Legacy/...-kalman-fast.mq5:387-502 -----------------------------------------------
def test_pla_piecewise_linear_input_recovers_joints(oracle):
    n = 256
    x = np.concatenate([np.full(60, 1.0), 1.0 + 0.002 * np.arange(1, 41), np.full(156, 1.08)])
    line, st, en, sl, ic = oracle.pla_build(x, 32, 0.0005)
    assert st[0] == 0 and en[0] == 59 and en[-1] == n - 1
    assert np.abs(line - x).max() < 1e-12
    assert np.all(st[1:] == en[:-1] + 1)


def test_pla_degenerate_split_at_segment_start_is_reproduced(oracle):
    """When the worst sample is the segment's first one, PlaSplit (:462-467) recurses on
    [s,s] and on the SAME [s,e] again, burning the segment budget with single-point segments.
    The oracle keeps that behaviour (it defines the reference's PLA line)."""
    x = np.concatenate([np.linspace(1.0, 1.1, 100), np.linspace(1.1, 1.0, 156)])
    line, st, en, sl, ic = oracle.pla_build(x, 32, 0.0005)
    assert st.size == 32 and np.all(st == 0) and np.all(en[:-1] == 0) and en[-1] == 255
    slope, icpt = np.polyfit(np.arange(256.0), x, 1)
    assert np.abs(line - (slope * np.arange(256.0) + icpt)).max() < 1e-12


def test_pla_segment_budget_and_flat(oracle):
    x = synth.random_walk(10, 1024)
    line, st, en, _, _ = oracle.pla_build(x, 8, 1e-8)
    assert 1 <= st.size <= 8 + 8 and st[0] == 0 and en[-1] == 1023
    flat = np.full(64, 1.5)
    line, st, en, sl, ic = oracle.pla_build(flat, 32, 0.0005)
    assert st.size == 1 and np.allclose(line, 1.5, atol=1e-12)


# ---- A12 ZigZag expansion ----------------------------------------------------------------------
def test_zigzag_feed_110_modes(oracle):
    main = np.zeros(10); main[[2, 5, 8]] = [1.0, 2.0, 1.5]
    hi = np.arange(10.0) + 1; lo = np.arange(10.0)
    step = oracle.zigzag_feed_110(main, hi, lo, 0)
    assert list(step) == [1, 1, 1, 1, 1, 2, 2, 2, 1.5, 1.5]
    interp = oracle.zigzag_feed_110(main, hi, lo, 1)
    assert interp[3] == pytest.approx(1 + 1 / 3) and interp[0] == 1.0 and interp[9] == 1.5
    mid = oracle.zigzag_feed_110(main, hi, lo, 2)
    assert np.array_equal(mid, (hi + lo) * 0.5)
    empty = oracle.zigzag_feed_110(np.zeros(4), hi[:4], lo[:4], 0, high0=3.0, low0=1.0)
    assert np.all(empty == 2.0)


def test_zigzag_series_legacy_modes(oracle):
    zz = np.zeros(10); zz[[2, 6]] = [1.0, 3.0]
    ok, cont = oracle.zigzag_series_legacy(zz, np.zeros(10), np.zeros(10), 0)
    assert ok and list(cont) == [1, 1, 1, 1.5, 2, 2.5, 3, 3, 3, 3]
    ok, alt = oracle.zigzag_series_legacy(zz, np.zeros(10), np.zeros(10), 1)
    assert ok and list(alt) == [1, 1, 1, 1, 1, 1, 3, 3, 3, 3]
    ok, _ = oracle.zigzag_series_legacy(np.zeros(10), np.zeros(10), np.zeros(10), 0)
    assert not ok


# ---- A13 tracker pool, Legacy/...-kalman-fast.mq5:1415-1667 -------------------------------------
def test_tracker_first_bar_drags_one_tracker_through_close_bins(oracle):
    """Neighbouring bins within 5 % of each other keep updating the SAME tracker (UpdateTracker
    rewrites the period later candidates are matched against), so bar 1 of a dense band leaves one
    tracker parked on the last bin."""
    n = 2048
    sp = np.ones(n // 2)
    t = oracle.Tracker()
    idx, per = t.step(sp, n, 18.0, 52.0)                 # band = bins 40..113
    assert len(t.trackers()) == 1 and t.trackers()[0][0] == 113
    assert idx[0] == 113 and per[0] == n / 113 and np.all(idx[1:] == 0) and np.all(per[1:] == 0.0)


def test_tracker_wide_bins_each_get_a_tracker_and_slots_fill_by_power(oracle):
    n = 64
    sp = np.zeros(n // 2); sp[4:9] = [1.0, 5.0, 3.0, 5.0, 2.0]        # periods 16, 12.8, 10.7, 9.1, 8 (>5 % apart)
    t = oracle.Tracker()
    idx, per = t.step(sp, n, 8.0, 16.0)
    assert [x[0] for x in t.trackers()] == [4, 5, 6, 7, 8]
    # descending power, equal powers keep bin order (stable bubble sort with '<')
    assert list(idx[:5]) == [5, 7, 6, 8, 4] and np.all(idx[5:] == 0)
    # slots are sticky: a new ordering of the powers does not move them
    sp[4:9] = [9.0, 1.0, 1.0, 1.0, 1.0]
    idx2, _ = t.step(sp, n, 8.0, 16.0)
    assert list(idx2[:5]) == [5, 7, 6, 8, 4]


def test_tracker_unseen_trackers_expire_and_slots_keep_raw_indices(oracle):
    n = 64
    sp = np.zeros(n // 2); sp[4:9] = [1.0, 5.0, 3.0, 4.0, 2.0]
    t = oracle.Tracker()
    t.step(sp, n, 8.0, 16.0)                              # trackers for bins 4..8
    for _ in range(3):                                    # band shrinks: bins 4,5 are no longer candidates
        idx, per = t.step(sp, n, 8.0, 10.7)               # band = bins 6..8
    assert [x[0] for x in t.trackers()] == [6, 7, 8]      # bins 4, 5 erased after 3 inactive bars
    # erase shifts the array but the slot table keeps raw indices (:1514-1519, :1584-1589):
    # slot 0 pointed at tracker index 1 (bin 5) and now silently shows what slid into index 1
    assert idx[0] == 7


# ---- pipeline (bar loop) ------------------------------------------------------------------------
def test_pipeline_series_matches_stagewise_calls(oracle):
    n = 256
    s = synth.random_walk(11, 600)
    cfg = oracle.default_cfg(n, detrend=1, trend_period=128.0, window_type=3, min_period=9.0,
                             max_period=60.0, top_k=5,
                             outputs=oracle.OUT_SPECTRA | oracle.OUT_BINS | oracle.OUT_WAVES |
                             oracle.OUT_KALMAN | oracle.OUT_ROWS | oracle.OUT_PHASE)
    r = oracle.pipeline_series(s, cfg)
    nw = 600 - n + 1
    assert r["spectra"].shape == (nw, n)
    for w in (0, 7, nw - 1):
        _, d = oracle.detrend_iir(s[w:w + n], 128.0)
        d = oracle.apply_window(d, 3)
        re, im = oracle.fft_forward(d)
        assert np.array_equal(r["spectra"][w, 0::2], re[:n // 2])
        b, p = oracle.topk_insertion(oracle.power(re, im), n, 9.0, 60.0, 5)
        assert np.array_equal(r["bins"][w], b)
        wv, _ = oracle.recon_last(re, im, b[0], p[0])
        assert r["waves"][w, 0] == wv
        row = r["rows"][w, 0]
        assert row[0] * np.sin(row[3]) == pytest.approx(oracle.contribution(re, im, b[0]), rel=1e-9)
        assert row[2] == n / b[0] and row[14] == 0.0
        ph, un, gd = oracle.phase_chain(re, im, n // 2)
        assert np.array_equal(r["phase"][w, 0], ph) and np.array_equal(r["phase"][w, 2], gd)
    assert np.array_equal(r["kalman"], oracle.kalman4d_series(s[n - 1:]))


def test_pipeline_batch_mt_equals_series_loop(oracle):
    n = 128
    batch = synth.random_walk_batch(20, 3, 700)
    cfg = oracle.default_cfg(n, min_period=9.0, max_period=60.0, top_k=4)
    done, out = oracle.pipeline_batch_mt(batch, cfg, threads=3, want=("bins", "spectra"))
    assert done == 3 * (700 - n + 1)
    for i in range(3):
        r = oracle.pipeline_series(batch[i], cfg, oracle.OUT_BINS | oracle.OUT_SPECTRA)
        assert np.array_equal(out["bins"][i], r["bins"])
        assert np.array_equal(out["spectra"][i], r["spectra"])


def test_applied_price_matches_the_reference_expressions(oracle):
    """A1 (…-fast.mq5:3308-3316): operand order matters for the last bit — (h+l+c)/3, not h/3+l/3+c/3."""
    c = synth.random_walk(77, 4000)
    h, l = synth.high_low(77, c)
    o = np.roll(c, 1); o[0] = c[0]
    assert np.array_equal(oracle.applied_price(o, h, l, c, 1), c)
    assert np.array_equal(oracle.applied_price(o, None, None, None, 2), o)
    assert np.array_equal(oracle.applied_price(None, h, None, None, 3), h)
    assert np.array_equal(oracle.applied_price(None, None, l, None, 4), l)
    assert np.array_equal(oracle.applied_price(None, h, l, None, 5), (h + l) / 2.0)
    assert np.array_equal(oracle.applied_price(None, h, l, c, 6), (h + l + c) / 3.0)
    assert np.array_equal(oracle.applied_price(None, h, l, c, 7), (h + l + 2 * c) / 4.0)
    assert not np.array_equal(oracle.applied_price(None, h, l, c, 6), h / 3.0 + l / 3.0 + c / 3.0)
    with pytest.raises(ValueError):
        oracle.applied_price(o, h, l, c, 9)


# ---- A8d: inverse of the gpu_fft_real_forward contract ------------------------------------------
@pytest.mark.parametrize("n", [4, 16, 256, 1024, 4096])
def test_fft_inverse_known_answers(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    spec = oracle.fft_interleaved(x)
    back = oracle.fft_inverse(spec)
    # the forward contract drops the Nyquist bin: the round trip loses exactly that component
    nyq = np.sum(x * (-1.0) ** np.arange(n))
    assert np.abs(back - (x - nyq * (-1.0) ** np.arange(n) / n)).max() < 1e-12
    # independent restatement: numpy's irfft of the same half spectrum with a zero Nyquist bin
    half = np.zeros(n // 2 + 1, dtype=complex)
    half[:n // 2] = spec[0::2] + 1j * spec[1::2]
    half[0] = half[0].real
    assert np.abs(back - np.fft.irfft(half, n)).max() < 1e-12
    # a single cosine bin comes back as that cosine with amplitude 2/n per unit of |X|
    k = n // 4 - 1 if n > 4 else 1
    one = np.zeros(n)
    one[2 * k] = 3.0; one[2 * k + 1] = -4.0
    t = np.arange(n)
    expect = (2.0 / n) * (3.0 * np.cos(2 * np.pi * k * t / n) + 4.0 * np.sin(2 * np.pi * k * t / n))
    assert np.abs(oracle.fft_inverse(one) - expect).max() < 1e-13
