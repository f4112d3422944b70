"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/wavespec_oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
fft_wavespec_b200 package.  Parity is unpinned by the reference (it ships no vectors).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Kalman4DParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "follow_strength", "q_pos", "q_vel", "q_acc", "q_jerk", "adapt_gain", "meas_noise",
        "init_var_pos", "init_var_vel", "init_var_acc", "init_var_jerk",
        "init_vel", "init_acc", "init_jerk", "clip_std", "ema_blend_period")]


class PipelineCfg(C.Structure):
    _fields_ = [
        ("window_len", C.c_int32), ("hop", C.c_int32), ("top_k", C.c_int32), ("row_stride", C.c_int32),
        ("min_period", C.c_double), ("max_period", C.c_double), ("sample_rate_seconds", C.c_double),
        ("feed", C.c_int32), ("detrend", C.c_int32), ("trend_period", C.c_double),
        ("window_type", C.c_int32), ("select", C.c_int32), ("pla_max_segments", C.c_int32),
        ("outputs", C.c_int32), ("pla_max_error", C.c_double),
        ("wk_process_noise", C.c_double), ("wk_meas_noise", C.c_double), ("wk_init_variance", C.c_double),
        ("kalman", Kalman4DParams),
        ("tracker_tolerance", C.c_double), ("tracker_max_inactive", C.c_int32), ("reserved0", C.c_int32)]


class WKalmanState(C.Structure):
    _fields_ = [("weights", C.c_double * 32), ("cov", C.c_double * 32)]


class TrackerState(C.Structure):
    _fields_ = [("count", C.c_int), ("period", C.c_double * 512), ("power", C.c_double * 512),
                ("fft_index", C.c_int * 512), ("is_active", C.c_int * 512), ("bars_inactive", C.c_int * 512),
                ("slot", C.c_int * 12)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("wavespec_oracle.cpp", "wavespec_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.oracle_fft_forward.argtypes = [_dp, C.c_int, _dp, _dp]
    L.oracle_fft_interleaved.argtypes = [_dp, C.c_int, _dp]
    L.oracle_fft_inverse.argtypes = [_dp, C.c_int, _dp]
    L.oracle_apply_window.argtypes = [_dp, C.c_int, C.c_int]
    L.oracle_detrend_iir.argtypes = [_dp, C.c_int, C.c_double, _dp, _dp]
    L.oracle_mean_hann.argtypes = [_dp, C.c_int, _dp]
    L.oracle_power.argtypes = [_dp, _dp, C.c_int, _dp]
    L.oracle_phase_chain.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _dp]
    L.oracle_topk_insertion.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, _ip, _dp]
    L.oracle_collect_sorted.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double, _ip, _dp]
    L.oracle_collect_sorted.restype = C.c_int
    L.oracle_recon_last.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.oracle_contribution.argtypes = [_dp, _dp, C.c_int, C.c_int]
    L.oracle_contribution.restype = C.c_double
    L.oracle_kalman4d_defaults.argtypes = [C.POINTER(Kalman4DParams)]
    L.oracle_kalman4d_series.argtypes = [_dp, C.c_int, C.POINTER(Kalman4DParams), _dp]
    L.oracle_wkalman_reset.argtypes = [C.POINTER(WKalmanState), C.c_double]
    L.oracle_wkalman_update.argtypes = [C.POINTER(WKalmanState), _dp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.oracle_wkalman_update.restype = C.c_double
    L.oracle_pla_build.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp, _ip, _ip, _dp, _dp]
    L.oracle_pla_build.restype = C.c_int
    L.oracle_zigzag_feed_110.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
    L.oracle_zigzag_series_legacy.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, _dp]
    L.oracle_zigzag_series_legacy.restype = C.c_int
    L.oracle_applied_price.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, _dp]
    L.oracle_applied_price.restype = C.c_int
    L.oracle_tracker_reset.argtypes = [C.POINTER(TrackerState)]
    L.oracle_tracker_step.argtypes = [C.POINTER(TrackerState), _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_int, _ip, _dp]
    L.oracle_cycle_cache.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                     C.c_int, C.c_double, C.c_double, C.c_double, _dp]
    L.oracle_default_cfg.argtypes = [C.POINTER(PipelineCfg), C.c_int]
    L.oracle_pipeline_series.argtypes = [_dp, C.c_int, C.POINTER(PipelineCfg)] + [C.c_void_p] * 7
    L.oracle_pipeline_series_trk.argtypes = [_dp, C.c_int, C.POINTER(PipelineCfg)] + [C.c_void_p] * 9
    L.oracle_pipeline_batch_mt.argtypes = [_dp, C.c_int, C.c_int, C.POINTER(PipelineCfg), C.c_int,
                                           C.c_int64] + [C.c_void_p] * 4
    L.oracle_pipeline_batch_mt.restype = C.c_int64
    _LIB = L
    return L


# ----------------------------------------------------------------------------------------------
# numpy-level helpers
def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def fft_forward(x):
    x = _f64(x); n = x.size
    re = np.empty(n); im = np.empty(n)
    lib().oracle_fft_forward(x, n, re, im)
    return re, im


def fft_interleaved(x):
    x = _f64(x); out = np.empty(x.size)
    lib().oracle_fft_interleaved(x, x.size, out)
    return out


def fft_inverse(spec):
    """n/2 interleaved bins -> n real samples (1/n normalised, Nyquist taken as 0)."""
    s = _f64(spec)
    out = np.empty(s.size)
    lib().oracle_fft_inverse(s, s.size, out)
    return out


def apply_window(x, wtype):
    y = _f64(x).copy()
    lib().oracle_apply_window(y, y.size, int(wtype))
    return y


def detrend_iir(x, trend_period):
    x = _f64(x); tr = np.empty_like(x); d = np.empty_like(x)
    lib().oracle_detrend_iir(x, x.size, float(trend_period), tr, d)
    return tr, d


def mean_hann(x):
    x = _f64(x); out = np.empty_like(x)
    lib().oracle_mean_hann(x, x.size, out)
    return out


def power(re, im):
    n = re.size; sp = np.empty(n // 2)
    lib().oracle_power(_f64(re), _f64(im), n, sp)
    return sp


def phase_chain(re, im, count):
    ph = np.empty(count); un = np.empty(count); gd = np.empty(count)
    lib().oracle_phase_chain(_f64(re), _f64(im), count, ph, un, gd)
    return ph, un, gd


def topk_insertion(spectrum, n, min_period, max_period, top_k=8):
    b = np.empty(top_k, dtype=np.int32); p = np.empty(top_k)
    lib().oracle_topk_insertion(_f64(spectrum), n, float(min_period), float(max_period), top_k, b, p)
    return b, p


def collect_sorted(re, im, min_period, max_period):
    n = re.size
    idx = np.empty(n // 2, dtype=np.int32); pw = np.empty(n // 2)
    c = lib().oracle_collect_sorted(_f64(re), _f64(im), n, float(min_period), float(max_period), idx, pw)
    return idx[:c].copy(), pw[:c].copy()


def recon_last(re, im, bin_, pw):
    w = C.c_double(); p = C.c_double()
    lib().oracle_recon_last(_f64(re), _f64(im), re.size, int(bin_), float(pw), C.byref(w), C.byref(p))
    return w.value, p.value


def contribution(re, im, k):
    return lib().oracle_contribution(_f64(re), _f64(im), re.size, int(k))


def kalman4d_defaults():
    p = Kalman4DParams()
    lib().oracle_kalman4d_defaults(C.byref(p))
    return p


def kalman4d_series(z, params=None):
    z = _f64(z); out = np.empty_like(z)
    p = params or kalman4d_defaults()
    lib().oracle_kalman4d_series(z, z.size, C.byref(p), out)
    return out


def pla_build(window, max_segments=32, max_error=0.0005):
    w = _f64(window); n = w.size
    line = np.zeros(n); st = np.empty(n + 4, dtype=np.int32); en = np.empty(n + 4, dtype=np.int32)
    sl = np.empty(n + 4); ic = np.empty(n + 4)
    c = lib().oracle_pla_build(w, n, max_segments, max_error, line, st, en, sl, ic)
    return line, st[:c].copy(), en[:c].copy(), sl[:c].copy(), ic[:c].copy()


def zigzag_feed_110(main_ch, high_ch, low_ch, mode, high0=0.0, low0=0.0):
    m = _f64(main_ch); out = np.empty_like(m)
    lib().oracle_zigzag_feed_110(m, _f64(high_ch), _f64(low_ch), m.size, int(mode), high0, low0, out)
    return out


def applied_price(open_, high, low, close, mode):
    """A1: per-bar applied price (MQL5 ENUM_APPLIED_PRICE value as mode); unused series may be None."""
    arrs = [None if a is None else _f64(a) for a in (open_, high, low, close)]
    n = next(a.size for a in arrs if a is not None)
    out = np.empty(n)
    ptrs = [None if a is None else C.c_void_p(a.ctypes.data) for a in arrs]
    if lib().oracle_applied_price(*ptrs, n, int(mode), out) != 0:
        raise ValueError("unknown applied-price mode")
    return out


def zigzag_series_legacy(zz_main, zz_high, zz_low, mode):
    m = _f64(zz_main); out = np.zeros_like(m)
    ok = lib().oracle_zigzag_series_legacy(m, _f64(zz_high), _f64(zz_low), m.size, int(mode), out)
    return bool(ok), out


def cycle_cache(rows, top_k, window_len, hop, bars, period_seconds=60.0, music_only=False, use_music_weights=False,
                min_coherence=0.05, min_score=0.01, min_snr_db=-40.0):
    """rows: [n_windows*top_k, stride] -> [bars, 20] cache record (EMPTY_VALUE = DBL_MAX where unwritten)."""
    r = _f64(rows).reshape(-1, np.asarray(rows).shape[-1])
    out = np.empty((bars, 20))
    lib().oracle_cycle_cache(r, r.shape[0], r.shape[1], top_k, window_len, hop, bars, period_seconds,
                             int(music_only), int(use_music_weights), min_coherence, min_score, min_snr_db, out)
    return out


class Tracker:
    """Period tracker pool + 12 stable slots (A13), one bar at a time."""

    def __init__(self):
        self.st = TrackerState()
        lib().oracle_tracker_reset(C.byref(self.st))

    def step(self, spectrum, n, min_period, max_period, tol=5.0, max_inactive=3):
        idx = np.zeros(12, dtype=np.int32); per = np.zeros(12)
        lib().oracle_tracker_step(C.byref(self.st), _f64(spectrum), n, float(min_period), float(max_period),
                                  float(tol), int(max_inactive), idx, per)
        return idx, per

    def trackers(self):
        c = self.st.count
        return [(self.st.fft_index[i], self.st.period[i], self.st.power[i], self.st.bars_inactive[i]) for i in range(c)]


def default_cfg(window_len, **over):
    cfg = PipelineCfg()
    lib().oracle_default_cfg(C.byref(cfg), int(window_len))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


OUT_SPECTRA, OUT_ROWS, OUT_BINS, OUT_WAVES, OUT_KALMAN, OUT_PHASE, OUT_WKALMAN, OUT_TRACKER = 1, 2, 4, 8, 16, 32, 64, 128


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_windows(series_len, window_len, hop):
    return 0 if series_len < window_len else 1 + (series_len - window_len) // hop


def pipeline_series(series, cfg, outputs=None):
    """Run the bar loop over one series; returns a dict of the requested planes."""
    s = _f64(series)
    outputs = cfg.outputs if outputs is None else outputs
    n, K = cfg.window_len, cfg.top_k
    nw = num_windows(s.size, n, cfg.hop)
    o = {
        "spectra": np.zeros((nw, n)) if outputs & OUT_SPECTRA else None,
        "rows": np.zeros((nw, K, cfg.row_stride)) if outputs & OUT_ROWS else None,
        "bins": np.full((nw, K), -1, dtype=np.int32) if outputs & OUT_BINS else None,
        "waves": np.zeros((nw, K)) if outputs & OUT_WAVES else None,
        "kalman": np.zeros(nw) if outputs & OUT_KALMAN else None,
        "phase": np.zeros((nw, 3, n // 2)) if outputs & OUT_PHASE else None,
        "wkalman": np.zeros(nw) if outputs & OUT_WKALMAN else None,
        "trk_index": np.zeros((nw, 12), dtype=np.int32) if outputs & OUT_TRACKER else None,
        "trk_period": np.zeros((nw, 12)) if outputs & OUT_TRACKER else None,
    }
    lib().oracle_pipeline_series_trk(s, s.size, C.byref(cfg), _ptr(o["spectra"]), _ptr(o["rows"]),
                                     _ptr(o["bins"]), _ptr(o["waves"]), _ptr(o["kalman"]),
                                     _ptr(o["phase"]), _ptr(o["wkalman"]), _ptr(o["trk_index"]),
                                     _ptr(o["trk_period"]))
    return {k: v for k, v in o.items() if v is not None}


def pipeline_batch_mt(series2d, cfg, threads, max_windows=0, want=("bins",)):
    """Stateless planes over a [n_series, series_len] batch with `threads` host threads."""
    s = _f64(series2d)
    ns, sl = s.shape
    n, K = cfg.window_len, cfg.top_k
    nw = num_windows(sl, n, cfg.hop)
    o = {
        "spectra": np.zeros((ns, nw, n)) if "spectra" in want else None,
        "rows": np.zeros((ns, nw, K, cfg.row_stride)) if "rows" in want else None,
        "bins": np.full((ns, nw, K), -1, dtype=np.int32) if "bins" in want else None,
        "waves": np.zeros((ns, nw, K)) if "waves" in want else None,
    }
    done = lib().oracle_pipeline_batch_mt(s, ns, sl, C.byref(cfg), int(threads), int(max_windows),
                                          _ptr(o["spectra"]), _ptr(o["rows"]), _ptr(o["bins"]),
                                          _ptr(o["waves"]))
    return done, {k: v for k, v in o.items() if v is not None}
