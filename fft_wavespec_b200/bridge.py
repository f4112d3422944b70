"""Host-side mirror of the reference's DLL import block (Include/imports.mqh:5-21) over
libwavespec.so.  Function names, argument meaning and status codes are the reference's; numpy
arrays stand in for MQL5 `double &a[]` buffers.  There is no fallback: if the CUDA library is
missing or no device opens, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WAVESPEC_LIB") or os.path.join(_HERE, "libwavespec.so")   # WAVESPEC_LIB: profiling builds

# status codes: WaveCyclesBatchFetcher.mq5:15-21
OK, BAD_ARGS, BACKEND_UNAVAILABLE, TIMEOUT, INTERNAL_ERROR, NOT_READY, NO_MEM = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "OK", -1: "BAD_ARGS", -2: "BACKEND_UNAVAILABLE", -3: "TIMEOUT",
                -4: "INTERNAL_ERROR", -5: "NOT_READY", -6: "NO_MEM"}

FEED_CLOSE, FEED_PLA = 0, 1
DETREND_NONE, DETREND_IIR, DETREND_MEAN = 0, 1, 2
WINDOW_NONE, WINDOW_HANN, WINDOW_HAMMING, WINDOW_BLACKMAN, WINDOW_BARTLETT, WINDOW_HANN_WIP = range(6)
SELECT_INSERTION, SELECT_SORT = 0, 1
OUT_SPECTRA, OUT_ROWS, OUT_BINS, OUT_WAVES, OUT_KALMAN, OUT_PHASE, OUT_WKALMAN, OUT_TRACKER = 1, 2, 4, 8, 16, 32, 64, 128
OUT_CONTRIB = 256
ROW_FIELDS = 15


class Kalman4DParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "follow_strength", "q_pos", "q_vel", "q_acc", "q_jerk", "adapt_gain", "meas_noise",
        "init_var_pos", "init_var_vel", "init_var_acc", "init_var_jerk",
        "init_vel", "init_acc", "init_jerk", "clip_std", "ema_blend_period")]


class PipelineCfg(C.Structure):
    """wavespec_pipeline_cfg (include/wavespec_abi.h)."""
    _fields_ = [
        ("window_len", C.c_int32), ("hop", C.c_int32), ("top_k", C.c_int32), ("row_stride", C.c_int32),
        ("min_period", C.c_double), ("max_period", C.c_double), ("sample_rate_seconds", C.c_double),
        ("feed", C.c_int32), ("detrend", C.c_int32), ("trend_period", C.c_double),
        ("window_type", C.c_int32), ("select", C.c_int32), ("pla_max_segments", C.c_int32),
        ("outputs", C.c_int32), ("pla_max_error", C.c_double),
        ("wk_process_noise", C.c_double), ("wk_meas_noise", C.c_double), ("wk_init_variance", C.c_double),
        ("kalman", Kalman4DParams),
        ("tracker_tolerance", C.c_double), ("tracker_max_inactive", C.c_int32), ("reserved0", C.c_int32)]


class CacheParams(C.Structure):
    """wavespec_cache_params (include/wavespec_abi.h)."""
    _fields_ = [("music_only", C.c_int32), ("use_music_weights", C.c_int32), ("min_coherence", C.c_double),
                ("min_score", C.c_double), ("min_snr_db", C.c_double)]


class Planes(C.Structure):
    """wavespec_planes (include/wavespec_abi.h): output pointers, NULL = not wanted."""
    _fields_ = [(n, C.c_void_p) for n in ("spectra", "rows", "bins", "waves", "contrib", "kalman", "phase",
                                          "wkalman", "trk_index", "trk_period")]


class WaveSpecError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {text}")
        self.status = status


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def lib():
    """Load libwavespec.so; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the package has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    i32, i64, dbl, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    L.gpu_init.argtypes = [i32, i32]; L.gpu_init.restype = i32
    L.gpu_shutdown.argtypes = []; L.gpu_shutdown.restype = None
    L.gpu_fft_real_forward.argtypes = [vp, i32, vp]; L.gpu_fft_real_forward.restype = i32
    L.gpu_fft_real_inverse.argtypes = [vp, i32, vp]; L.gpu_fft_real_inverse.restype = i32
    L.gpu_fft_real_forward_batch.argtypes = [vp, i32, i32, vp]; L.gpu_fft_real_forward_batch.restype = i32
    L.gpu_extract_cycles.argtypes = [vp, i32, i32, dbl, dbl, dbl, i32, i32, vp, i32, i32, _ip]
    L.gpu_extract_cycles.restype = i32
    L.gpu_submit_extract_cycles.argtypes = [vp, i32, i32, dbl, dbl, dbl, i32, i32, C.POINTER(i64)]
    L.gpu_submit_extract_cycles.restype = i32
    L.gpu_try_get_cycles.argtypes = [i64, vp, i32, i32, _ip, _ip]; L.gpu_try_get_cycles.restype = i32
    L.gpu_submit_extract_cycles_batch.argtypes = [vp, i32, i32, i32, i32, dbl, dbl, dbl, i32, i32, i32, C.POINTER(i64)]
    L.gpu_submit_extract_cycles_batch.restype = i32
    L.gpu_try_get_cycles_batch.argtypes = [i64, vp, i32, _ip, _ip]; L.gpu_try_get_cycles_batch.restype = i32
    L.gpu_free_job.argtypes = [i64]; L.gpu_free_job.restype = i32
    L.gpu_get_last_error_w.argtypes = [C.POINTER(C.c_uint16), i32]; L.gpu_get_last_error_w.restype = i32
    L.wavespec_default_cfg.argtypes = [C.POINTER(PipelineCfg), i32]; L.wavespec_default_cfg.restype = None
    L.wavespec_num_windows.argtypes = [i32, i32, i32]; L.wavespec_num_windows.restype = i64
    L.wavespec_pipeline_host.argtypes = [vp, i32, i32, C.POINTER(PipelineCfg)] + [vp] * 9
    L.wavespec_pipeline_host.restype = i32
    L.wavespec_pipeline_device.argtypes = [vp, i32, i32, C.POINTER(PipelineCfg)] + [vp] * 10
    L.wavespec_pipeline_device.restype = i32
    L.wavespec_fft_real_forward_sliding.argtypes = [vp, i32, i32, i32, vp]
    L.wavespec_fft_real_forward_sliding.restype = i32
    L.wavespec_pla_windows_host.argtypes = [vp, i32, i32, i32, i32, dbl, vp, vp, vp]
    L.wavespec_pla_windows_host.restype = i32
    L.wavespec_zigzag_feed_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, dbl, i32, vp, vp]
    L.wavespec_zigzag_feed_host.restype = i32
    L.wavespec_applied_price_host.argtypes = [vp, vp, vp, vp, C.c_int64, i32, vp]
    L.wavespec_applied_price_host.restype = i32
    L.wavespec_applied_price_device.argtypes = [vp, vp, vp, vp, C.c_int64, i32, vp, vp]
    L.wavespec_applied_price_device.restype = i32
    L.wavespec_cycle_cache_host.argtypes = [vp, i32, i32, i32, i32, i32, i32, dbl, C.POINTER(CacheParams), vp]
    L.wavespec_cycle_cache_host.restype = i32
    L.wavespec_pipeline_host_planes.argtypes = [vp, i32, i32, C.POINTER(PipelineCfg), C.POINTER(Planes)]
    L.wavespec_pipeline_host_planes.restype = i32
    L.wavespec_pipeline_device_planes.argtypes = [vp, i32, i32, C.POINTER(PipelineCfg), C.POINTER(Planes), vp]
    L.wavespec_pipeline_device_planes.restype = i32
    L.wavespec_fft_real_inverse_batch_host.argtypes = [vp, i32, i32, vp]
    L.wavespec_fft_real_inverse_batch_host.restype = i32
    L.wavespec_fft_real_inverse_batch_device.argtypes = [vp, i32, i64, vp, vp]
    L.wavespec_fft_real_inverse_batch_device.restype = i32
    L.wavespec_reconstruct_topk_host.argtypes = [vp, vp, i32, i32, i32, vp]
    L.wavespec_reconstruct_topk_host.restype = i32
    L.wavespec_reconstruct_topk_device.argtypes = [vp, vp, i32, i32, i64, vp, vp]
    L.wavespec_reconstruct_topk_device.restype = i32
    L.wavespec_try_get_cycles_batch64.argtypes = [i64, vp, i64, C.POINTER(i64), _ip]
    L.wavespec_try_get_cycles_batch64.restype = i32
    L.wavespec_submit_cycle_cache_batch.argtypes = [vp, i32, i32, i32, i32, dbl, dbl, dbl, i32, i32,
                                                    C.POINTER(CacheParams), C.POINTER(i64)]
    L.wavespec_submit_cycle_cache_batch.restype = i32
    L.wavespec_try_get_cycle_cache.argtypes = [i64, vp, i64, _ip, _ip]
    L.wavespec_try_get_cycle_cache.restype = i32
    L.wavespec_device_count.argtypes = []; L.wavespec_device_count.restype = i32
    L.wavespec_job_device.argtypes = [i64]; L.wavespec_job_device.restype = i32
    L.wavespec_launch_count.argtypes = []; L.wavespec_launch_count.restype = i64
    L.wavespec_last_kernel.argtypes = []; L.wavespec_last_kernel.restype = C.c_char_p
    L.wavespec_version.argtypes = []; L.wavespec_version.restype = i32
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "gpu_init", "gpu_shutdown", "gpu_fft_real_forward", "gpu_extract_cycles", "gpu_submit_extract_cycles",
    "gpu_try_get_cycles", "gpu_submit_extract_cycles_batch", "gpu_try_get_cycles_batch", "gpu_free_job",
    "gpu_get_last_error_w", "gpu_fft_real_inverse", "gpu_fft_real_forward_batch",
    "wavespec_default_cfg", "wavespec_num_windows", "wavespec_pipeline_host", "wavespec_pipeline_device",
    "wavespec_fft_real_forward_sliding", "wavespec_pla_windows_host", "wavespec_zigzag_feed_host",
    "wavespec_cycle_cache_host", "wavespec_applied_price_host", "wavespec_applied_price_device",
    "wavespec_launch_count",
    "wavespec_last_kernel", "wavespec_version",
    "wavespec_pipeline_host_planes", "wavespec_pipeline_device_planes",
    "wavespec_fft_real_inverse_batch_host", "wavespec_fft_real_inverse_batch_device",
    "wavespec_try_get_cycles_batch64", "wavespec_submit_cycle_cache_batch", "wavespec_try_get_cycle_cache",
    "wavespec_device_count", "wavespec_job_device",
    "wavespec_reconstruct_topk_host", "wavespec_reconstruct_topk_device",
]


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def last_error() -> str:
    buf = (C.c_uint16 * 512)()
    n = lib().gpu_get_last_error_w(buf, 512)
    return "".join(chr(buf[i]) for i in range(max(0, n - 1)))


def _check(st):
    if st != OK:
        raise WaveSpecError(st, last_error())


# ---- the imports.mqh functions, same names ------------------------------------------------------
def gpu_init(device_index: int = 0, stream_count: int = 2) -> int:
    return lib().gpu_init(device_index, stream_count)


def gpu_shutdown() -> None:
    lib().gpu_shutdown()


def gpu_fft_real_forward(inp, out=None):
    x = _f64(inp)
    out = np.empty(x.size) if out is None else out
    _check(lib().gpu_fft_real_forward(_ptr(x), x.size, _ptr(out)))
    return out


def gpu_fft_real_inverse(spec):
    s = _f64(spec); out = np.empty(s.size)
    _check(lib().gpu_fft_real_inverse(_ptr(s), s.size, _ptr(out)))
    return out


def gpu_fft_real_forward_batch(inp, window_len, n_windows):
    x = _f64(inp); out = np.empty(window_len * n_windows)
    _check(lib().gpu_fft_real_forward_batch(_ptr(x), window_len, n_windows, _ptr(out)))
    return out.reshape(n_windows, window_len)


def fft_real_forward_sliding(series, window_len, hop=1):
    x = _f64(series)
    nw = lib().wavespec_num_windows(x.size, window_len, hop)
    out = np.empty((nw, window_len))
    _check(lib().wavespec_fft_real_forward_sliding(_ptr(x), x.size, window_len, hop, _ptr(out)))
    return out


def gpu_extract_cycles(series, top_k, min_period, max_period, sample_rate_seconds=60.0, method=0,
                       ar_order=10, out_stride=15, out_capacity=None):
    x = _f64(series)
    cap = top_k if out_capacity is None else out_capacity
    out = np.zeros((max(cap, 1), out_stride))
    n = C.c_int32(0)
    st = lib().gpu_extract_cycles(_ptr(x), x.size, top_k, min_period, max_period, sample_rate_seconds,
                                  method, ar_order, _ptr(out), out_stride, cap, C.byref(n))
    _check(st)
    return out[:n.value]


def gpu_submit_extract_cycles(series, top_k, min_period, max_period, sample_rate_seconds=60.0, method=0,
                              ar_order=10):
    x = _f64(series); jid = C.c_int64(0)
    st = lib().gpu_submit_extract_cycles(_ptr(x), x.size, top_k, min_period, max_period,
                                         sample_rate_seconds, method, ar_order, C.byref(jid))
    return st, jid.value


def gpu_try_get_cycles(job_id, out, out_stride, out_capacity):
    n = C.c_int32(0); ready = C.c_int32(0)
    st = lib().gpu_try_get_cycles(job_id, _ptr(out), out_stride, out_capacity, C.byref(n), C.byref(ready))
    return st, n.value, ready.value


def gpu_submit_extract_cycles_batch(series, window_len, hop, top_k, min_period, max_period,
                                    sample_rate_seconds=60.0, method=0, ar_order=10, stride=15):
    x = _f64(series); jid = C.c_int64(0)
    st = lib().gpu_submit_extract_cycles_batch(_ptr(x), x.size, window_len, hop, top_k, min_period,
                                               max_period, sample_rate_seconds, method, ar_order, stride,
                                               C.byref(jid))
    return st, jid.value


def gpu_try_get_cycles_batch(job_id, out, out_cap=None):
    n = C.c_int32(0); ready = C.c_int32(0)
    cap = out.size if out_cap is None else out_cap
    st = lib().gpu_try_get_cycles_batch(job_id, _ptr(out), cap, C.byref(n), C.byref(ready))
    return st, n.value, ready.value


def gpu_free_job(job_id) -> int:
    return lib().gpu_free_job(job_id)


def try_get_cycles_batch64(job_id, out):
    n = C.c_int64(0); ready = C.c_int32(0)
    st = lib().wavespec_try_get_cycles_batch64(job_id, _ptr(out), out.size, C.byref(n), C.byref(ready))
    return st, n.value, ready.value


def submit_cycle_cache_batch(series, window_len, hop, top_k, min_period, max_period, sample_rate_seconds=60.0,
                             method=0, ar_order=10, music_only=False, use_music_weights=False,
                             min_coherence=0.05, min_score=0.01, min_snr_db=-40.0):
    """Cycle cache record (20 doubles per bar, WaveSpecZZ_1.1.0-gpuopt.mq5:294-324) as a job product."""
    x = _f64(series); jid = C.c_int64(0)
    cp = CacheParams(int(music_only), int(use_music_weights), min_coherence, min_score, min_snr_db)
    st = lib().wavespec_submit_cycle_cache_batch(_ptr(x), x.size, window_len, hop, top_k, min_period, max_period,
                                                 sample_rate_seconds, method, ar_order, C.byref(cp), C.byref(jid))
    return st, jid.value


def try_get_cycle_cache(job_id, out):
    n = C.c_int32(0); ready = C.c_int32(0)
    st = lib().wavespec_try_get_cycle_cache(job_id, _ptr(out), out.size, C.byref(n), C.byref(ready))
    return st, n.value, ready.value


def device_count() -> int:
    return lib().wavespec_device_count()


def job_device(job_id) -> int:
    return lib().wavespec_job_device(job_id)


def fft_real_inverse_batch(spec, window_len):
    s = _f64(spec).reshape(-1, window_len)
    out = np.empty_like(s)
    _check(lib().wavespec_fft_real_inverse_batch_host(_ptr(s), window_len, s.shape[0], _ptr(out)))
    return out


# ---- new-build extensions -----------------------------------------------------------------------
def default_cfg(window_len, **over) -> PipelineCfg:
    cfg = PipelineCfg()
    lib().wavespec_default_cfg(C.byref(cfg), int(window_len))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def num_windows(series_len, window_len, hop=1) -> int:
    return 0 if series_len < window_len else 1 + (series_len - window_len) // hop


def pipeline_host(series, cfg: PipelineCfg, outputs=None):
    """Fused per-bar pipeline over host series ([n_series, series_len] or 1-D)."""
    s = _f64(series)
    squeeze = s.ndim == 1
    s2 = s.reshape(1, -1) if squeeze else s
    ns, sl = s2.shape
    outputs = cfg.outputs if outputs is None else outputs
    n, K = cfg.window_len, cfg.top_k
    nw = num_windows(sl, n, cfg.hop)
    o = {
        "spectra": np.empty((ns, nw, n)) if outputs & OUT_SPECTRA else None,
        "rows": np.empty((ns, nw, K, cfg.row_stride)) if outputs & OUT_ROWS else None,
        "bins": np.empty((ns, nw, K), dtype=np.int32) if outputs & OUT_BINS else None,
        "waves": np.empty((ns, nw, K)) if outputs & OUT_WAVES else None,
        "kalman": np.empty((ns, nw)) if outputs & OUT_KALMAN else None,
        "phase": np.empty((ns, nw, 3, n // 2)) if outputs & OUT_PHASE else None,
        "wkalman": np.empty((ns, nw)) if outputs & OUT_WKALMAN else None,
        "trk_index": np.empty((ns, nw, 12), dtype=np.int32) if outputs & OUT_TRACKER else None,
        "trk_period": np.empty((ns, nw, 12)) if outputs & OUT_TRACKER else None,
        "contrib": np.empty((ns, nw, K)) if outputs & OUT_CONTRIB else None,
    }
    pl = Planes(**{k: (None if v is None else v.ctypes.data) for k, v in o.items()})
    st = lib().wavespec_pipeline_host_planes(_ptr(s2), ns, sl, C.byref(cfg), C.byref(pl))
    _check(st)
    return {k: (v[0] if squeeze else v) for k, v in o.items() if v is not None}


def pipeline_device(d_series, n_series, series_len, cfg: PipelineCfg, spectra=0, rows=0, bins=0, waves=0,
                    kalman=0, phase=0, wkalman=0, trk_index=0, trk_period=0, contrib=0, stream=0):
    """Device-pointer pipeline; every buffer is a raw device address (e.g. torch.Tensor.data_ptr())."""
    vp = C.c_void_p
    pl = Planes(spectra=spectra or None, rows=rows or None, bins=bins or None, waves=waves or None,
                contrib=contrib or None, kalman=kalman or None, phase=phase or None, wkalman=wkalman or None,
                trk_index=trk_index or None, trk_period=trk_period or None)
    st = lib().wavespec_pipeline_device_planes(vp(d_series), n_series, series_len, C.byref(cfg), C.byref(pl),
                                               vp(stream or None))
    _check(st)


def pla_windows_host(series, window_len, hop=1, max_segments=32, max_error=0.0005):
    x = _f64(series)
    nw = num_windows(x.size, window_len, hop)
    cap = 2 * max(1, max_segments) + 2
    lines = np.empty((nw, window_len)); bounds = np.empty((nw, cap, 2), dtype=np.int32)
    counts = np.empty(nw, dtype=np.int32)
    _check(lib().wavespec_pla_windows_host(_ptr(x), x.size, window_len, hop, max_segments, max_error,
                                           _ptr(lines), _ptr(bounds), _ptr(counts)))
    return lines, bounds, counts


def zigzag_feed_host(zz_main, zz_high, zz_low, window_len, hop=1, pivot_rule=0, mode=0, fallback=0.0,
                     min_pivots=0):
    m, h, lo = _f64(zz_main), _f64(zz_high), _f64(zz_low)
    nw = num_windows(m.size, window_len, hop)
    lines = np.empty((nw, window_len)); valid = np.empty(nw, dtype=np.int32)
    _check(lib().wavespec_zigzag_feed_host(_ptr(m), _ptr(h), _ptr(lo), m.size, window_len, hop, pivot_rule, mode,
                                           float(fallback), min_pivots, _ptr(lines), _ptr(valid)))
    return lines, valid


PRICE_CLOSE, PRICE_OPEN, PRICE_HIGH, PRICE_LOW, PRICE_MEDIAN, PRICE_TYPICAL, PRICE_WEIGHTED = range(1, 8)


def applied_price_host(open_, high, low, close, mode):
    """A1: per-bar applied price series (mode = MQL5 ENUM_APPLIED_PRICE value); unused inputs may be None."""
    arrs = [None if a is None else _f64(a) for a in (open_, high, low, close)]
    sizes = {a.size for a in arrs if a is not None}
    if len(sizes) != 1:
        raise ValueError("price series must be present and equally long")
    n = sizes.pop()
    out = np.empty(n)
    _check(lib().wavespec_applied_price_host(*[_ptr(a) for a in arrs], n, int(mode), _ptr(out)))
    return out


def cycle_cache_host(rows, top_k, window_len, hop, bars, period_seconds=60.0, music_only=False,
                     use_music_weights=False, min_coherence=0.05, min_score=0.01, min_snr_db=-40.0):
    r = _f64(rows)
    stride = r.shape[-1]
    r2 = r.reshape(-1, stride)
    n_windows = r2.shape[0] // top_k
    cp = CacheParams(int(music_only), int(use_music_weights), min_coherence, min_score, min_snr_db)
    out = np.empty((bars, 20))
    _check(lib().wavespec_cycle_cache_host(_ptr(r2), n_windows, top_k, stride, window_len, hop, bars,
                                           float(period_seconds), C.byref(cp), _ptr(out)))
    return out


def reconstruct_topk(spectra, bins):
    """Inverse FFT of every window's spectrum masked to its selected bins: [nwin, N] samples."""
    s = _f64(spectra)
    n = s.shape[-1]
    s2 = s.reshape(-1, n)
    b = np.ascontiguousarray(bins, dtype=np.int32).reshape(s2.shape[0], -1)
    out = np.empty_like(s2)
    _check(lib().wavespec_reconstruct_topk_host(_ptr(s2), _ptr(b), n, b.shape[1], s2.shape[0], _ptr(out)))
    return out


def launch_count() -> int:
    return lib().wavespec_launch_count()


def last_kernel() -> str:
    return lib().wavespec_last_kernel().decode()
