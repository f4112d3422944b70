"""Times the five BASELINE configs (reduced series counts, full window lengths) through the device
pipeline: which kernel serves them and at what rate.  Usage: python profiles/prof_configs.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge as br, synth  # noqa: E402

assert br.gpu_init(0, 2) == 0
st = torch.cuda.current_stream().cuda_stream


def run(name, n, series, bars, outputs, **over):
    cfg = br.default_cfg(n, **over)
    K = cfg.top_k
    nwin = bars - n + 1
    d = torch.from_numpy(synth.random_walk_batch(0, series, bars)).cuda()
    bufs = {}
    if outputs & br.OUT_SPECTRA: bufs["spectra"] = torch.empty((series, nwin, n), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_ROWS: bufs["rows"] = torch.empty((series, nwin, K, 15), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_BINS: bufs["bins"] = torch.empty((series, nwin, K), dtype=torch.int32, device="cuda")
    if outputs & br.OUT_WAVES: bufs["waves"] = torch.empty((series, nwin, K), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_KALMAN: bufs["kalman"] = torch.empty((series, nwin), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_PHASE: bufs["phase"] = torch.empty((series, nwin, 3, n // 2), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_WKALMAN: bufs["wkalman"] = torch.empty((series, nwin), dtype=torch.float64, device="cuda")
    if outputs & br.OUT_TRACKER:
        bufs["trk_index"] = torch.empty((series, nwin, 12), dtype=torch.int32, device="cuda")
        bufs["trk_period"] = torch.empty((series, nwin, 12), dtype=torch.float64, device="cuda")
    ptrs = {k: v.data_ptr() for k, v in bufs.items()}
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        br.pipeline_device(d.data_ptr(), series, bars, cfg, stream=st, **ptrs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"{name:34s} N={n:5d} {series:3d} x {bars:7d}  {ms:9.3f} ms  {series * nwin / ms / 1e3:9.2f} M windows/s  "
          f"kernel={br.last_kernel()}", flush=True)


S, R, B, W, KA, PH, WK = br.OUT_SPECTRA, br.OUT_ROWS, br.OUT_BINS, br.OUT_WAVES, br.OUT_KALMAN, br.OUT_PHASE, br.OUT_WKALMAN
run("C1 mean+Hann top-5", 512, 1, 100000, S | R | B, top_k=5, min_period=9.0, max_period=200.0,
    detrend=br.DETREND_MEAN, window_type=br.WINDOW_HANN_WIP)
run("C1 (8 series)", 512, 8, 100000, S | R | B, top_k=5, min_period=9.0, max_period=200.0,
    detrend=br.DETREND_MEAN, window_type=br.WINDOW_HANN_WIP)
run("C2 plain top-8", 1024, 4, 300000, S | R, top_k=8, min_period=18.0, max_period=200.0)
run("C3 IIR+Blackman (spectra+bins)", 2048, 4, 100000, S | B, top_k=8, min_period=18.0, max_period=52.0,
    detrend=br.DETREND_IIR, trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("C3 Kalman4D only", 2048, 256, 1000000, KA)
run("C3 bins + Kalman4D (side stream)", 2048, 64, 200000, B | KA, top_k=8, min_period=18.0, max_period=52.0,
    detrend=br.DETREND_IIR, trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("C3 bins only (same shape)", 2048, 64, 200000, B, top_k=8, min_period=18.0, max_period=52.0,
    detrend=br.DETREND_IIR, trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("C4 Hann+sort+waves+wkalman", 1024, 4, 100000, S | B | W | WK, top_k=8, min_period=12.0, max_period=256.0,
    window_type=br.WINDOW_HANN, select=br.SELECT_SORT)
run("C4 phase chain (Hann)", 1024, 2, 50000, PH, window_type=br.WINDOW_HANN)
run("C4 plain: phase+waves+rows", 1024, 4, 100000, PH | W | R, top_k=8, min_period=18.0, max_period=200.0)
run("C4 plain: spectra+phase+rows", 1024, 4, 100000, S | PH | R, top_k=8, min_period=18.0, max_period=200.0)
run("C5 N=4096 K=4 rows", 4096, 2, 200000, R, top_k=4, min_period=9.0, max_period=200.0)
run("C5 N=4096 K=4 spectra+rows", 4096, 2, 200000, S | R, top_k=4, min_period=9.0, max_period=200.0)
run("N=1024 IIR+Blackman (spectra+bins)", 1024, 4, 100000, S | B, top_k=8, min_period=18.0, max_period=52.0,
    detrend=br.DETREND_IIR, trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("N=4096 IIR+Blackman (spectra+bins)", 4096, 2, 100000, S | B, top_k=8, min_period=18.0, max_period=52.0,
    detrend=br.DETREND_IIR, trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("N=4096 Hann K=4 band 9-200 (rows)", 4096, 2, 100000, R, top_k=4, min_period=9.0, max_period=200.0,
    window_type=br.WINDOW_HANN)
run("PLA feed + FFT", 1024, 1, 30000, S | B, feed=br.FEED_PLA)
run("PLA feed + FFT (4 x 100k)", 1024, 4, 100000, S | B, feed=br.FEED_PLA)
TR = br.OUT_TRACKER
run("C3 tracker slots only", 2048, 16, 200000, TR, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR,
    trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("C3 tracker + bins", 2048, 4, 100000, TR | B, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR,
    trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
run("C3 tracker + bins (64 x 200k)", 2048, 64, 200000, TR | B, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR,
    trend_period=1024.0, window_type=br.WINDOW_BLACKMAN)
br.gpu_shutdown()
