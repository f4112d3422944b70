"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): one short call through every
kernel family, including the round-2 ones (staged sliding kernel, warp tracker, lockstep PLA, batched
inverse, cycle-cache job).  Usage: compute-sanitizer --tool racecheck python profiles/sanitize.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge as br, synth  # noqa: E402

assert br.gpu_init(0, 4) == 0
s = synth.random_walk_batch(0, 2, 1024 + 150)
# sliding kernels: producer/consumer (spectra + rows), rows only, spectra only, N = 512 and 4096
for n in (1024, 512):
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0)
    br.pipeline_host(s, cfg, br.OUT_SPECTRA | br.OUT_ROWS | br.OUT_BINS)
    print(n, br.last_kernel())
    br.pipeline_host(s, cfg, br.OUT_ROWS)
    br.pipeline_host(s, cfg, br.OUT_SPECTRA)
br.pipeline_host(synth.random_walk(3, 4096 + 40), br.default_cfg(4096, top_k=4, min_period=9.0, max_period=200.0), br.OUT_ROWS)
# staged variant (opt-in) in a child process would need its own env: run it here when asked to
if os.environ.get("WAVESPEC_STAGED") == "1":
    print("staged:", br.last_kernel())
# per-window kernels: warp (N = 1024, 4096), CTA (N = 256), phase chain, PLA feed, tracker, Kalman
for n in (1024, 4096, 256):
    cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR, trend_period=float(n // 2),
                         window_type=br.WINDOW_BLACKMAN)
    br.pipeline_host(synth.random_walk(4, n + 70), cfg, br.OUT_SPECTRA | br.OUT_BINS | br.OUT_ROWS | br.OUT_KALMAN)
    print(n, br.last_kernel())
cfg = br.default_cfg(1024, top_k=8, min_period=18.0, max_period=200.0)
br.pipeline_host(s[0], cfg, br.OUT_PHASE | br.OUT_ROWS)
br.pipeline_host(s[0][:1024 + 40], br.default_cfg(1024, feed=br.FEED_PLA), br.OUT_BINS | br.OUT_KALMAN)
br.pipeline_host(s, br.default_cfg(1024, min_period=18.0, max_period=52.0), br.OUT_TRACKER | br.OUT_BINS)
br.pipeline_host(s[0], br.default_cfg(1024, window_type=br.WINDOW_HANN, select=br.SELECT_SORT), br.OUT_WKALMAN | br.OUT_CONTRIB)
# inverse / reconstruction
got = br.pipeline_host(s[0], cfg, br.OUT_SPECTRA | br.OUT_BINS)
br.fft_real_inverse_batch(got["spectra"], 1024)
br.reconstruct_topk(got["spectra"], got["bins"])
# job table: rows, cache record, window jobs
os.environ["WAVESPEC_JOB_CHUNK"] = "64"
out = np.empty(151 * 8 * 15); rec = np.empty(1174 * 20)
st, j1 = br.gpu_submit_extract_cycles_batch(s[0], 1024, 1, 8, 18.0, 200.0)
st, j2 = br.submit_cycle_cache_batch(s[1], 1024, 1, 2, 18.0, 200.0)
st, j3 = br.gpu_submit_extract_cycles(s[0][:1024], 2, 9.0, 200.0)
buf = np.zeros((2, 15))
for _ in range(20000):
    a = br.gpu_try_get_cycles_batch(j1, out); b = br.try_get_cycle_cache(j2, rec); c = br.gpu_try_get_cycles(j3, buf, 15, 2)
    if a[2] and b[2] and c[2]:
        break
    time.sleep(0.001)
assert a[2] and b[2] and c[2]
for j in (j1, j2, j3):
    br.gpu_free_job(j)
br.gpu_shutdown()
print("sanitize driver done")
