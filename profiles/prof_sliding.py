"""Small driver for ncu: launches the sliding kernel in three output modes (2 launches each):
spectra+rows, spectra only, rows only.  Usage: python profiles/prof_sliding.py [N] [series] [bars] [top_k] [min_period]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300000
assert bridge.gpu_init(0, 2) == 0
K = int(sys.argv[4]) if len(sys.argv) > 4 else 8
MINP = float(sys.argv[5]) if len(sys.argv) > 5 else 18.0
cfg = bridge.default_cfg(N, top_k=K, min_period=MINP, max_period=200.0)
nwin = T - N + 1
d = torch.from_numpy(synth.random_walk_batch(0, S, T)).cuda()
spec = torch.empty((S, nwin, N), dtype=torch.float64, device="cuda")
rows = torch.empty((S, nwin, K, 15), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for mode in ("both", "spectra", "rows"):
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bridge.pipeline_device(d.data_ptr(), S, T, cfg, spectra=spec.data_ptr() if mode != "rows" else 0,
                               rows=rows.data_ptr() if mode != "spectra" else 0, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"N={N} mode={mode:8s} {ms:8.3f} ms  {S * nwin / ms / 1e3:8.2f} M spectra/s  kernel={bridge.last_kernel()}")
bridge.gpu_shutdown()
