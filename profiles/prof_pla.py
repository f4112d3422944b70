"""Driver for the PLA feed (A11): 4 series x 50k bars, N=1024, PLA line -> FFT -> top-8 bins.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge as br, synth  # noqa: E402

assert br.gpu_init(0, 2) == 0
n, S, T = 1024, 4, 50000
cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=200.0, feed=br.FEED_PLA)
nwin = T - n + 1
d = torch.from_numpy(synth.random_walk_batch(0, S, T)).cuda()
bins = torch.empty((S, nwin, 8), dtype=torch.int32, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    br.pipeline_device(d.data_ptr(), S, T, cfg, bins=bins.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
print(f"PLA feed + FFT N={n} {S} x {T}: {e0.elapsed_time(e1):.3f} ms {S * nwin / e0.elapsed_time(e1) / 1e3:.2f} M windows/s {br.last_kernel()}")
br.gpu_shutdown()
