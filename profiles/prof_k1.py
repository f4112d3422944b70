"""ncu driver for the per-window FFT kernel (config 3 shape by default)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge as br, synth
assert br.gpu_init(0, 2) == 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
S, T = 4, 60000
cfg = br.default_cfg(n, top_k=8, min_period=18.0, max_period=52.0, detrend=br.DETREND_IIR, trend_period=1024.0,
                     window_type=br.WINDOW_BLACKMAN)
nwin = T - n + 1
d = torch.from_numpy(synth.random_walk_batch(0, S, T)).cuda()
spec = torch.empty((S, nwin, n), dtype=torch.float64, device="cuda")
bins = torch.empty((S, nwin, 8), dtype=torch.int32, device="cuda")
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    br.pipeline_device(d.data_ptr(), S, T, cfg, spectra=spec.data_ptr(), bins=bins.data_ptr(),
                       stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
print(f"N={n} {e0.elapsed_time(e1):.3f} ms {S*nwin/e0.elapsed_time(e1)/1e3:.2f} M windows/s {br.last_kernel()}")
