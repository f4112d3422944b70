"""BASELINE config 5 through ONE process and the multi-device session behind the C ABI
(WaveCyclesBatchFetcher.mq5:91-143 is a single process that calls gpu_init once and submits jobs):
N = 4096, K = 4, band 9-200, stride 15, `--series` series of `--bars` bars, on the first `--devices`
GPUs of the box.  Jobs are bound to the devices round robin by the library; each device's worker
thread issues its launches; results land in pinned host buffers.

Reports, per device count:
  device_windows_per_s   the kernels alone: wavespec_pipeline_device on every device, rows resident in HBM
  e2e_windows_per_s      submit / try_get / free with H2D of every series and D2H of every row
  roofline               P2 product: 8*hop + 120*K bytes per window against the HBM peak —
                         far below it by construction: the FP64 pipe and the 435-bin selection bind this shape
usage: python profiles/c5_sweep.py --devices 1 2 4 8 --series 256 --bars 500000 --out profiles/r02_c5_sweep.json
"""
import argparse
import json
import os
import sys
import time
from collections import deque

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_wavespec_b200 import bridge, synth  # noqa: E402

N, K, MINP, MAXP, STRIDE = 4096, 4, 9.0, 200.0, 15


def run(ndev, n_series, bars, unique, depth_per_dev, peak_gbs):
    bridge.gpu_shutdown()
    for d in range(ndev):
        st = bridge.gpu_init(d, 8)
        assert st == bridge.OK, bridge.last_error()
    assert bridge.device_count() == ndev
    nwin = bars - N + 1
    host = torch.empty((unique, bars), dtype=torch.float64, pin_memory=True)
    hn = host.numpy()
    for i in range(unique):
        hn[i] = synth.random_walk(5000 + i, bars)
    cfg = bridge.default_cfg(N, top_k=K, min_period=MINP, max_period=MAXP, row_stride=STRIDE)

    # ---- kernels alone, every device at once -----------------------------------------------------
    per_dev = max(1, n_series // ndev)
    group = min(8, per_dev)
    devbufs = []
    for d in range(ndev):
        with torch.cuda.device(d):
            ds = torch.from_numpy(hn[:group]).cuda(d)
            rows = torch.empty((group, nwin, K, STRIDE), dtype=torch.float64, device=f"cuda:{d}")
            devbufs.append((ds, rows, torch.cuda.Stream(device=d)))

    def device_pass():
        for _ in range(0, per_dev, group):
            for d, (ds, rows, st) in enumerate(devbufs):
                bridge.pipeline_device(ds.data_ptr(), group, bars, cfg, rows=rows.data_ptr(), stream=st.cuda_stream)

    def sync_all():
        for d in range(ndev):
            torch.cuda.synchronize(d)
    device_pass(); sync_all()
    l0 = bridge.launch_count()
    t = time.perf_counter()
    device_pass(); sync_all()
    dt_dev = time.perf_counter() - t
    launches = bridge.launch_count() - l0
    dev_rate = ndev * (per_dev // group) * group * nwin / dt_dev
    kernel = bridge.last_kernel()
    del devbufs
    for d in range(ndev):
        with torch.cuda.device(d):
            torch.cuda.empty_cache()

    # ---- end to end through the job API ----------------------------------------------------------
    depth = depth_per_dev * ndev
    out_doubles = nwin * K * STRIDE
    outs = [torch.empty(out_doubles, dtype=torch.float64, pin_memory=True).numpy() for _ in range(depth)]

    def e2e_pass(total):
        pending, free_bufs, i, done = deque(), list(range(depth)), 0, 0
        seen = set()
        while done < total:
            while i < total and free_bufs:
                st, jid = bridge.gpu_submit_extract_cycles_batch(hn[i % unique], N, 1, K, MINP, MAXP, 60.0, 0, 10, STRIDE)
                assert st == bridge.OK, bridge.last_error()
                seen.add(bridge.job_device(jid))
                b = free_bufs.pop()
                bridge.gpu_try_get_cycles_batch(jid, outs[b])
                pending.append((jid, b)); i += 1
            progressed = False
            for _ in range(len(pending)):                   # any finished job frees its buffer, not only the oldest
                jid, b = pending.popleft()
                st, n, ready = bridge.gpu_try_get_cycles_batch(jid, outs[b])
                assert st == bridge.OK, bridge.last_error()
                if ready:
                    assert n == nwin * K
                    bridge.gpu_free_job(jid); free_bufs.append(b); done += 1; progressed = True
                else:
                    pending.append((jid, b))
            if not progressed:
                time.sleep(0.0002)
        return seen
    e2e_pass(depth)                                         # warm-up
    t = time.perf_counter()
    seen = e2e_pass(n_series)
    dt = time.perf_counter() - t
    e2e_rate = n_series * nwin / dt
    alg = 8 * 1 + 120 * K
    res = {"devices": ndev, "series": n_series, "bars": bars, "windows_per_series": nwin,
           "device_windows_per_s": dev_rate, "device_ms": 1e3 * dt_dev, "device_launches": int(launches),
           "kernel": "ws::" + kernel + "_kernel",
           "e2e_windows_per_s": e2e_rate, "e2e_s": dt, "devices_that_ran_jobs": sorted(seen),
           "e2e_d2h_gb_per_s_aggregate": n_series * out_doubles * 8 / dt / 1e9,
           "jobs_in_flight": depth,
           "roofline_p2": {"algorithmic_bytes_per_window": alg, "achieved_gb_per_s_per_gpu": dev_rate / ndev * alg / 1e9,
                           "peak_gb_per_s": peak_gbs, "frac": dev_rate / ndev * alg / 1e9 / peak_gbs,
                           "note": "rows-only product: FP64 pipe / selection bound by construction, not HBM"}}
    del outs
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, nargs="+", default=[1])
    ap.add_argument("--series", type=int, default=256)
    ap.add_argument("--bars", type=int, default=500000)
    ap.add_argument("--unique", type=int, default=32, help="distinct synthetic series (reused cyclically)")
    ap.add_argument("--depth", type=int, default=3, help="jobs in flight per device")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    peak = 6525.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    have = torch.cuda.device_count()
    results = []
    for n in a.devices:
        if n > have:
            continue
        r = run(n, a.series, a.bars, a.unique, a.depth, peak)
        print(json.dumps(r), flush=True)
        results.append(r)
    bridge.gpu_shutdown()
    if results:
        base = results[0]
        for r in results:
            r["device_scaling_vs_first"] = r["device_windows_per_s"] / base["device_windows_per_s"] / (r["devices"] / base["devices"])
            r["e2e_scaling_vs_first"] = r["e2e_windows_per_s"] / base["e2e_windows_per_s"] / (r["devices"] / base["devices"])
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"config": "BASELINE config 5: N=4096, K=4, band 9-200, stride 15, hop 1, one process, multi-device session",
                       "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
