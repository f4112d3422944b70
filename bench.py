#!/usr/bin/env python
"""bench.py — sliding spectra per second (N=1024, FP64), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path over the batch: every hop-1 window of every series goes
through the shared-butterfly sliding FFT kernel, which writes the interleaved half spectrum
(the judged "P1" product, 8*hop + 8*N algorithmic bytes per spectrum) and the fused top-K cycle
rows.  Workload at one GPU: BASELINE config 2 — 64 synthetic random-walk series x 1M bars,
N=1024, top-8, band 18-200 (WaveSpecZZ_1.1.0-gpuopt / ...-gpuopt-nodetrend path).  Multi-GPU is
weak scaling: every rank owns its own 64 series (series shard with no collective on the data
path); torch.distributed/NCCL only carries the barrier and the max-over-ranks time.

  value        device-resident inputs, CUDA-event time over exactly K steps, max over ranks
  e2e          the same metric through the imports.mqh API (gpu_submit_extract_cycles_batch /
               gpu_try_get_cycles_batch / gpu_free_job) with pinned HOST buffers: H2D of every
               series and D2H of every result row inside the timed region
  roofline     algorithmic HBM bytes of the dominant kernel / its measured duration, against
               MEASURED_PEAKS.json
  cpu_baseline the oracle (line-faithful port of the reference's CPU path) on the host cores,
               bounded sample; --impl reference times the same thing as its own arm
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sliding spectra/sec (N=1024, FP64)"
UNIT = "spectra/s"
N_WINDOW = 1024
TOP_K = 8
MIN_P, MAX_P = 18.0, 200.0
ROW_STRIDE = 15


def config_dict(n_series, bars, gpus):
    return {"workload": f"config 2: {n_series} synthetic random-walk series x {bars} bars per GPU, N={N_WINDOW} hop=1 "
                        f"sliding real FFT (FP64) + top-{TOP_K} cycles, band {MIN_P:g}-{MAX_P:g}, no detrend/window",
            "series_per_gpu": n_series, "bars": bars, "window": N_WINDOW, "hop": 1, "top_k": TOP_K,
            "row_stride": ROW_STRIDE, "parallelism": f"series-sharded x{gpus}, no collective",
            "outputs": "interleaved half spectra (8 KiB/window) + cycle rows, written to HBM",
            "l2": "per-launch output (>= 8 GB) and the spectra ring exceed the 126 MB L2; no flush needed"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.samples if t0 <= t <= t1 + 0.1] or [l for (_, l) in self.samples[-3:]]
        for line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def cpu_oracle_rate(n_series, target_seconds, threads):
    """Times the oracle (port of the reference CPU path) on a bounded sample of the workload."""
    from fft_wavespec_b200 import synth
    from oracle import oracle as orc
    cfg = orc.default_cfg(N_WINDOW, top_k=TOP_K, min_period=MIN_P, max_period=MAX_P, row_stride=ROW_STRIDE)
    ns = max(1, min(n_series, threads))
    cal = 4000
    series = synth.random_walk_batch(0, ns, N_WINDOW - 1 + cal)
    t = time.perf_counter()
    orc.pipeline_batch_mt(series, cfg, threads, want=("bins",))
    rate = ns * cal / (time.perf_counter() - t)
    per_series = int(max(cal, min(900000, rate * target_seconds / ns)))
    series = synth.random_walk_batch(0, ns, N_WINDOW - 1 + per_series)
    t = time.perf_counter()
    done, _ = orc.pipeline_batch_mt(series, cfg, threads, want=("bins",))
    dt = time.perf_counter() - t
    return done / dt, f"first {per_series} windows of {ns} series of config 2 ({done} spectra, {dt:.1f} s)", (series, cfg)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; MQL5 cannot be
    compiled or run here) on all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    _, _, (series, cfg) = cpu_oracle_rate(64, 2.0, threads)      # sizes one step at ~2 s of CPU work
    per_step = series.shape[0] * (series.shape[1] - N_WINDOW + 1)
    for _ in range(args.warmup):
        orc.pipeline_batch_mt(series, cfg, threads, want=("bins",))
    t = time.perf_counter()
    for _ in range(args.steps):
        orc.pipeline_batch_mt(series, cfg, threads, want=("bins",))
    dt = time.perf_counter() - t
    value = per_step * args.steps / dt
    sample = f"each step = first {series.shape[1] - N_WINDOW + 1} windows of {series.shape[0]} series of config 2"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(64, 1000000, args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit_json(out)


# --------------------------------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Native libraries write there too (NCCL prints its
    version banner to fd 1 whenever NCCL_DEBUG is set on the box), so the real stdout is kept aside
    for the JSON line and fd 1 is pointed at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json(obj):
    print(json.dumps(obj), file=_JSON_OUT if _JSON_OUT is not None else sys.stdout, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series", type=int, default=64, help="series per GPU (config 2: 64)")
    ap.add_argument("--bars", type=int, default=1000000)
    ap.add_argument("--group", type=int, default=4, help="series per kernel launch (spectra ring size)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fft_wavespec_b200 import bridge, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    st = bridge.gpu_init(local_rank, 8)
    if st != bridge.OK:
        raise SystemExit(f"gpu_init failed: {st} {bridge.last_error()}")

    S, T, G = args.series, args.bars, args.group
    nwin = T - N_WINDOW + 1
    cfg = bridge.default_cfg(N_WINDOW, top_k=TOP_K, min_period=MIN_P, max_period=MAX_P, row_stride=ROW_STRIDE)

    # ---- inputs resident in HBM before the timed region ------------------------------------------
    host = torch.empty((S, T), dtype=torch.float64, pin_memory=True)
    host_np = host.numpy()
    for i in range(S):
        host_np[i] = synth.random_walk(rank * S + i, T)
    d_series = host.cuda(non_blocking=False)
    d_spectra = torch.empty((G, nwin, N_WINDOW), dtype=torch.float64, device="cuda")        # ring, overwritten per group
    d_rows = torch.empty((S, nwin, TOP_K, ROW_STRIDE), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step(with_spectra=True):
        for g0 in range(0, S, G):
            g = min(G, S - g0)
            bridge.pipeline_device(d_series[g0].data_ptr(), g, T, cfg,
                                   spectra=d_spectra.data_ptr() if with_spectra else 0,
                                   rows=d_rows[g0].data_ptr(), stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = bridge.launch_count()
        e0.record(stream)
        for _ in range(nsteps):
            step(**kw)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, bridge.launch_count() - l0

    for _ in range(args.warmup):
        step()
    assert bridge.last_kernel() in ("sliding_shared", "sliding_overlap"), bridge.last_kernel()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t0 = time.time()
    ms, launches = timed(args.steps)
    t1 = time.time()
    main_kernel = bridge.last_kernel()                      # the kernel of the timed region
    clocks = sampler.stop(t0, t1)
    spectra_per_step = S * nwin
    value = world * spectra_per_step * args.steps / (ms * 1e-3)

    # rows-only product (P2) for reference: same kernel, spectra store disabled
    for _ in range(2):
        step(with_spectra=False)
    ms_rows, _ = timed(max(2, args.steps // 2), with_spectra=False)
    rows_only = world * spectra_per_step * max(2, args.steps // 2) / (ms_rows * 1e-3)

    peak, peak_src = load_peaks()
    alg_bytes = 8 * 1 + 8 * N_WINDOW                       # 8*hop in + 8*N out per spectrum
    per_gpu = value / world
    achieved = per_gpu * alg_bytes / 1e9
    launch_ms = ms / launches
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "ws::" + main_kernel + "_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_spectrum": alg_bytes, "spectra_per_launch": G * nwin,
                "avg_launch_ms": launch_ms,
                "note": "rows (960 B/window) are extra traffic not counted in the algorithmic bytes"}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: the imports.mqh job API with pinned host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        out_doubles = nwin * TOP_K * ROW_STRIDE
        outs = [torch.empty(out_doubles, dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)]
        depth = 4

        def e2e_step():
            pending, got_rows, k = [], 0, 0
            def drain():
                nonlocal got_rows, k
                jid = pending.pop(0)
                while True:
                    stt, n, ready = bridge.gpu_try_get_cycles_batch(jid, outs[k % 2], out_doubles)
                    if stt == bridge.OK and ready:
                        break
                    if stt != bridge.NOT_READY:
                        raise RuntimeError(f"try_get failed {stt}: {bridge.last_error()}")
                    time.sleep(0.0002)
                bridge.gpu_free_job(jid)
                got_rows += n; k += 1
            for i in range(S):
                stt, jid = bridge.gpu_submit_extract_cycles_batch(host_np[i], N_WINDOW, 1, TOP_K, MIN_P, MAX_P, 60.0,
                                                                  0, 10, ROW_STRIDE)
                if stt != bridge.OK:
                    raise RuntimeError(f"submit failed {stt}: {bridge.last_error()}")
                pending.append(jid)
                if len(pending) >= depth:
                    drain()
            while pending:
                drain()
            assert got_rows == S * nwin * TOP_K
        e2e_step()                                         # warm-up
        barrier()
        tt = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - tt
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * spectra_per_step * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": world * S * T * 8, "d2h_bytes_per_step": world * S * out_doubles * 8,
               "api": "gpu_submit_extract_cycles_batch / gpu_try_get_cycles_batch / gpu_free_job, pinned host buffers",
               "ms_per_step": 1e3 * dt / args.e2e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sample, _ = cpu_oracle_rate(S, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(S, T, world),
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
               "clocks": clocks, "extra": {"rows_only_spectra_per_s": rows_only,
                                           "rows_only_kernel": "ws::" + bridge.last_kernel() + "_kernel"}}
        emit_json(out)
    bridge.gpu_shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
