"""Generates tests/golden/wavespec_golden.npz from the CPU oracle (oracle/wavespec_oracle.cpp).

The reference ships no golden vectors, so these are OUR pins: they freeze
the oracle's answers on small instances of the five BASELINE configs so that neither the oracle
nor the CUDA path can drift silently.  Run from the repo root:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fft_wavespec_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

# name -> (series index, bars, cfg overrides, outputs, windows whose spectra are stored)
CASES = {
    "c1": (0, 512 + 190, dict(window_len=512, top_k=5, min_period=9.0, max_period=200.0, detrend=2, window_type=5),
           orc.OUT_SPECTRA | orc.OUT_BINS | orc.OUT_ROWS | orc.OUT_WAVES, (0, 95, 190)),
    "c2": (1, 1024 + 76, dict(window_len=1024, top_k=8, min_period=18.0, max_period=200.0),
           orc.OUT_SPECTRA | orc.OUT_BINS | orc.OUT_ROWS | orc.OUT_WAVES, (0, 38, 76)),
    "c3": (2, 2048 + 40, dict(window_len=2048, top_k=8, min_period=18.0, max_period=52.0, detrend=1,
                              trend_period=1024.0, window_type=3),
           orc.OUT_SPECTRA | orc.OUT_BINS | orc.OUT_KALMAN | orc.OUT_TRACKER, (0, 40)),
    "c4": (3, 1024 + 60, dict(window_len=1024, top_k=8, min_period=12.0, max_period=256.0, window_type=1, select=1),
           orc.OUT_SPECTRA | orc.OUT_BINS | orc.OUT_WAVES | orc.OUT_WKALMAN | orc.OUT_PHASE, (0, 60)),
    "c5": (4, 4096 + 12, dict(window_len=4096, top_k=4, min_period=9.0, max_period=200.0),
           orc.OUT_SPECTRA | orc.OUT_BINS | orc.OUT_ROWS, (0, 12)),
}


def case_cfg(over):
    cfg = orc.default_cfg(over["window_len"])
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def build():
    out = {}
    for name, (sidx, bars, over, outputs, keep) in CASES.items():
        s = synth.random_walk(sidx, bars)
        r = orc.pipeline_series(s, case_cfg(over), outputs)
        out[f"{name}_series"] = s
        for k, v in r.items():
            if k in ("spectra", "phase"):
                out[f"{name}_{k}"] = v[list(keep)]
            else:
                out[f"{name}_{k}"] = v
    s = synth.random_walk(5, 512 + 30)
    for w in (0, 15, 30):
        line, st, en, _, _ = orc.pla_build(s[w:w + 512], 32, 0.0005)
        out[f"pla_line_{w}"] = line; out[f"pla_start_{w}"] = st; out[f"pla_end_{w}"] = en
    out["pla_series"] = s
    z = synth.random_walk(6, 3000)
    out["kalman_in"] = z
    out["kalman_out"] = orc.kalman4d_series(z)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wavespec_golden.npz")
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path), "bytes")
