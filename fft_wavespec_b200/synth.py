"""Seeded synthetic price series shared by the tests, the oracle runs and bench.py.

SURVEY.md section 8(d): close[s][t] = p0 + sigma * cumsum(g), g ~ N(0,1) from
PCG64(0x5EED0000 + s), p0 = 1.10000, sigma = 1e-4 (EURUSD-like M1), rounded to 5 decimals so
that genuinely repeated quotes (and therefore exact ties) occur.
"""
from __future__ import annotations

import numpy as np

P0 = 1.10000
SIGMA = 1.0e-4
SEED_BASE = 0x5EED0000


def random_walk(series_index: int, bars: int) -> np.ndarray:
    g = np.random.Generator(np.random.PCG64(SEED_BASE + int(series_index))).standard_normal(bars)
    return np.round(P0 + SIGMA * np.cumsum(g), 5)


def random_walk_batch(first_index: int, n_series: int, bars: int) -> np.ndarray:
    out = np.empty((n_series, bars), dtype=np.float64)
    for s in range(n_series):
        out[s] = random_walk(first_index + s, bars)
    return out


def high_low(series_index: int, close: np.ndarray):
    """High/low channels for the ZigZag / MID feeds: close +- |N(0,1)| * sigma / 2."""
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + 0x10000 + int(series_index)))
    up = np.abs(rng.standard_normal(close.size)) * SIGMA / 2
    dn = np.abs(rng.standard_normal(close.size)) * SIGMA / 2
    return np.round(close + up, 5), np.round(close - dn, 5)
