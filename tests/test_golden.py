"""Committed golden vectors (tests/golden/, produced by tests/golden/make_golden.py from the
oracle): the oracle must still reproduce them (CPU), and the CUDA path must match them through the
C ABI (GPU).  Integer planes exactly, floating planes to 1e-9 relative."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "wavespec_golden.npz"))


def test_oracle_reproduces_golden_vectors():
    fresh = make_golden.build()
    assert set(fresh) == set(G.files)
    for k in G.files:
        a, b = fresh[k], G[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == "i":
            assert np.array_equal(a, b), k
        else:
            den = max(1e-300, np.abs(b).max())
            assert np.abs(a - b).max() / den < 1e-12, k


def test_golden_sanity_known_structure():
    # period = N / bin in every stored row; bins inside the configured band
    rows, bins = G["c2_rows"], G["c2_bins"]
    assert np.array_equal(rows[..., 2], 1024.0 / bins)
    assert bins.min() >= 6 and bins.max() <= 56
    assert G["c5_bins"].min() >= 21 and G["c5_bins"].max() <= 455
    assert G["c3_bins"].min() >= 40 and G["c3_bins"].max() <= 113


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_cuda_matches_golden(name):
    import ctypes as C
    from fft_wavespec_b200 import bridge as br
    assert br.gpu_init(0, 4) == br.OK, br.last_error()
    sidx, bars, over, outputs, keep = make_golden.CASES[name]
    cfg = br.default_cfg(over["window_len"])
    for k, v in over.items():
        setattr(cfg, k, v)
    got = br.pipeline_host(G[f"{name}_series"], cfg, outputs)
    for plane in ("bins",):
        assert np.array_equal(got[plane], G[f"{name}_{plane}"]), plane
    sp = got["spectra"][list(keep)]
    ref = G[f"{name}_spectra"]
    assert (np.abs(sp - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-9
    if f"{name}_kalman" in G.files:
        assert np.array_equal(got["kalman"], G[f"{name}_kalman"])
    if f"{name}_trk_index" in G.files:
        assert np.array_equal(got["trk_index"], G[f"{name}_trk_index"])
        assert np.array_equal(got["trk_period"], G[f"{name}_trk_period"])
    if f"{name}_waves" in G.files:
        w = G[f"{name}_waves"]
        assert np.abs(got["waves"] - w).max() <= 1e-9 * np.abs(w).max()
    if f"{name}_wkalman" in G.files:
        w = G[f"{name}_wkalman"]
        assert np.abs(got["wkalman"] - w).max() <= 1e-9 * max(1.0, np.abs(w).max())
    if f"{name}_rows" in G.files:
        r = G[f"{name}_rows"]
        assert np.abs(got["rows"][..., 0] - r[..., 0]).max() <= 1e-9 * r[..., 0].max()
        assert np.array_equal(got["rows"][..., 2], r[..., 2])


@pytest.mark.gpu
def test_cuda_pla_and_kalman_match_golden():
    from fft_wavespec_b200 import bridge as br
    assert br.gpu_init(0, 4) == br.OK
    lines, bounds, counts = br.pla_windows_host(G["pla_series"], 512, 1, 32, 0.0005)
    for w in (0, 15, 30):
        n = G[f"pla_start_{w}"].size
        assert counts[w] == n
        assert np.array_equal(bounds[w, :n, 0], G[f"pla_start_{w}"])
        assert np.array_equal(bounds[w, :n, 1], G[f"pla_end_{w}"])
        assert np.array_equal(lines[w], G[f"pla_line_{w}"])
    z = G["kalman_in"]
    cfg = br.default_cfg(64)
    got = br.pipeline_host(np.concatenate([np.zeros(63), z]), cfg, br.OUT_KALMAN)
    assert np.array_equal(got["kalman"], G["kalman_out"])
