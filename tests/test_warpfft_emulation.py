"""CPU emulation of the warp-per-window FFT kernel: the same host/device arithmetic the CUDA kernel
runs (fft_wavespec_b200/csrc/ws_warpfft_core.cuh), executed lane by lane and phase by phase on the
host, compared with the oracle's FFT.  Covers the in-place DIF pass geometry, the digit-reversed
real-input split and the shared-memory swizzle (bank-conflict freedom) without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fft_wavespec_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu_warpfft.cpp")
    so = os.path.join(HERE, "emu", "libemu_warpfft.so")
    core = os.path.join(os.path.dirname(HERE), "fft_wavespec_b200", "csrc", "ws_warpfft_core.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(so):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.emu_warpfft.argtypes = [dp, C.c_int, dp]
    L.emu_warpfft.restype = C.c_int
    L.emu_warp_ifft.argtypes = [dp, C.c_int, C.c_void_p, dp]
    L.emu_warp_ifft.restype = C.c_int
    L.emu_warpfft_conflicts.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.emu_warpfft_conflicts.restype = C.c_int
    return L


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_emulated_warp_fft_matches_oracle(emu, oracle, n):
    for seed in (0, 1):
        x = synth.random_walk(700 + seed, n)
        out = np.full(n, np.nan)
        # rc -2: a bin was produced twice or never; -3: split_bin disagrees with split_phase
        assert emu.emu_warpfft(x, n, out) == 0
        ref = oracle.fft_interleaved(x)
        assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-13
        assert np.abs(out[2:] - ref[2:]).max() / np.abs(ref[2:]).max() < 1e-12
        assert out[1] == 0.0


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_swizzle_keeps_quarter_warps_conflict_free(emu, n):
    g = C.c_int(0)
    worst = emu.emu_warpfft_conflicts(n, C.byref(g))
    assert worst == 1, "a pass access has two lanes of a quarter warp in one 16-byte bank group"
    # the mirrored side of the gather (bin M-k) breaks the digit pattern at k = 0 mod 8 only
    assert g.value <= 2


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_emulated_warp_inverse_matches_oracle(emu, oracle, n):
    """The inverse kernel's arithmetic (ws_inverse.cu) against oracle_fft_inverse, and as the inverse
    of the forward contract: inverse(forward(x)) = x minus the dropped Nyquist component."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    spec = oracle.fft_interleaved(x)
    out = np.full(n, np.nan)
    assert emu.emu_warp_ifft(spec, n, None, out) == 0
    ref = oracle.fft_inverse(spec)
    assert np.abs(out - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    nyq = np.sum(x * (-1.0) ** np.arange(n))
    assert np.abs(out - (x - nyq * (-1.0) ** np.arange(n) / n)).max() < 1e-12
    if n >= 64:
        # spectrum masked to a few bins (top-K reconstruction): sum of those cycles over the window
        keep = np.zeros(n // 2, dtype=np.uint8)
        sel = rng.choice(np.arange(1, n // 2), size=5, replace=False)
        keep[sel] = 1
        assert emu.emu_warp_ifft(spec, n, keep.ctypes.data, out) == 0
        masked = spec.copy().reshape(-1, 2)
        masked[keep == 0] = 0.0
        ref = oracle.fft_inverse(masked.reshape(-1))
        assert np.abs(out - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
        t = np.arange(n)
        direct = sum((2.0 / n) * (spec[2 * k] * np.cos(2 * np.pi * k * t / n) - spec[2 * k + 1] * np.sin(2 * np.pi * k * t / n))
                     for k in sel)
        assert np.abs(out - direct).max() < 1e-11 * max(1.0, np.abs(direct).max())
