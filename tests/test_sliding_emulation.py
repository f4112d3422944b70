"""CPU emulation of the shared-butterfly sliding FFT kernel: the same host/device arithmetic the
CUDA kernel runs (fft_wavespec_b200/csrc/ws_sliding_core.cuh), executed thread by thread and pass
by pass on the host, compared with the oracle's per-window FFT.  Covers the tile/halo geometry,
the packed real-DFT butterflies and ragged last tiles without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fft_wavespec_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu_sliding.cpp")
    so = os.path.join(HERE, "emu", "libemu_sliding.so")
    core = os.path.join(os.path.dirname(HERE), "fft_wavespec_b200", "csrc", "ws_sliding_core.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(so):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.emu_sliding.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.emu_sliding.restype = C.c_int
    L.emu_sliding_top.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.emu_sliding_top.restype = C.c_int
    L.emu_sliding_staged.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.emu_sliding_staged.restype = C.c_int
    return L


# (N, T, S) as launched by ws_sliding.cu::pick_plan, plus off-nominal tilings
PLANS = [(256, 128, 16), (512, 64, 8), (1024, 32, 4), (2048, 16, 2), (4096, 8, 1),
         (1024, 16, 1), (1024, 64, 8), (512, 24, 3), (1024, 8, 2),
         (1024, 32, 2), (1024, 32, 1), (512, 64, 4), (2048, 16, 1), (256, 128, 4), (1024, 40, 2)]   # shapes of the producer/consumer kernel


@pytest.mark.parametrize("n,t,s", PLANS)
def test_emulated_kernel_matches_oracle(emu, oracle, n, t, s):
    for extra in (0, t - 1, 2 * t + 3):
        x = synth.random_walk(300 + n, n + extra)
        nw = x.size - n + 1
        out = np.full((nw, n), np.nan)
        assert emu.emu_sliding(x, x.size, n, t, s, 256, out) == 0
        assert not np.isnan(out).any(), "a bin of some window was never produced"
        ref = np.stack([oracle.fft_interleaved(x[w:w + n]) for w in range(nw)])
        err = np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert err.max() < 1e-12
        off_dc = np.abs(out[:, 2:] - ref[:, 2:]).max(axis=1) / np.abs(ref[:, 2:]).max(axis=1)
        assert off_dc.max() < 1e-12
        assert np.all(out[:, 1] == 0.0)          # Im X[0]: the Nyquist bin is dropped, not packed


def test_emulated_kernel_thread_count_independent(emu):
    x = synth.random_walk(301, 1024 + 70)
    a = np.empty((71, 1024)); b = np.empty((71, 1024))
    assert emu.emu_sliding(x, x.size, 1024, 32, 4, 256, a) == 0
    assert emu.emu_sliding(x, x.size, 1024, 32, 4, 96, b) == 0
    assert np.array_equal(a, b)


def test_emulated_kernel_rejects_unsupported_plans(emu):
    x = np.zeros(5000); out = np.zeros((1, 128))
    assert emu.emu_sliding(x, 200, 128, 32, 4, 256, out) == -1      # N < 256
    assert emu.emu_sliding(x, 2000, 1024, 30, 4, 256, np.zeros((977, 1024))) == -1   # T % S != 0
    assert emu.emu_sliding(x, 2000, 1024, 8, 8, 256, np.zeros((977, 1024))) == -1    # windows per sub-chain % 4 != 0


# radix-4 top pass (levels 2 -> 0), the low-register variant: (N, T, S)
PLANS4 = [(256, 64, 8), (512, 32, 4), (1024, 16, 2), (1024, 32, 4), (2048, 16, 2), (4096, 8, 2), (4096, 16, 2)]


@pytest.mark.parametrize("n,t,s", PLANS4)
def test_emulated_radix4_top_pass_matches_oracle(emu, oracle, n, t, s):
    for extra in (0, t - 1, 2 * t + 3):
        x = synth.random_walk(400 + n, n + extra)
        nw = x.size - n + 1
        out = np.full((nw, n), np.nan)
        assert emu.emu_sliding_top(x, x.size, n, t, s, 2, 256, out) == 0
        assert not np.isnan(out).any(), "a bin of some window was never produced"
        ref = np.stack([oracle.fft_interleaved(x[w:w + n]) for w in range(nw)])
        assert (np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-12
        off_dc = np.abs(out[:, 2:] - ref[:, 2:]).max(axis=1) / np.abs(ref[:, 2:]).max(axis=1)
        assert off_dc.max() < 1e-12
        assert np.all(out[:, 1] == 0.0)


@pytest.mark.parametrize("n,t,s", [(1024, 24, 2), (1024, 16, 2), (1024, 32, 2), (1024, 16, 1), (1024, 48, 2)])
def test_emulated_staged_rows_match_the_direct_store_kernel(emu, oracle, n, t, s):
    """sliding_staged_kernel's data path (64 threads per segment, rows collected and stored whole):
    every bin of every row written exactly once (rc -3 otherwise), and the rows are bit for bit the
    direct-store kernel's."""
    for extra in (0, t - 1, 2 * t + 5):
        x = synth.random_walk(310 + n, n + extra)
        nw = x.size - n + 1
        a = np.full((nw, n), np.nan); b = np.full((nw, n), np.nan)
        assert emu.emu_sliding_staged(x, x.size, n, t, s, a) == 0
        assert emu.emu_sliding(x, x.size, n, t, s, 256, b) == 0
        assert np.array_equal(a, b)
        ref = np.stack([oracle.fft_interleaved(x[w:w + n]) for w in range(nw)])
        assert (np.abs(a - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-12
