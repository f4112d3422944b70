"""fft_wavespec_b200 — B200-native (sm_100a) sliding-window spectral hot path of the WaveSpecZZ
indicators behind the reference's DLL import surface (Include/imports.mqh).

The package is a thin host mirror (ctypes) over libwavespec.so, a C-ABI CUDA library; it has
no CPU compute path.  `bridge` carries the reference-named entry points, `synth` the seeded
synthetic price series used by tests and bench.
"""
from . import synth  # noqa: F401

__all__ = ["synth", "bridge", "build"]
__version__ = "0.1.0"
