"""Builds fft_wavespec_b200/libwavespec.so in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwavespec.so")

# (source, extra flags).  ws_series.cu / ws_pla.cu hold the recursions whose results must be
# bit-identical to the CPU statement: no FMA contraction there.
SOURCES = [
    ("ws_abi.cu", []),
    ("ws_runtime.cu", []),
    ("ws_jobs.cu", []),
    ("ws_pipeline.cu", []),
    ("ws_window_fft.cu", []),
    ("ws_window_fft_warp.cu", []),
    ("ws_sliding.cu", []),
    ("ws_rows.cu", []),
    ("ws_inverse.cu", []),
    ("ws_phase.cu", ["-fmad=false"]),
    ("ws_series.cu", ["-fmad=false"]),
    ("ws_pla.cu", ["-fmad=false"]),
    ("ws_zigzag.cu", ["-fmad=false"]),
    ("ws_cache.cu", ["-fmad=false"]),
]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "wavespec_abi.h"))
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
    if force or _stale(OUT, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
