"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/wavespec_abi.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bridge():
    import __graft_entry__ as g
    g.build()
    from fft_wavespec_b200 import bridge as b
    return b


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wavespec_abi.h")).read()
    return re.findall(r"WAVESPEC_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)


def test_header_declares_the_imports_mqh_surface():
    syms = declared_symbols()
    # Include/imports.mqh:6-20 + the two Legacy declarations
    for s in ["gpu_init", "gpu_shutdown", "gpu_fft_real_forward", "gpu_extract_cycles",
              "gpu_submit_extract_cycles", "gpu_try_get_cycles", "gpu_submit_extract_cycles_batch",
              "gpu_try_get_cycles_batch", "gpu_free_job", "gpu_get_last_error_w",
              "gpu_fft_real_inverse", "gpu_fft_real_forward_batch"]:
        assert s in syms


def test_library_exports_every_declared_symbol(bridge):
    L = bridge.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in wavespec_abi.h but not exported"
    assert set(syms) == set(bridge.EXPORTED_SYMBOLS)


def test_cfg_struct_matches_oracle_and_defaults(bridge, oracle):
    assert C.sizeof(bridge.PipelineCfg) == C.sizeof(oracle.PipelineCfg)
    a = bridge.default_cfg(1024)
    b = oracle.default_cfg(1024)
    for (name, _) in bridge.PipelineCfg._fields_:
        if name == "kalman":
            for (kn, _) in bridge.Kalman4DParams._fields_:
                assert getattr(a.kalman, kn) == getattr(b.kalman, kn), kn
        else:
            assert getattr(a, name) == getattr(b, name), name
    assert bridge.lib().wavespec_num_windows(1000000, 1024, 1) == 998977
    assert bridge.lib().wavespec_num_windows(100, 1024, 1) == 0


def test_no_device_means_backend_unavailable(bridge):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    st = bridge.gpu_init(0, 4)
    assert st == bridge.BACKEND_UNAVAILABLE
    assert "CUDA" in bridge.last_error() or "device" in bridge.last_error()
    with pytest.raises(bridge.WaveSpecError) as e:
        bridge.gpu_fft_real_forward(np.zeros(64))
    assert e.value.status == bridge.BACKEND_UNAVAILABLE      # fails loudly, never computes on the CPU
    st, jid = bridge.gpu_submit_extract_cycles_batch(np.zeros(4096), 1024, 1, 4, 9.0, 200.0)
    assert st == bridge.BACKEND_UNAVAILABLE and jid == 0


def test_package_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "fft_wavespec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
