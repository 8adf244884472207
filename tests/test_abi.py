"""The C-ABI library builds for sm_100a without a GPU, loads, exports every symbol include/gpr_sm100a.h
declares, and fails loudly (no CPU fallback) when no device is present.  No compute calls here."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(gpr):
    from gpr_sm100a import _ffi
    L = _ffi.lib()
    declared = _ffi.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/gpr_sm100a.h but not exported"
        assert name in _ffi._SIGS, f"{name} has no ctypes signature in _ffi.py"
    assert L.gpr_version() >= 100


def test_library_is_plain_c_abi_and_sm100a(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert "gpr_nlml_grad" in syms and "gpr_predict" in syms
    assert not any("torch" in s.lower() or "at::" in s for s in syms)
    sass = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_dmma_and_ldgsts_in_sass(built_lib):
    """FP64 tensor pipe evidence: the GEMM lowers to DMMA; operands are staged with cp.async (LDGSTS)."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert sass.count("DMMA") >= 300
    assert "LDGSTS" in sass


def test_dim_hp_host_only(gpr):
    from gpr_sm100a import _ffi
    L = _ffi.lib()
    arr = (ctypes.c_int * 3)(1, 1, 2)
    assert L.gpr_dim_hp(arr, 3, 8) == 19      # SquaredExp + SquaredExp + WhiteNoise at D = 8
    bad = (ctypes.c_int * 1)(9)
    assert L.gpr_dim_hp(bad, 1, 8) < 0


def test_no_cpu_fallback_without_gpu(gpr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gpr.GPRError, match="no CPU fallback|CUDA"):
        gpr.Context(0)
    import numpy as np
    with pytest.raises(gpr.GPRError):
        gpr.kernel(gpr.SquaredExp(), np.ones(3), np.random.rand(2, 10))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gaussianprocessregression.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gpr_oracle" not in text and "import oracle" not in text, f
