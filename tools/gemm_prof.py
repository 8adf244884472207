"""A few large launches of the DMMA GEMM (for `ncu --set full -k regex:dgemm128`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

ctx = _ffi.get_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((n, n)))
C0 = np.zeros((n, n), order="F")
for tA, tB in (("T", "N"), ("N", "N"), ("N", "T")):
    _, ms = _ffi.dbg_dgemm(ctx, tA, tB, 1.0, A, A, 1.0, C0, reps=2)
    print(tA + tB, n, f"{ms:.3f} ms", f"{2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
