"""A few large launches of the DMMA GEMM (for `ncu --set full -k regex:dgemm128`): the T,N form through the LDGSTS ring
(dgemm_sm100.cuh) and through the TMA ring (dgemm_tma.cuh), then the N,N and N,T forms (LDGSTS only)."""
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
for tma in (0, 1):
    ctx.set_option("gemm_tma", tma)
    _, ms = _ffi.dbg_dgemm(ctx, "T", "N", 1.0, A, A, 1.0, C0, reps=2)
    print("TN", "tma" if tma else "ldgsts", n, f"{ms:.3f} ms", f"{2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
for tA, tB in (("N", "N"), ("N", "T")):
    _, ms = _ffi.dbg_dgemm(ctx, tA, tB, 1.0, A, A, 1.0, C0, reps=2)
    print(tA + tB, n, f"{ms:.3f} ms", f"{2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
