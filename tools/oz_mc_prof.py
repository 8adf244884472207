"""One FP64 product through the two-window 8-digit INT8 kernels, single CTAs (flag 512) and multicast cluster pairs (512 | 8192), for ncu:
ncu --set full -k regex:oz_gemm_win python tools/oz_mc_prof.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _ffi.get_context()
A = np.asfortranarray(np.random.default_rng(0).standard_normal((n, n)))
for fl in (512, 512 | 8192):
    _, ms = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, np.zeros((n, n)), S=8, flags=fl)
    print(f"n={n} flags={fl}: {ms:.3f} ms", flush=True)
