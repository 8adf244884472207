"""One large FP64 product through the INT8 tensor cores for ncu: 8192^3, eight digits (oz_gemm_kernel<8>) and nine digits (the three
diagonal windows of the inverse's W^T W).   ncu --set full -k regex:oz_gemm python tools/ozaki_prof.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = _ffi.get_context()
A = np.asfortranarray(np.random.default_rng(0).standard_normal((n, n)))
for S, fl in ((8, 0), (9, 0)):
    _, ms = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, np.zeros((n, n)), S=S, flags=fl)
    print(f"S={S} flags={fl}: {ms:.3f} ms", flush=True)
