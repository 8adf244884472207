"""How long does one 128 x 128 diagonal-block factorization (potrf_leaf_kernel: Cholesky + explicit inverse, one CTA) take in situ, i.e. with
warm instruction caches and no profiler?  gpr_dbg_factor(mode 0) on n = 128 (one leaf launch) .. 4096, repeated.   python tools/leaf_time.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

ctx = _ffi.get_context()
rng = np.random.default_rng(0)
for n in (128, 256, 512, 1024, 2048, 4096):
    X = rng.standard_normal((n, n))
    K = X @ X.T / n + np.eye(n)
    ts = []
    for rep in range(6):
        _, ms = _ffi.dbg_factor(ctx, K.copy(order="F"), mode=0)
        ts.append(ms)
    print(f"potrf n={n:5d} ({n // 128:3d} leaves): first {ts[0] * 1e3:8.1f} us, then min {min(ts[1:]) * 1e3:8.1f} us, median {np.median(ts[1:]) * 1e3:8.1f} us "
          f"-> per leaf <= {min(ts[1:]) * 1e3 / (n // 128):6.1f} us", flush=True)
