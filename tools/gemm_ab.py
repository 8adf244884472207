"""A/B of the T,N GEMM: LDGSTS ring (dgemm_sm100.cuh) vs TMA ring (dgemm_tma.cuh), isolated kernel and whole evaluation.
    python tools/gemm_ab.py [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = _ffi.get_context()
rng = np.random.default_rng(0)
for n in (4096, 8192):
    A = np.asfortranarray(rng.standard_normal((n, n)))
    C0 = np.zeros((n, n), order="F")
    for tma in (0, 1):
        ctx.set_option("gemm_tma", tma)
        C, ms = _ffi.dbg_dgemm(ctx, "T", "N", 1.0, A, A, 0.0, C0, reps=6)
        print(f"dgemm T,N {n}^3 tma={tma}: {ms:.3f} ms = {2 * n ** 3 / ms / 1e9:.2f} TFLOP/s, checksum {float(C[::97, ::89].sum()):.6e}", flush=True)
    del A, C0
D = 8
x = rng.random((D, N))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
for tma in (0, 1, 0, 1):
    ctx.set_option("gemm_tma", tma)
    mh.nlml_grad(np.log(hp * 1.001), log_scale=True)
    F, G = mh.nlml_grad(np.log(hp), log_scale=True)
    t = mh.timings()
    print(f"N={N} tma={tma}: eval {t['eval']:.1f} ms (potrf {t['potrf']:.1f}, trtri {t['trtri']:.1f}, lauum {t['lauum']:.1f}, kbuild {t['kbuild']:.2f}, grad {t['grad']:.2f}) F={F!r} |G|={float(np.linalg.norm(G))!r}", flush=True)
ctx.set_option("gemm_tma", 0)
mh.close()
