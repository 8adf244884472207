"""Whole NLML + gradient evaluation (bench workload, N = 32768) with the FP64 products on the DMMA pipe vs through the INT8
tensor cores (option "ozaki" = number of 7-bit digits), against the committed oracle golden.   python tools/ozaki_eval.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gpr_sm100a import _ffi
import make_golden_config3 as m3

ctx0 = _ffi.get_context()
rng = np.random.default_rng(1)
for n in (2048,):
    X = rng.standard_normal((n, n))
    Kmat = X @ X.T / n + np.eye(n)
    for mode in (0, 3):
        res = []
        for oz in (0, 8):
            ctx0.set_option("ozaki", oz)
            ctx0.set_option("ozaki_min", 512)
            res.append(_ffi.dbg_factor(ctx0, Kmat.copy(order="F"), mode=mode)[0])
        ref = np.linalg.cholesky(Kmat).T if mode == 0 else np.linalg.inv(Kmat)
        print(f"dbg_factor n={n} mode={mode}: dmma err {np.abs(np.triu(res[0]) - np.triu(ref)).max():.2e}, ozaki err {np.abs(np.triu(res[1]) - np.triu(ref)).max():.2e}", flush=True)
ctx0.set_option("ozaki", 0)
g = np.load(os.path.join(ROOT, "tests", "golden", "config3_n32768.npz"))
x, y, hp = m3.inputs()
ctx = _ffi.get_context()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], 8, x, y)
for oz, omin, ph, pan in ((0, 1024, 15, 4096), (8, 1024, 15, 32768), (8, 1024, 15, 8192), (8, 1024, 15, 4096), (8, 1024, 15, 2048), (8, 1024, 15, 1024), (8, 1024, 3, 4096)):
    ctx.set_option("ozaki", oz)
    ctx.set_option("ozaki_min", omin)
    ctx.set_option("ozaki_phases", ph)
    ctx.set_option("ozaki_panel", pan)
    mh.nlml_grad(hp * 1.001)
    F, G = mh.nlml_grad(hp)
    t = mh.timings()
    relF = abs(F - float(g["F"])) / abs(float(g["F"]))
    relG = float((np.abs(G - g["G"]) / np.maximum(np.abs(g["G"]), 1e-8 * np.linalg.norm(g["G"]))).max())
    alpha = mh.fetch(_ffi.FETCH_ALPHA)
    rela = float(np.abs(alpha - g["alpha"]).max() / np.abs(g["alpha"]).max())
    print(f"ozaki={oz} min={omin} phases={ph} panel={pan}: eval {t['eval']:.1f} ms (potrf {t['potrf']:.1f}, trtri {t['trtri']:.1f}, lauum {t['lauum']:.1f}); "
          f"vs oracle: relF {relF:.2e} relG {relG:.2e} alpha {rela:.2e}", flush=True)
ctx.set_option("ozaki", 0)
mh.close()
