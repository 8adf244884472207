"""[Historical A/B script: it produced profiles/ozaki_sweep_r2u.log with the flag meanings of that commit -- launch flag 4096 / option
ozaki_windows = 8 selected the two-window nine-digit product, flag 2048 / bit 2 the previous digit extraction.  Since then the two-window
form is the default, flag 4096 / ozaki_windows bit 2 selects the THREE-window form, and the slower extraction variant is gone.]
A/B of the INT8 route's knobs on the benchmark model (N = 32768) against the committed oracle golden: digit extraction kernel
(option ozaki_windows bit 2 = previous one-warp-per-column form), smallest routed product (ozaki_min).   python tools/ozaki_sweep.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gpr_sm100a import _ffi
import make_golden_config3 as m3

ctx = _ffi.get_context()
rng = np.random.default_rng(11)
for nn, kk in ((2048 + 128, 1024), (640, 640)):      # column counts that are not multiples of 96 / 256
    B = rng.standard_normal((kk, nn)) * np.exp(rng.uniform(-10, 0, (kk, nn)))
    r0, _ = _ffi.dbg_ozaki_dgemm(ctx, 1.0, B, B, 0.0, np.zeros((nn, nn)), S=9, flags=0)
    r1, _ = _ffi.dbg_ozaki_dgemm(ctx, 1.0, B, B, 0.0, np.zeros((nn, nn)), S=9, flags=4096)
    r2, _ = _ffi.dbg_ozaki_dgemm(ctx, 1.0, B, B, 0.0, np.full((nn, nn), 7.0), S=9, flags=4096 | 1)
    ref = B.astype(np.longdouble).T @ B.astype(np.longdouble)
    den = np.abs(B).T @ np.abs(B)
    print(f"n={nn} k={kk}: three windows err {np.max(np.abs(r0 - ref) / den):.2e}, two windows (128x96) err {np.max(np.abs(r1 - ref) / den):.2e}, "
          f"max|diff| {np.abs(r1 - r0).max():.2e}; upper-only: upper diff {np.abs(np.triu(r2) - np.triu(r1)).max():.2e}, "
          f"strict lower untouched {bool(np.all(np.tril(r2, -1) == np.tril(np.full((nn, nn), 7.0), -1)))}", flush=True)
n = 8192
A = rng.standard_normal((n, n))
for S, fl in ((8, 0), (8, 2048), (9, 1024), (9, 4096), (9, 4096 | 3), (9, 1024 | 3)):
    _, ms = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, np.zeros((n, n)), S=S, flags=fl, reps=5)
    print(f"timing 8192^3 S={S} flags={fl}: {ms:.3f} ms -> {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s FP64-equivalent (full product count)", flush=True)
del A
g = np.load(os.path.join(ROOT, "tests", "golden", "config3_n32768.npz"))
x, y, hp = m3.inputs()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], 8, x, y)
for win, omin in ((2 | 4, 1024), (2, 1024), (8, 1024), (2, 512), (2, 256), (8, 512), (2, 1024)):
    ctx.set_option("ozaki_windows", win)
    ctx.set_option("ozaki_min", omin)
    mh.nlml_grad(hp * 1.001)
    ts = []
    for rep in range(3):
        F, G = mh.nlml_grad(hp * (1 + 1e-9 * rep) if rep < 2 else hp)
        ts.append(mh.timings())
    t = {k: float(np.mean([q[k] for q in ts])) for k in ts[0]}
    relF = abs(F - float(g["F"])) / abs(float(g["F"]))
    relG = float((np.abs(G - g["G"]) / np.maximum(np.abs(g["G"]), 1e-8 * np.linalg.norm(g["G"]))).max())
    print(f"windows={win} ozaki_min={omin}: eval {t['eval']:.1f} ms (potrf {t['potrf']:.1f}, trtri {t['trtri']:.1f}, lauum {t['lauum']:.1f}); "
          f"vs oracle: relF {relF:.2e} relG {relG:.2e}", flush=True)
ctx.set_option("ozaki_windows", 0)
ctx.set_option("ozaki_min", 1024)
mh.close()
