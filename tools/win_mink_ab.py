"""A/B: two-diagonal-window form of the 8-digit INT8 product for the LARGE products of potrf / trtri only (option ozaki_win_mink) on the
benchmark model (N = 32768) against the committed oracle golden.   python tools/win_mink_ab.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gpr_sm100a import _ffi
import make_golden_config3 as m3

ctx = _ffi.get_context()
g = np.load(os.path.join(ROOT, "tests", "golden", "config3_n32768.npz"))
x, y, hp = m3.inputs()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], 8, x, y)
for mink in (0, 16384, 8192, 4096, 2048, 0):
    ctx.set_option("ozaki_win_mink", mink)
    mh.nlml_grad(hp * 1.001)
    ts = []
    for rep in range(3):
        F, G = mh.nlml_grad(hp * (1 + 1e-9 * rep) if rep < 2 else hp)
        ts.append(mh.timings())
    t = {k: float(np.mean([q[k] for q in ts])) for k in ts[0]}
    relF = abs(F - float(g["F"])) / abs(float(g["F"]))
    relG = float((np.abs(G - g["G"]) / np.maximum(np.abs(g["G"]), 1e-8 * np.linalg.norm(g["G"]))).max())
    print(f"ozaki_win_mink={mink}: eval {t['eval']:.1f} ms (potrf {t['potrf']:.1f}, trtri {t['trtri']:.1f}, lauum {t['lauum']:.1f}); relF {relF:.2e} relG {relG:.2e}", flush=True)
ctx.set_option("ozaki_win_mink", 0)
mh.close()
