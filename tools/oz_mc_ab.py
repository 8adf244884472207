"""A/B: cluster pairs with a multicast op(B) tile for the 128 x 128 window kernels of the INT8 route (option ozaki_mc, launch flag 8192).
  1. isolated products (gpr_dbg_ozaki_dgemm): time and bit-identity against the single-CTA windows;
  2. the benchmark model (N = 32768) against the committed oracle golden with the option off / on / off.
python tools/oz_mc_ab.py [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gpr_sm100a import _ffi

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = _ffi.get_context()
rng = np.random.default_rng(5)
t0 = time.time()
n = 4096 if quick else 8192
A = np.asfortranarray(rng.standard_normal((n, n)))
C0 = np.zeros((n, n), order="F")
for (S, base, what) in ((8, 512, "8 digits, two windows"), (8, 512 | 1 | 2, "8 digits, two windows, upper + K-from-N"), (9, 0, "9 digits (second window)"),
                        (9, 1 | 2, "9 digits, upper + K-from-N (lauum)")):
    C1, ms1 = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, C0, S=S, flags=base, reps=4)
    C2, ms2 = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, C0, S=S, flags=base | 8192, reps=4)
    same = bool(np.array_equal(C1, C2))
    print(f"{n}^3 {what}: single CTAs {ms1:.2f} ms, multicast pairs {ms2:.2f} ms ({ms1 / ms2:.3f} x), bit-identical {same}", flush=True)
    del C1, C2
del A, C0
print(f"[isolated products done in {time.time() - t0:.0f} s]", flush=True)

import make_golden_config3 as m3

g = np.load(os.path.join(ROOT, "tests", "golden", "config3_n32768.npz"))
x, y, hp = m3.inputs()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], 8, x, y)
res = {}
for mc in (0, 1, 0, 1):
    ctx.set_option("ozaki_mc", mc)
    mh.nlml_grad(hp * 1.001)
    ts = []
    for rep in range(3):
        F, G = mh.nlml_grad(hp * (1 + 1e-9 * rep) if rep < 2 else hp)
        ts.append(mh.timings())
    t = {k: float(np.mean([q[k] for q in ts])) for k in ts[0]}
    relF = abs(F - float(g["F"])) / abs(float(g["F"]))
    relG = float((np.abs(G - g["G"]) / np.maximum(np.abs(g["G"]), 1e-8 * np.linalg.norm(g["G"]))).max())
    if mc in res:
        print(f"    F identical to the first ozaki_mc={mc} run: {F == res[mc][0]}, G identical: {bool(np.array_equal(G, res[mc][1]))}")
    res.setdefault(mc, (F, G))
    print(f"ozaki_mc={mc}: eval {t['eval']:.1f} ms (potrf {t['potrf']:.1f}, trtri {t['trtri']:.1f}, lauum {t['lauum']:.1f}); relF {relF:.2e} relG {relG:.2e}", flush=True)
print(f"F identical between the routes: {res[0][0] == res[1][0]}, G identical: {bool(np.array_equal(res[0][1], res[1][1]))}")
mh.close()
print(f"[all done in {time.time() - t0:.0f} s]")
