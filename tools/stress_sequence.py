"""Reproducer harness for the open issue of DESIGN.md section 9 (seen once, in GPU call AR, never explained): after
  (a) a few dozen debug products through the INT8 kernels, (b) the block-cyclic driver with 2 ranks on device 0, (c) a forced-route
  N = 8192 model with prediction and split prediction on the default context,
two of three batches of N = 32768 evaluations on that same context came back wrong (F off by 1.8e-5 once, a failed Cholesky once).
This script replays the sequence with the SHIPPED library and checks EVERY evaluation against the committed oracle golden
(tests/golden/config3_n32768.npz, rel 1e-8) and against the first one bit for bit, then repeats it with the stages of the preamble
switched off one at a time so that a failure can be bisected.

    python tools/stress_sequence.py [evals_per_leg = 6] [stages = abc]      # e.g. "python tools/stress_sequence.py 10 c"

Not collected by pytest (it has never run on a GPU: the budget of round 2 ended with the call that showed the issue)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from gpr_sm100a import _ffi
import make_golden_config2 as mg2
import make_golden_config3 as m3

nev = int(sys.argv[1]) if len(sys.argv) > 1 else 6
stages = sys.argv[2] if len(sys.argv) > 2 else "abc"
t0 = time.time()
ctx = _ffi.get_context()
rng = np.random.default_rng(31)

if "a" in stages:      # debug products in many forms (what part 1 of tools/oz_epi_check.py did, without the option that no longer exists)
    for beta in (0.0, 0.5):
        for (M, N, K, S, fl) in ((256, 384, 640, 8, 0), (512, 512, 1024, 8, 1), (384, 384, 384, 8, 1 | 2), (512, 512, 512, 8, 64), (512, 640, 1024, 7, 0),
                                 (256, 256, 512, 6, 0), (512, 512, 1024, 8, 512), (1024, 1024, 1024, 8, 512 | 1 | 2), (512, 512, 1024, 8, 512 | 8192),
                                 (640, 640, 1024, 9, 0), (640, 640, 640, 9, 1 | 2), (640, 640, 1024, 9, 4096), (640, 640, 1024, 9, 4096 | 1024)):
            if fl & 2:
                A = np.tril(rng.standard_normal((K, M)))
                for J in range(M // 128):
                    A[:128 * J, 128 * J:128 * (J + 1)] = 1e30
                B = A
            else:
                A, B = rng.standard_normal((K, M)), (rng.standard_normal((K, N)) if S != 9 else None)
                B = A if B is None else B
            _ffi.dbg_ozaki_dgemm(ctx, -1.0, np.asfortranarray(A), np.asfortranarray(B), beta, np.asfortranarray(rng.standard_normal((M, N))), S=S, flags=fl)
    print(f"[a] debug products done  [{time.time() - t0:.0f} s]", flush=True)

x2, y2, sets = mg2.inputs()
hpA = sets["A"]
if "b" in stages:      # block-cyclic driver, two ranks on device 0, INT8 tile-mapped forms
    mc = _ffi.MultiContext([0, 0], nb=1024)
    mc.set_option("ozaki", 8)
    mc.set_option("ozaki_kchunk", 2048)
    mm = _ffi.MultiModelHandle(mc, [1, 2], 8, x2, y2)
    mm.nlml_grad(hpA)
    mm.close(); mc.close()
    print(f"[b] multi-rank driver done  [{time.time() - t0:.0f} s]", flush=True)

if "c" in stages:      # forced-route model with prediction and split prediction on the default context
    xp = np.asfortranarray(rng.random((8, 4096)))
    xe, xq = np.asfortranarray(0.5 * rng.random((8, 1024))), np.asfortranarray(0.5 * rng.random((8, 1024)))
    ctx.set_option("ozaki", 8)
    mh = _ffi.ModelHandle(ctx, [1, 2], 8, x2, y2)
    mh.nlml_grad(hpA)
    mh.predict(xp, want_var=True)
    mh.split_predict(xe, xq, var_range=None, want_var=False)
    mh.close()
    ctx.set_option("ozaki", -1)
    print(f"[c] forced-route model done  [{time.time() - t0:.0f} s]", flush=True)

g = np.load(os.path.join(ROOT, "tests", "golden", "config3_n32768.npz"))
x, y, hp = m3.inputs()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], 8, x, y)
first, bad = None, 0
for leg in range(3):
    for rep in range(nev):
        try:
            mh.nlml_grad(hp * (1 + 1e-6 * (rep + 1)))          # a different point in between, as an optimiser would
            F, G = mh.nlml_grad(hp)
        except Exception as e:                                 # a failed factorization is exactly what the issue looked like
            bad += 1
            print(f"leg {leg} evaluation {rep}: EXCEPTION {e}", flush=True)
            continue
        relF = abs(F - float(g["F"])) / abs(float(g["F"]))
        relG = float((np.abs(G - g["G"]) / np.maximum(np.abs(g["G"]), 1e-8 * np.linalg.norm(g["G"]))).max())
        if first is None:
            first = (F, G.copy())
        same = F == first[0] and bool(np.array_equal(G, first[1]))
        ok = relF <= 1e-8 and relG <= 1e-8 and same
        bad += not ok
        if not ok or rep == nev - 1:
            print(f"leg {leg} evaluation {rep}: relF {relF:.2e} relG {relG:.2e} identical to the first {same} {'' if ok else '  <-- WRONG'}", flush=True)
mh.close()
print(f"RESULT stages={stages}: {'OK' if bad == 0 else str(bad) + ' BAD evaluations'}  [{time.time() - t0:.0f} s]")
sys.exit(1 if bad else 0)
