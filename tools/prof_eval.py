"""One pass over every kernel of the hot path, for ncu (`--set full -k regex:...` / the launch list):
one NLML + gradient evaluation (3-ref model), one tile of prediction (mean + variance), one split prediction,
and one evaluation through the block-cyclic multi-rank path (2 ranks on cuda:0).
    python tools/prof_eval.py [N] [M]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
D = 8
rng = np.random.default_rng(3003)
x = rng.random((D, N))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
ctx = _ffi.get_context()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
F, G = mh.nlml_grad(np.log(hp), log_scale=True)
print("nlml", F, float(np.linalg.norm(G)), {k: round(v, 2) for k, v in mh.timings().items() if v > 0})
xp = np.asfortranarray(rng.random((D, M)))
mu, var, _ = mh.predict(xp, want_var=True)
print("predict", float(mu.mean()), float(var.mean()), {k: round(v, 2) for k, v in mh.timings().items() if k.startswith("pred")})
xe, xq = 0.5 * rng.random((D, 64)), 0.5 * rng.random((D, 64))
smu, svar = mh.split_predict(xe, xq, var_range=(1, 3))
print("split", float(smu.mean()), float(svar.mean()))
mh.close()
mc = _ffi.MultiContext([0, 0], nb=512)
mm = _ffi.MultiModelHandle(mc, [1, 1, 2], D, x, y)
F2, G2 = mm.nlml_grad(np.log(hp), log_scale=True)
print("mgpu", F2, abs(F2 - F) / abs(F), {k: round(v, 2) for k, v in mm.timings().items() if v > 0})
mm.close(); mc.close()
