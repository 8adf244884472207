"""One NLML + gradient evaluation of the benchmark model (N = 32768) after one warm-up, for the ncu launch list."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = 8
rng = np.random.default_rng(3003)
x = rng.random((D, N)); y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
ctx = _ffi.get_context()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
for rep in range(2):
    F, G = mh.nlml_grad(np.log(hp * (1 + 0.001 * rep)), log_scale=True)
print(F, float(np.linalg.norm(G)), {k: round(v, 2) for k, v in mh.timings().items() if v > 0})
