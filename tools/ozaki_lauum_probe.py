"""Where does the INT8 W^T W product lose accuracy?  K^-1 of the benchmark model at N = 8192 with lauum on DMMA vs INT8."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi
N, D = 8192, 8
rng = np.random.default_rng(3003)
x = rng.random((D, N)); y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
ctx = _ffi.get_context()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
res = {}
for name, oz, ph in (("dmma", 0, 15), ("int8_potrf_trtri", 8, 3), ("int8_all", 8, 7), ("int8_lauum_only", 8, 4)):
    ctx.set_option("ozaki", oz); ctx.set_option("ozaki_phases", ph); ctx.set_option("ozaki_min", 1024); ctx.set_option("ozaki_panel", 32768)
    F, G = mh.nlml_grad(hp * (1 + 1e-13 * len(res)))
    Ki = mh.fetch(_ffi.FETCH_KINV)
    res[name] = (F, G, Ki)
F0, G0, K0 = res["dmma"]
for name in ("int8_potrf_trtri", "int8_all", "int8_lauum_only"):
    F, G, Ki = res[name]
    dg = np.diag(Ki) - np.diag(K0)
    off = Ki - K0
    print(f"{name}: relF {abs(F - F0) / abs(F0):.2e} relG {np.abs((G - G0) / np.maximum(np.abs(G0), 1e-8 * np.linalg.norm(G0))).max():.2e}; "
          f"Kinv: max|diff| {np.abs(off).max():.2e} (max|Kinv| {np.abs(K0).max():.1f}), diag rel diff mean {np.mean(dg / np.diag(K0)):.2e} "
          f"rms {np.sqrt(np.mean((dg / np.diag(K0)) ** 2)):.2e}, trace rel diff {dg.sum() / np.trace(K0):.2e}, "
          f"sum(all) rel diff {off.sum() / np.abs(K0).sum():.2e}, rms offdiag diff {np.sqrt(np.mean(off ** 2)):.2e}", flush=True)
print("column maxima of Kinv rows (first 5):", np.abs(K0).max(0)[:5], " typical |offdiag|:", np.median(np.abs(K0)))
