"""Covariance-side kernels in isolation, for ncu and for CUDA-event timing: the training build, the gradient
reduction and the mean-only K* build at N (default 8192), with the TMA/Gram build and with the direct-difference one.
    python tools/prof_cov.py [N] [M]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
D = 8
rng = np.random.default_rng(3003)
x = rng.random((D, N))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
xp = np.asfortranarray(rng.random((D, M)))
ctx = _ffi.get_context()
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
for flag in (1, 0):
    ctx.set_option("kbuild_gram", flag)
    for rep in range(2):
        F, G = mh.nlml_grad(np.log(hp * (1 + 0.001 * rep)), log_scale=True)
    t = mh.timings()
    mu, _, _ = mh.predict(xp, want_var=False)
    mu, _, _ = mh.predict(xp, want_var=False)
    tp = mh.timings()
    print(f"kbuild_gram={flag}: F={F:.6f} kbuild {t['kbuild']:.3f} ms grad {t['grad']:.3f} ms; K* (mean only, {M} pts) {tp['pred_kstar']:.3f} ms; mean[0]={mu[0, 0]:.12f}")
mh.close()
