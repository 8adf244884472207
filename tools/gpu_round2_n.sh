#!/bin/bash
# mgpu INT8 potrf test (virtual ranks on one GPU) + config5-like timing at N = 32768 with 1 rank
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mgpu or ozaki or tma" > gpurun_out/r2n_pytest.log 2>&1
tail -5 gpurun_out/r2n_pytest.log; grep -E "mgpu potrf|ozaki |INT8" gpurun_out/r2n_pytest.log | head -20
python - <<'PY' > gpurun_out/r2n_mgpu_timing.log 2>&1
import sys, os, time
sys.path.insert(0, "gaussianprocessregression.jl_b200")
import numpy as np
from gpr_sm100a import _ffi
N, D = 32768, 16
rng = np.random.default_rng(5005)
x = np.asfortranarray(rng.random((D, N))); y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.4 * np.ones(D), [0.1]])
for oz in (0, -1):
    mc = _ffi.MultiContext([0], nb=1024); mc.set_option("ozaki", oz)
    mm = _ffi.MultiModelHandle(mc, [1, 2], D, x, y)
    mm.nlml_grad(hp * 0.99); F, G = mm.nlml_grad(hp); t = mm.timings()
    print(f"mgpu 1 rank N={N} ozaki={oz}: potrf {t['potrf']:.1f} trtri {t['trtri']:.1f} lauum {t['lauum']:.1f} total {t['total']:.1f} ms F={F!r} |G|={float(np.linalg.norm(G))!r}", flush=True)
    mm.close(); mc.close()
PY
cat gpurun_out/r2n_mgpu_timing.log
