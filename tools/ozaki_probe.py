import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi
ctx = _ffi.get_context()
rng = np.random.default_rng(0)
n = 8192
A = np.asfortranarray(rng.standard_normal((n, n)))
C0 = np.zeros((n, n), order="F")
for S in (8, 6):
    for flags, name in ((512, "two diagonal windows, 128 x 128 tiles" if S == 8 else "single pass"), (0, "single pass, 128 x 64 tiles"), (256, "single pass, no reload")):
        C, ms = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, C0, S=S, flags=flags, reps=4)
        pairs = S * (S + 1) // 2
        tiles = (n // 128) * (n // 64)
        waves = -(-tiles // 148)
        clk_per_mma = ms * 1e-3 * 1.965e9 / waves / (n // 64) / (pairs * 2)
        print(f"S={S} {name}: {ms:.3f} ms, {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s-equiv, ~{clk_per_mma:.1f} clk per UTCIMMA (M128 N64 K32), int8 {2 * n ** 3 * pairs / ms / 1e12:.2f} POPS", flush=True)
