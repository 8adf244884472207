"""Prototype check of the INT8-tensor-core FP64 GEMM (csrc/ozaki_i8.cuh): correctness against numpy / longdouble on small
and medium sizes, then throughput against the DMMA kernel.    python tools/ozaki_check.py [quick]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi

ctx = _ffi.get_context()
rng = np.random.default_rng(0)


def check(M, N, K, S, alpha=1.0, beta=0.0, flags=0, kind="gauss"):
    if kind == "gauss":
        A, B = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    elif kind == "ones":
        A, B = np.ones((K, M)), np.ones((K, N))
    else:   # decaying magnitudes along k and across rows, like a Cholesky panel
        A = rng.standard_normal((K, M)) * np.exp(-3 * rng.random((K, M))) * np.exp(-2 * rng.random(M))[None, :]
        B = rng.standard_normal((K, N)) * np.exp(-3 * rng.random((K, N))) * np.exp(-2 * rng.random(N))[None, :]
    A, B = np.asfortranarray(A), np.asfortranarray(B)
    C0 = np.asfortranarray(rng.standard_normal((M, N)))
    C, ms = _ffi.dbg_ozaki_dgemm(ctx, alpha, A, B, beta, C0, S=S, flags=flags)
    ref = alpha * (A.astype(np.longdouble).T @ B.astype(np.longdouble)) + beta * C0
    den = np.abs(A).T @ np.abs(B) + 1e-300
    mask = np.ones((M, N), dtype=bool)
    if flags & 1:
        mask = np.triu(mask)
    err = float(np.max((np.abs(C - ref) / den)[mask]))
    d64 = alpha * (A.T @ B) + beta * C0
    err64 = float(np.max((np.abs(d64 - ref) / den)[mask]))
    print(f"M={M} N={N} K={K} S={S} {kind} alpha={alpha} beta={beta} flags={flags}: err/|a||b| = {err:.2e} (numpy dgemm {err64:.2e})  "
          f"C[0,0]={C[0, 0]:.6g} ref {float(ref[0, 0]):.6g}", flush=True)
    return err


check(128, 128, 128, 2, kind="ones")
check(128, 128, 128, 8, kind="ones")
check(128, 128, 128, 8)
check(256, 384, 512, 8)
check(256, 256, 1024, 8, alpha=-1.0, beta=1.0)
check(384, 384, 384, 8, flags=1, beta=1.0)
check(512, 512, 2048, 8, kind="decay")
check(512, 512, 2048, 7, kind="decay")
check(256, 256, 32768, 8)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(0)
for n in (4096, 8192):
    A = np.asfortranarray(rng.standard_normal((n, n)))
    C0 = np.zeros((n, n), order="F")
    for S in (8, 7, 6):
        C, ms = _ffi.dbg_ozaki_dgemm(ctx, 1.0, A, A, 0.0, C0, S=S, reps=4)
        print(f"ozaki S={S} {n}^3: {ms:.3f} ms = {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s FP64-equivalent (slicing included)", flush=True)
    C, ms = _ffi.dbg_dgemm(ctx, "T", "N", 1.0, A, A, 0.0, C0, reps=4)
    print(f"dmma (tma) {n}^3: {ms:.3f} ms = {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
