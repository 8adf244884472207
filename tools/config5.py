"""BASELINE.json config 5: N = 131072, D = 16, SquaredExp()+WhiteNoise() (P = 18) NLML + gradient with the block-row
covariance build and the block-cyclic distributed Cholesky / inverse over 1/2/4/8 B200.

    python tools/config5.py [--n 131072] [--gpus 8,4] [--nb 1024] [--evals 1]

The oracle is infeasible at this N (>= 4 x 137 GB on the host), so parity at full size is established by
invariants: (1) F and G do not depend on the number of GPUs (rtol 1e-9), (2) K alpha = y on sampled columns,
(3) the directional derivative of F along a random direction matches G (two more loss evaluations; --fd).
tests/test_gpu_parity.py holds the oracle comparison of the same code path at small N.
One JSON line per GPU count."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))


def se_noise_columns(x, hp, cols):
    """K[:, cols] of SquaredExp()+WhiteNoise() incl. jitter and noise on the diagonal (src/covariance.jl:85-95,
    src/compose_covar.jl:63-77), plain numpy -- only used for the sampled residual check."""
    sig, ell, sn = hp[0], hp[1:-1], hp[-1]
    xs = x * ell[:, None]
    d = ((xs[:, :, None] - xs[:, None, cols]) ** 2).sum(0)
    K = sig * sig * np.exp(-d)
    K[cols, np.arange(len(cols))] += 1e-8 + sn * sn
    return K


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=131072)
    ap.add_argument("--d", type=int, default=16)
    ap.add_argument("--gpus", default="8")
    ap.add_argument("--nb", type=int, default=1024)
    ap.add_argument("--evals", type=int, default=1, help="timed evaluations per GPU count (after one warm-up if > 1)")
    ap.add_argument("--prefetch", default="", help="prefetch_trtri,prefetch_lauum (0 none, 1 SM pulls, 2 copy engines)")
    ap.add_argument("--fd", action="store_true", help="also check the gradient by a central difference of F")
    args = ap.parse_args()
    from gpr_sm100a import _ffi

    N, D = args.n, args.d
    rng = np.random.default_rng(5005)
    x = np.asfortranarray(rng.random((D, N)))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    hp = np.concatenate([[1.0], 0.4 * np.ones(D), [0.1]])
    ndev = _ffi.device_count()
    ref = None
    for G in [int(g) for g in args.gpus.split(",")]:
        devs = [r % ndev for r in range(G)]
        t0 = time.perf_counter()
        mc = _ffi.MultiContext(devs, nb=args.nb)
        if args.prefetch:
            pt, pl = [int(v) for v in args.prefetch.split(",")]
            mc.set_option("prefetch_trtri", pt)
            mc.set_option("prefetch_lauum", pl)
        mm = _ffi.MultiModelHandle(mc, [1, 2], D, x, y)
        t_setup = time.perf_counter() - t0
        if args.evals > 1:
            mm.nlml_grad(hp * 0.99)
        ts, tms = [], []
        for e in range(args.evals):
            t0 = time.perf_counter()
            F, Gd = mm.nlml_grad(hp)
            ts.append(time.perf_counter() - t0)
            tms.append(mm.timings())
        alpha = mm.fetch(_ffi.FETCH_ALPHA)
        cols = np.random.default_rng(1).choice(N, 32, replace=False)
        Kc = se_noise_columns(x, hp, cols)
        resid = float(np.abs(Kc.T @ alpha - y[cols]).max() / np.abs(y).max())
        out = {"config": 5, "N": N, "D": D, "P": len(hp), "gpus": G, "devices": devs, "nb": args.nb, "prefetch": args.prefetch or "default (2,2)", "setup_s": round(t_setup, 2),
               "s_per_eval": [round(t, 3) for t in ts], "evals_per_s": 1.0 / min(ts),
               "phase_ms": {k: round(v, 1) for k, v in tms[-1].items() if v > 0},
               "dense_tflops_aggregate": N ** 3 / ((tms[-1]["potrf"] + tms[-1]["trtri"] + tms[-1]["lauum"]) * 1e-3) / 1e12,
               "F": F, "G_norm": float(np.linalg.norm(Gd)), "resid_K_alpha_minus_y": resid, "launches": mc.launch_count()}
        if ref is None:
            ref = (F, Gd.copy(), G)
        else:
            out["vs_gpus"] = ref[2]
            out["relF_vs_ref"] = abs(F - ref[0]) / abs(ref[0])
            out["relG_vs_ref"] = float((np.abs(Gd - ref[1]) / np.maximum(np.abs(ref[1]), 1e-8 * np.linalg.norm(ref[1]))).max())
        if args.fd and ref[2] == G:   # once, on the first GPU count
            v = np.random.default_rng(2).standard_normal(len(hp))
            v /= np.linalg.norm(v)
            h = 1e-5
            Fp, _ = mm.nlml_grad(hp + h * v, want_g=False)
            Fm, _ = mm.nlml_grad(hp - h * v, want_g=False)
            out["fd_directional"] = (Fp - Fm) / (2 * h)
            out["G_dot_v"] = float(Gd @ v)
        print(json.dumps(out), flush=True)
        mm.close()
        mc.close()


if __name__ == "__main__":
    main()
