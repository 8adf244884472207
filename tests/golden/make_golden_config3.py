"""Oracle values at the BENCHMARKED configuration (BASELINE.json config 3 "3-ref" + config 4):
SquaredExp()+SquaredExp()+WhiteNoise(), N = 32768, D = 8, P = 19, the exact inputs bench.py times (seed 3003).

Stored in config3_n32768.npz (inputs are regenerated from the seeds, only results are stored):
  F, G (natural space), G_log (log space), alpha (all N), cond_est,
  pred_mean / pred_var at 4096 general test points (seed 4004),
  split predict on the ne = nq = 4096 grid of config 4: mean rows e = 1..3 (all q), mean at 4096 sampled (e, q),
  variance of the e rows 1..3 (the reference's default var_range, q fastest).
Arithmetic: oracle/gpr_oracle_big.nlml_grad_lean (same LAPACK calls as the reference, pinned to gpr_oracle.loss_grad
in tests/test_oracle_big.py) and gpr_oracle.predict / split_predict on that factor.  OpenBLAS's dpotrf fails at this
order in this image (see oracle/gpr_oracle_big.py), so factor and solves are blocked over LAPACK/BLAS calls of order
<= 8192 and the factor / inverse are verified against K on sampled columns before anything is stored.
Needs ~30 GB of host memory and ~25 min on 8 cores:  python tests/golden/make_golden_config3.py
"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import gpr_oracle as o  # noqa: E402
import gpr_oracle_big as ob  # noqa: E402

D, N, SEED, SEED_TEST = 8, 32768, 3003, 4004
M_TEST, NE, NQ, N_SAMPLED = 4096, 4096, 4096, 4096
COV = (o.SE, o.SE, o.NOISE)


def inputs(n=N):
    """Identical to bench.make_problem(N, D): the benchmarked x, y, hp."""
    rng = np.random.default_rng(SEED)
    x = rng.random((D, n))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(n)
    hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
    return x, y, hp


def test_inputs(m=M_TEST, ne=NE, nq=NQ, nsamp=N_SAMPLED):
    rng = np.random.default_rng(SEED_TEST)
    xp = rng.random((D, m))
    xe = 0.5 * rng.random((D, ne))
    xq = 0.5 * rng.random((D, nq))
    samp = rng.choice(ne * nq, nsamp, replace=False)       # flat indices into mean[e, q] laid out e fastest
    return xp, xe, xq, samp


def make(n=N, m=M_TEST, ne=NE, nq=NQ, nsamp=N_SAMPLED, log=print):
    x, y, hp = inputs(n)
    xp, xe, xq, samp = test_inputs(m, ne, nq, nsamp)
    t0 = time.perf_counter()
    F, G, alpha, U, Kinv = ob.nlml_grad_lean(COV, hp, x, y, log=log)
    # independent checks of the big intermediate results against K itself (sampled columns)
    vf, vi = ob.verify_factor(U, COV, hp, x), ob.verify_inverse_columns(Kinv, COV, hp, x)
    log(f"verify: max|U^T U - K|/max|K| = {vf:.2e} (48 columns), max|K K^-1 - I| = {vi:.2e} (32 columns)")
    assert vf < 1e-12 and vi < 1e-9
    v = np.ones(n) / np.sqrt(n)
    for _ in range(30):                                     # lambda_min(K) from power iteration on K^-1
        v = Kinv @ v
        lam = np.linalg.norm(v)
        v /= lam
    lmin = 1.0 / lam
    del Kinv
    # lambda_max(K) <= trace bound is loose; use power iteration through the factor: K v = U^T (U v) on the upper triangle
    w = np.ones(n) / np.sqrt(n)
    for _ in range(20):
        t = np.zeros(n)
        for k0 in range(0, n, 4096):
            k1 = min(n, k0 + 4096)
            Wb = np.triu(U[k0:k1, k0:], 0)
            t[k0:k1] = Wb @ w[k0:]
        t2 = np.zeros(n)
        for k0 in range(0, n, 4096):
            k1 = min(n, k0 + 4096)
            Wb = np.triu(U[k0:k1, k0:], 0)
            t2[k0:] += Wb.T @ t[k0:k1]
        lmax = np.linalg.norm(t2)
        w = t2 / lmax
    out = {"F": F, "G": G, "G_log": G * hp, "alpha": alpha, "hp": hp, "cond_est": lmax / lmin, "verify_factor": vf, "verify_inverse": vi}
    log(f"F={F!r} |G|={np.linalg.norm(G):.6e} cond_est={out['cond_est']:.3e} ({time.perf_counter() - t0:.0f}s)")
    md = o.GPRModel(COV, hp, x, y)
    if n > ob.BIG_N:           # the oracle's dtrsm calls go through the blocked solve as well (oracle/gpr_oracle_big.py)
        o.sl.solve_triangular = lambda a, b, trans=0, lower=False, **kw: ob._solve_ut(a, b, "T" if trans in ("T", 1) else "N")
    pc = ob.FactorCache(U, alpha)          # solve_triangular references the upper triangle only
    mu, var = o.predict(md, xp, diagonal_var=True, pc=pc, same=False)
    out["pred_mean"], out["pred_var"] = mu, var
    log(f"predict done ({time.perf_counter() - t0:.0f}s)")
    smu, svar = o.split_predict(md, o.Cmap(xe, xq), (1, 3), pc)
    out["split_mean_rows"] = smu[:3, :].copy()
    out["split_mean_sampled"] = smu.reshape(-1, order="F")[samp]
    out["split_var_rows"] = svar[:3 * nq].copy()
    log(f"split predict done ({time.perf_counter() - t0:.0f}s)")
    return out


if __name__ == "__main__":
    res = make()
    np.savez(os.path.join(HERE, "config3_n32768.npz"), **res)
    print("written", os.path.join(HERE, "config3_n32768.npz"))
