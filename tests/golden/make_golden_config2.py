"""Oracle NLML + gradient for BASELINE.json config 2 (ARD SE + noise, N=8192, D=8, three hyper-parameter sets;
SURVEY.md 8d) plus the "cond-limited" set D (plain SquaredExp(), jitter only, hyper-parameters of set A without the
noise term).  Inputs are regenerated from the seed, only F and G are stored (config2_n8192.npz).
Takes a few minutes of CPU:  python tests/golden/make_golden_config2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import gpr_oracle as o  # noqa: E402

D, N, SEED = 8, 8192, 2002


def inputs():
    rng = np.random.default_rng(SEED)
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    sets = {"A": np.concatenate([[1.0], 0.5 * np.ones(D), [0.1]]),
            "B": np.concatenate([[1.5], np.linspace(0.3, 1.2, D), [0.05]]),
            "C": np.concatenate([[0.7], 2.0 * np.ones(D), [0.3]])}
    return x, y, sets


if __name__ == "__main__":
    x, y, sets = inputs()
    out = {}
    for name, hp in sets.items():
        md = o.GPRModel((o.SE, o.NOISE), hp, x, y)
        F, G = o.loss_grad(hp, md)
        out["F_" + name], out["G_" + name], out["hp_" + name] = F, G, hp
        print(name, F, np.linalg.norm(G), flush=True)
    # set D: jitter-only SquaredExp(); cond(K) from lambda_max (power iteration) / lambda_min (inverse iteration)
    hpD = sets["A"][:-1]
    md = o.GPRModel(o.SE, hpD, x, y)
    tc = o.MllGradCache(md)
    F, G = o.loss_grad(hpD, md, tc)
    v = np.ones(N) / np.sqrt(N)
    K = o.kernel(o.SE, hpD, x)
    for _ in range(50):
        v = K @ v
        lmax = np.linalg.norm(v)
        v /= lmax
    w = np.random.default_rng(0).standard_normal(N)
    for _ in range(50):
        w = tc.Kinv @ w
        linv = np.linalg.norm(w)
        w /= linv
    out["F_D"], out["G_D"], out["hp_D"], out["cond_D"] = F, G, hpD, lmax * linv
    print("D", F, np.linalg.norm(G), "cond", lmax * linv, flush=True)
    np.savez(os.path.join(HERE, "config2_n8192.npz"), **out)
