"""Oracle NLML + gradient + alpha for the parity leg of BASELINE.json config 5: the model of config 5
(ARD SquaredExp()+WhiteNoise(), D = 16, sigma = 1, l = 0.4, sigma_n = 0.1, seed 5005) at N = 16384, the size at which
the CPU oracle is still feasible; the distributed GPU path is checked against it for G = 1, 2, 4, 8 ranks.
Arithmetic: oracle/gpr_oracle_big.nlml_grad_lean (same LAPACK calls as the reference: dpotrf, dpotrs, dpotrs on the
identity).  ~5 min on 8 cores:  python tests/golden/make_golden_config5.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import gpr_oracle as o  # noqa: E402
import gpr_oracle_big as ob  # noqa: E402

D, N, SEED = 16, 16384, 5005


def inputs(n=N):
    rng = np.random.default_rng(SEED)
    x = rng.random((D, n))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(n)
    hp = np.concatenate([[1.0], 0.4 * np.ones(D), [0.1]])
    return x, y, hp


if __name__ == "__main__":
    x, y, hp = inputs()
    F, G, alpha, U, Kinv = ob.nlml_grad_lean((o.SE, o.NOISE), hp, x, y, log=print)
    print(F, np.linalg.norm(G))
    np.savez(os.path.join(HERE, "config5_n16384.npz"), F=F, G=G, alpha=alpha, hp=hp)
