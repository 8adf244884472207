"""The large-N companions of the oracle (oracle/gpr_oracle_big.py) against the plain oracle on the same inputs:
`nlml_grad_lean` (generator of tests/golden/config3_n32768.npz and config5_n16384.npz) and `reference_shaped_eval`
(the reference-as-executed CPU arm that bench.py times).  Reference path: src/cost.jl:60-70,96-127."""
import os
import sys

import numpy as np
import pytest

import gpr_oracle as o
import gpr_oracle_big as ob

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("cov,D,N,blk", [((o.SE, o.SE, o.NOISE), 8, 1100, 256), ((o.SE, o.NOISE), 16, 700, 128), ((o.NOISE, o.SE), 3, 500, 512)])
def test_lean_matches_oracle(cov, D, N, blk):
    rng = np.random.default_rng(N)
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    hp = 0.3 + rng.random(o.dim_hp(cov, D))
    hp[[i for i, k in enumerate(np.repeat(cov, [o.dim_hp(k, D) for k in cov])) if k == o.NOISE]] = 0.1
    md = o.GPRModel(cov, hp, x, y)
    tc = o.MllGradCache(md)
    Fo, Go = o.loss_grad(hp, md, tc)
    F, G, alpha, U, Kinv = ob.nlml_grad_lean(cov, hp, x, y, blk=blk)
    assert abs(F - Fo) <= 1e-13 * abs(Fo)
    assert np.abs(G - Go).max() <= 1e-11 * np.abs(Go).max()
    assert np.abs(alpha - tc.alpha).max() <= 1e-12 * np.abs(tc.alpha).max()
    assert np.abs(Kinv - tc.Kinv).max() <= 1e-12 * np.abs(tc.Kinv).max()
    assert np.abs(U - tc.kchol_base).max() <= 1e-12 * np.abs(U).max()        # incl. the strict lower triangle (= K)


def test_reference_shaped_eval_matches_oracle():
    rng = np.random.default_rng(5)
    D, N = 8, 900
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
    cov = (o.SE, o.SE, o.NOISE)
    Fo, Go = o.log_loss_grad(np.log(hp), o.GPRModel(cov, hp, x, y))
    F, G, st, ws = ob.reference_shaped_eval(cov, np.log(hp), x, y)
    F2, G2, _, _ = ob.reference_shaped_eval(cov, np.log(hp * 1.01), x, y, ws=ws)      # workspace reuse
    F3, G3, _, _ = ob.reference_shaped_eval(cov, np.log(hp), x, y, ws=ws)
    assert abs(F - Fo) <= 1e-13 * abs(Fo) and F3 == F
    assert np.abs(G - Go).max() <= 1e-11 * np.abs(Go).max() and np.array_equal(G3, G)
    assert F2 != F
    assert set(st) == {"kbuild", "sum", "potrf", "potrs_identity", "gradient"}


def test_golden_generators_at_reduced_size():
    """make_golden_config3.make at a reduced size equals the plain oracle (the full-size fixture is produced by the
    same code path)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_config3 as m
    r = m.make(n=768, m=32, ne=12, nq=9, nsamp=15, log=lambda *a: None)
    x, y, hp = m.inputs(768)
    xp, xe, xq, samp = m.test_inputs(32, 12, 9, 15)
    md = o.GPRModel(m.COV, hp, x, y)
    F, G = o.loss_grad(hp, md)
    assert abs(F - r["F"]) <= 1e-13 * abs(F) and np.abs(G - r["G"]).max() <= 1e-11 * np.abs(G).max()
    mu, var = o.predict(md, xp, diagonal_var=True)
    assert np.abs(mu - r["pred_mean"]).max() <= 1e-11 and np.abs(var - r["pred_var"]).max() <= 1e-12
    smu, svar = o.split_predict(md, o.Cmap(xe, xq))
    assert np.abs(smu[:3] - r["split_mean_rows"]).max() <= 1e-11
    assert np.abs(smu.reshape(-1, order="F")[samp] - r["split_mean_sampled"]).max() <= 1e-11
    assert np.abs(svar[:27] - r["split_var_rows"]).max() <= 1e-12
