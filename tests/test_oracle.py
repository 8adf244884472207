"""Pins the CPU oracle against every identity / known-answer test the reference's own suite holds for the
hot path (the reference ships no golden vectors).  Ported from
/root/reference/test/test_covariance.jl, test_loss.jl, test_models.jl, test_split_kernel.jl."""
import math

import numpy as np
import pytest

import gpr_oracle as o


def rng_for(*key):
    return np.random.default_rng(abs(hash(key)) % (2 ** 32))


# ---------------------------------------------------------------- test/test_covariance.jl
@pytest.mark.parametrize("n", [100, 200])
@pytest.mark.parametrize("dim", [1, 3, 7])
def test_covariance_structure_and_compose(n, dim):
    rng = np.random.default_rng(100 * n + dim)
    x, xp = rng.random((dim, n)), rng.random((dim, 2 * n))
    hp = rng.random(o.dim_hp(o.SE, dim))
    Kxx, Kxp = o.kernel(o.SE, hp, x), o.kernel(o.SE, hp, x, xp)
    assert np.array_equal(Kxx, Kxx.T)                          # :28
    assert np.all(np.linalg.eigvalsh(Kxx) > 0)                 # :29 (isposdef; jitter 1e-8)
    assert Kxx.shape == (n, n) and Kxp.shape == (n, 2 * n)     # :30-31
    I = np.eye(n)
    # :35-41
    k = (o.SE, o.NOISE)
    hps = rng.random(o.dim_hp(k, dim))
    np.testing.assert_allclose(o.kernel(k, hps, x), o.kernel(o.SE, hps[:-1], x) + hps[-1] ** 2 * I, rtol=1.5e-8)
    np.testing.assert_allclose(o.kernel(k, hps, x, xp), o.kernel(o.SE, hps[:-1], x, xp), rtol=1.5e-8)
    # :43-52
    k = (o.SE, o.SE)
    hps = rng.random(o.dim_hp(k, dim))
    np.testing.assert_allclose(o.kernel(k, hps, x), o.kernel(o.SE, hps[:dim + 1], x) + o.kernel(o.SE, hps[dim + 1:], x), rtol=1.5e-8)
    np.testing.assert_allclose(o.kernel(k, hps, x, xp), o.kernel(o.SE, hps[:dim + 1], x, xp) + o.kernel(o.SE, hps[dim + 1:], x, xp), rtol=1.5e-8)
    # :54-63
    k = (o.SE, o.SE, o.NOISE)
    hps = rng.random(o.dim_hp(k, dim))
    np.testing.assert_allclose(o.kernel(k, hps, x), o.kernel(o.SE, hps[:dim + 1], x) + o.kernel(o.SE, hps[dim + 1:-1], x) + hps[-1] ** 2 * I, rtol=1.5e-8)
    np.testing.assert_allclose(o.kernel(k, hps, x, xp), o.kernel(o.SE, hps[:dim + 1], x, xp) + o.kernel(o.SE, hps[dim + 1:-1], x, xp), rtol=1.5e-8)
    # :65-70 noise first
    k = (o.NOISE, o.SE)
    hps = rng.random(o.dim_hp(k, dim))
    np.testing.assert_allclose(o.kernel(k, hps, x), o.kernel(o.SE, hps[1:], x) + hps[0] ** 2 * I, rtol=1.5e-8)
    np.testing.assert_allclose(o.kernel(k, hps, x, xp), o.kernel(o.SE, hps[1:], x, xp), rtol=1.5e-8)
    # :72-81 noise in the middle
    k = (o.SE, o.NOISE, o.SE)
    hps = rng.random(o.dim_hp(k, dim))
    np.testing.assert_allclose(o.kernel(k, hps, x), o.kernel(o.SE, hps[:dim + 1], x) + o.kernel(o.SE, hps[dim + 2:], x) + hps[dim + 1] ** 2 * I, rtol=1.5e-8)
    np.testing.assert_allclose(o.kernel(k, hps, x, xp), o.kernel(o.SE, hps[:dim + 1], x, xp) + o.kernel(o.SE, hps[dim + 2:], x, xp), rtol=1.5e-8)


@pytest.mark.parametrize("dim", [1, 2, 5])
def test_kernel_grad_fd(dim):
    """:84-87 (the reference only checks i = dim; all i are checked here)"""
    rng = np.random.default_rng(dim)
    n = 100
    x = rng.random((dim, n))
    hp = rng.random(dim + 1)
    K = o.kernel(o.SE, hp, x)
    for i in range(1, dim + 2):
        hpe = hp.copy()
        hpe[i - 1] += 1e-7
        fd = (o.kernel(o.SE, hpe, x) - K) / 1e-7
        np.testing.assert_allclose(o.grad_kernel(o.SE, i, hp, x), fd, atol=1e-3)


def test_kernel_grad_compose_index_map():
    """:89-105"""
    rng = np.random.default_rng(7)
    dim = 2
    x = rng.random((dim, 100))
    k = (o.SE, o.NOISE, o.SE)
    hp = rng.random(o.dim_hp(k, dim))
    hps = o.split(hp, [dim + 1, 1, dim + 1])
    for i, (c, li) in {1: (0, 1), 2: (0, 2), 3: (0, 3), 5: (2, 1), 6: (2, 2), 7: (2, 3)}.items():
        np.testing.assert_allclose(o.grad_kernel(k, i, hp, x), o.grad_kernel(o.SE, li, hps[c], x), rtol=1.5e-8)
    g4 = o.grad_kernel(k, 4, hp, x)
    assert g4[0] == "I" and g4[1] == pytest.approx(2 * hps[1][0])
    assert o.find_idx([3, 1, 3], 4) == (2, 1)


# ---------------------------------------------------------------- test/test_loss.jl
@pytest.mark.parametrize("n", [10, 20, 100])
def test_nlml_known_answer_diagonal(n):
    """:1-11 closed form for a diagonal K"""
    rng = np.random.default_rng(n)
    x, y = rng.random(n) + 0.01, rng.random(n)
    mle = 0.5 * (np.dot(y, y / x) + np.sum(np.log(x)) + n * math.log(2 * math.pi))
    U = np.diag(np.sqrt(x))
    wt = y / x
    assert o.loss_from_chol(U, y, wt) == pytest.approx(mle, rel=1e-13)


@pytest.mark.parametrize("n", [10, 20, 100])
@pytest.mark.parametrize("dim", [2, 5])
@pytest.mark.parametrize("ny", [1, 5])
def test_grad_marginal_likelihood(n, dim, ny):
    """:21-56 (vector y) and :58-97 (matrix y, random train_axis)"""
    rng = np.random.default_rng(1000 * n + 10 * dim + ny)
    x = rng.random((dim, n))
    y1 = np.sin(x).sum(0)
    hp = rng.random(dim + 1)
    K = o.kernel(o.SE, hp, x)
    # :32 trace identity
    Kinv = np.linalg.inv(K)
    assert o.grad_term(K, y1, Kinv) == pytest.approx(-0.5 * (np.trace(np.outer(y1, y1) @ K) - n), rel=1e-5)

    if ny == 1:
        y, ta = y1, 1
    else:
        y = np.stack([rng.random() * y1 for _ in range(ny)], axis=1)
        ta = int(rng.integers(1, ny + 1))
    cov = (o.SE, o.NOISE)
    hp = rng.random(o.dim_hp(cov, dim))
    md = o.GPRModel(cov, hp, x, y, train_axis=ta)
    yt = o.get_sample(md)
    tc = o.MllGradCache(md)
    F, G = o.loss_grad(hp, md, tc)
    assert o.loss_functional(cov, hp, x, yt) == pytest.approx(F, rel=1.5e-8)      # :35
    Kf = o.kernel(cov, hp, x)
    import scipy.linalg as sl
    U = sl.cholesky(Kf, lower=False)
    np.testing.assert_allclose(np.triu(tc.kchol_base), U, rtol=1.5e-8, atol=1e-12)     # :46 upper = U
    np.testing.assert_allclose(np.tril(tc.kchol_base, -1), np.tril(Kf, -1), rtol=0, atol=0)   # strict lower keeps K
    np.testing.assert_allclose(tc.alpha, np.linalg.solve(Kf, yt), rtol=1e-6)       # :47
    np.testing.assert_allclose(tc.Kinv, np.linalg.inv(Kf), rtol=1e-6, atol=1e-6 * np.abs(np.linalg.inv(Kf)).max())  # :48
    for i in range(len(hp)):                                                       # :50-55
        hpe = hp.copy()
        hpe[i] += 1e-6
        fd = (o.loss_functional(cov, hpe, x, yt) - F) / 1e-6
        assert G[i] == pytest.approx(fd, rel=2e-3, abs=1e-4)
    # log-space chain rule (src/cost.jl:60-70)
    Fl, Gl = o.log_loss_grad(np.log(hp), md)
    assert Fl == pytest.approx(F, rel=1e-12)
    np.testing.assert_allclose(Gl, G * hp, rtol=1e-9)


# ---------------------------------------------------------------- test/test_models.jl
@pytest.mark.parametrize("n,npred,dim", [(100, 100, 1), (200, 100, 2), (100, 200, 5)])
def test_prediction(n, npred, dim):
    rng = np.random.default_rng(n + npred + dim)
    x, xp = rng.random((dim, n)), rng.random((dim, npred))
    y = np.sin(x.sum(0)) ** 2
    # :17-24 interpolation with SE+SE (jitter only)
    md2 = o.GPRModel((o.SE, o.SE), rng.random(2 * (dim + 1)), x, y)
    np.testing.assert_allclose(o.predict_mean(md2, x), y, rtol=1e-5, atol=1e-5)
    _, S = o.predict(md2, x)
    assert np.abs(S).max() < 1e-5
    # :26-31
    hp3 = rng.random(dim + 2)
    hp3[-1] = 1e-5
    md3 = o.GPRModel((o.SE, o.NOISE), hp3, x, y)
    np.testing.assert_allclose(o.predict_mean(md3, x), y, rtol=1e-3, atol=1e-3)
    # :34-48 diag variance == diag(full covariance)
    md = o.GPRModel((o.SE, o.NOISE), rng.random(dim + 2), x, y)
    _, Sf = o.predict(md, xp)
    _, Sd = o.predict(md, xp, diagonal_var=True)
    np.testing.assert_allclose(np.diag(Sf), Sd, atol=1e-5)
    md1 = o.GPRModel(o.SE, rng.random(dim + 1), x, y)
    _, Sf = o.predict(md1, xp)
    _, Sd = o.predict(md1, xp, diagonal_var=True)
    np.testing.assert_allclose(np.diag(Sf), Sd, atol=1e-5)


# ---------------------------------------------------------------- test/test_split_kernel.jl
COVS = [o.SE, (o.SE, o.NOISE), (o.SE, o.SE), (o.SE, o.SE, o.NOISE)]


@pytest.mark.parametrize("dim,n", [(2, 100), (3, 200), (5, 26)])
def test_split_kernel(dim, n):
    rng = np.random.default_rng(dim * 1000 + n)
    x = rng.random((dim, n))
    xe, xq = rng.random((dim, n + 7)), rng.random((dim, max(n - 10, 3)))
    cm = o.Cmap(xe, xq)
    flat = cm.flatten()
    ne, nq = xe.shape[1], xq.shape[1]
    # :10-22 flatten order: e fastest
    for e, q in ((0, 0), (3, 2), (ne - 1, nq - 1)):
        np.testing.assert_allclose(flat[:, q * ne + e], xe[:, e] + xq[:, q])
    # :24-33 D = DA + DB + DC
    DA, DB, DC = o.distance(xe, xq, "SplitA"), o.distance(xe, x, "Euclidean"), o.distance(x, xq, "SplitC")
    Dm = o.distance(flat, x)
    for e, q, s in ((0, 0, 0), (5, 1, n - 1), (ne - 1, nq - 1, 3)):
        assert Dm[q * ne + e, s] == pytest.approx(DA[e, q] + DB[e, s] + DC[s, q], rel=1e-12, abs=1e-13)
    # :36-44 split covariance == dense covariance
    for cov in COVS:
        hp = rng.random(o.dim_hp(cov, dim))
        KK = o.kernel(cov, hp, flat, x, same=False) if o.is_composed(cov) else o.kernel_single(cov, hp, flat, x, False)
        A, B, C = o.split_kernel(cov, hp, cm, x)
        for e, q, s in ((0, 0, n - 1), (5, 1, n - 1), (ne - 1, nq - 1, 0)):
            assert np.sum(A[e, q, :] * B[e, s, :] * C[s, q, :]) == pytest.approx(KK[q * ne + e, s], rel=1e-10)


@pytest.mark.parametrize("cov", COVS)
@pytest.mark.parametrize("dim,n,e,q", [(1, 100, 10, 10), (2, 200, 20, 30), (5, 100, 50, 10)])
def test_split_prediction(cov, dim, n, e, q):
    """:47-78 mean == dense mean; variance == dense variance on the first 3*q entries (q-fast order) and not beyond"""
    rng = np.random.default_rng(n + e + q + dim)
    x, xe, xq = rng.random((dim, n)), rng.random((dim, e)), rng.random((dim, q))
    y = np.sin(x.sum(0)) ** 2
    hp = 0.2 + rng.random(o.dim_hp(cov, dim))
    md = o.GPRModel(cov, hp, x, y)
    cm = o.Cmap(xe, xq)
    yp, _ = o.predict(md, cm.flatten(), diagonal_var=True)
    yps, varps = o.split_predict(md, cm)
    np.testing.assert_allclose(yps.reshape(-1, order="F"), yp, rtol=1e-6, atol=1e-9)
    _, varpt = o.predict(md, o.Cmap(xq, xe).flatten(), diagonal_var=True)
    np.testing.assert_allclose(varps[:3 * q], varpt[:3 * q], rtol=1e-5, atol=1e-9)
    assert not np.allclose(varps[:3 * q + 1], varpt[:3 * q + 1], rtol=1e-5, atol=0)
