"""SURVEY.md 8f "next" rows -- host logic that needs no GPU: the BFGS-quad updater and finite-difference Hessian
(/root/reference/test/test_update.jl:1-46 ported line by line), the M-estimator losses and the k-fold splitter.
Run against BOTH the product's host mirror (gpr_sm100a.api) and the oracle restatement."""
import numpy as np
import pytest

import gpr_oracle as o


@pytest.fixture(scope="module")
def api():
    import gpr_sm100a
    return gpr_sm100a


def _impls(api):
    return [("api", api.hessian_fd, api.bfgs_quad, api.bfgs_hessian), ("oracle", o.hessian_fd, o.bfgs_quad, o.bfgs_hessian)]


@pytest.mark.parametrize("dim", list(range(2, 31, 4)))
def test_bfgs_quad_and_hessian_fd(api, dim):
    """test/test_update.jl:1-23"""
    rng = np.random.default_rng(dim)
    J = rng.random(dim)
    L = np.tril(rng.random((dim, dim)))
    H = L @ L.T + 1e-7 * np.eye(dim)
    x0 = rng.random(dim)
    jac = lambda x: J + H @ x
    xma = -np.linalg.solve(H, J)
    eps, max_iter = 1e-4, 100000
    for name, hessian_fd, bfgs_quad, _ in _impls(api):
        assert np.all(np.linalg.eigvalsh(H) > 0)
        np.testing.assert_allclose(hessian_fd(jac, rng.random(dim)), H, atol=eps, err_msg=name)
        xm, Jm, Hm, iters = bfgs_quad(x0, jac(x0), None, jac, eps, max_iter)
        assert iters < max_iter, name
        np.testing.assert_allclose(xm, xma, rtol=10 * eps, atol=10 * eps * np.abs(xma).max(), err_msg=name)
        assert np.linalg.norm(Jm) < eps, name


@pytest.mark.parametrize("dim", list(range(2, 31, 4)))
def test_bfgs_hessian_update(api, dim):
    """test/test_update.jl:25-46"""
    rng = np.random.default_rng(100 + dim)
    L = np.tril(rng.random((dim, dim)))
    B = L @ L.T + 1e-7 * np.eye(dim)
    s, t = rng.random(dim), rng.random(dim)
    p = 1.0 / float(s @ t)
    st, ts, ss = np.outer(s, t), np.outer(t, s), np.outer(s, s)
    for name, _, _, bfgs_hessian in _impls(api):
        Bs = bfgs_hessian(B, s, t)
        assert np.array_equal(Bs, Bs.T), name
        assert np.all(np.linalg.eigvalsh(Bs) > -1e-12 * np.abs(Bs).max()), name
        np.testing.assert_allclose(bfgs_hessian(B, s, t, 0.0), B, rtol=1e-14, err_msg=name)
        np.testing.assert_allclose(bfgs_hessian(None, s, t), (np.eye(dim) - p * (st + ts)) + (p ** 2 * float(t @ t) + p) * ss,
                                   rtol=1e-12, atol=1e-12, err_msg=name)


def test_hessian_fd_uses_batched_gradient(api):
    """hessian_fd! hands its P + 1 points to `many` when the gradient closure offers it (ReplicaGradient: one per GPU)."""
    rng = np.random.default_rng(3)
    H = np.diag(1.0 + rng.random(6))

    class Jac:
        calls = 0

        def __call__(self, x):
            return H @ x

        def many(self, pts):
            Jac.calls += 1
            return [H @ p for p in pts]

    np.testing.assert_allclose(api.hessian_fd(Jac(), rng.random(6)), H, atol=1e-6)
    assert Jac.calls == 1


def test_m_estimators_match_oracle(api):
    rng = np.random.default_rng(0)
    n = 40
    y, yp = rng.random(n), rng.random(n)
    A = rng.standard_normal((n, n))
    S = A @ A.T / n + np.eye(n)
    assert api.m_loss(api.MSE(), y, yp, S) == pytest.approx(o.loss_mse(y, yp), rel=1e-14)
    assert api.m_loss(api.ChiSq(), y, yp, S) == pytest.approx(o.loss_chisq(y, yp, S), rel=1e-14)
    assert api.m_loss(api.ChiSq(), y, yp, api.Diagonal(np.diag(S).copy())) == pytest.approx(o.loss_chisq(y, yp, S), rel=1e-14)
    assert api.m_loss(api.Mahalanobis(), y, yp, S) == pytest.approx(o.loss_mahalanobis(y, yp, S), rel=1e-12)
    d = y - yp
    assert o.loss_mahalanobis(y, yp, S) == pytest.approx(float(d @ np.linalg.solve(S, d)), rel=1e-10)


def test_kfoldcv_partition(api):
    """src/crossval.jl:1-12: nb = div(n, k) folds of k test points; train = the other n - k positions of the shuffle."""
    for fn in (api.kfoldcv, o.kfoldcv):
        trn, tst = fn(103, 10, rng=np.random.default_rng(5))
        assert len(trn) == len(tst) == 10
        seen = np.concatenate(tst)
        assert len(set(seen.tolist())) == 100
        for a, b in zip(trn, tst):
            assert len(b) == 10 and len(a) == 93
            assert not set(a.tolist()) & set(b.tolist())
            assert sorted(a.tolist() + b.tolist()) == list(range(103))
    a1, b1 = api.kfoldcv(50, 5, rng=np.random.default_rng(9))
    a2, b2 = o.kfoldcv(50, 5, rng=np.random.default_rng(9))
    assert all(np.array_equal(u, v) for u, v in zip(a1 + b1, a2 + b2))


def test_oracle_update_sample_reoptimises():
    """test/test_update.jl:50-76 on the oracle: after y += dy the quasi-Newton update drives the log-space gradient
    below eps_J in a few iterations when started from a converged model."""
    import scipy.optimize as so
    rng = np.random.default_rng(11)
    D, N = 3, 120
    x = rng.random((D, N))
    hp_true = np.concatenate([[1.0], 1.0 + rng.random(D), [0.05]])
    y = o.sample_mvn((o.SE, o.NOISE), hp_true, x, rng.standard_normal(N))
    md = o.GPRModel((o.SE, o.NOISE), np.ones(D + 2), x, y)
    tc = o.MllGradCache(md)
    res = so.minimize(lambda v: o.log_loss_grad(v, md, tc), np.zeros(D + 2), jac=True, method="L-BFGS-B", options={"gtol": 1e-6, "ftol": 1e-15})
    md.params[...] = np.exp(res.x)
    it = o.update_sample(md, 0.01 * y ** 2, eps_j=1e-3)
    _, G = o.log_loss_grad(np.log(md.params), md, want_f=False)
    assert np.linalg.norm(G) < 1e-3
    assert it < 10


def test_oracle_sample_mvn_covariance():
    """distributions.jl:20-35: s = L z + mu with L L^T = Sigma .+ 1e-7 (shift on every entry)."""
    rng = np.random.default_rng(2)
    x = rng.random((2, 30))
    hp = np.array([1.3, 0.7, 1.1, 0.2])
    S = o.kernel((o.SE, o.NOISE), hp, x) + 1e-7
    Z = np.eye(30)
    Lcols = np.stack([o.sample_mvn((o.SE, o.NOISE), hp, x, Z[:, j]) for j in range(30)], axis=1)
    np.testing.assert_allclose(Lcols @ Lcols.T, S, rtol=1e-12, atol=1e-14)
    assert np.allclose(np.triu(Lcols, 1), 0.0)
    mu = rng.random(30)
    np.testing.assert_allclose(o.sample_mvn((o.SE, o.NOISE), hp, x, np.zeros(30), mu), mu)


def test_gaussian_integrals_against_quadrature(api):
    """test/test_integrate.jl:3-21: closed forms vs numerical quadrature (scipy.integrate.quad for QuadGK)."""
    import scipy.integrate as si
    rng = np.random.default_rng(8)
    for _ in range(25):
        xs, w = 3.0 * rng.random(), 0.05 + 3.0 * rng.random()
        a, b = -3.0 + 6 * rng.random(), -3.0 + 6 * rng.random()
        q1 = si.quad(lambda t: np.exp(-w ** 2 * (t - xs) ** 2), a, b, epsrel=1e-10)[0]
        for gi, ei in ((api.gauss_integ, api.erf_integ), (o.gauss_integ, o.erf_integ)):
            assert float(gi(xs, w, a, b)) == pytest.approx(q1, rel=1e-8, abs=1e-12)
            q2 = si.quad(lambda t: float(gi(t, w, a, b)), a, b, epsrel=1e-10)[0]
            assert float(ei(w, a, b)) == pytest.approx(q2, abs=1e-5)


@pytest.mark.parametrize("dim,n", [(2, 100), (3, 300), (5, 500)])
def test_antiderivative_product_form(dim, n):
    """test/test_integrate.jl:23-37"""
    rng = np.random.default_rng(dim * n)
    xs = rng.random((dim, n))
    hp = 0.05 + 5.0 * rng.random(dim + 1)
    a = -2 + 4 * rng.random(dim)
    b = a + 2.0 * rng.random(dim)
    w = hp[1:dim + 1]
    integ = hp[0] ** 2 * np.prod(o.gauss_integ(xs, w[:, None], a[:, None], b[:, None]), axis=0)
    np.testing.assert_allclose(o.antideriv(xs, hp, a, b), integ, rtol=1e-12)


def test_oracle_integration_noise_paths_agree():
    """test/test_integrate.jl:118-161: zero sample noise reproduces the Cholesky path; vector noise matches per-column
    refactorization."""
    rng = np.random.default_rng(4)
    dim, n, k = 2, 120, 5
    x = rng.random((dim, n))
    y = rng.random((n, k))
    hp = np.array([1.0, 1.5, 0.8])
    md = o.GPRModel(o.SE, hp, x, y)
    a, b = np.zeros(dim), np.ones(dim)
    mu, var = o.integrate(md, hp, a, b, None, eps=1e-6)
    mu0, var0 = o.integrate(md, hp, a, b, np.zeros(k), eps=1e-6)
    np.testing.assert_allclose(mu0, mu, rtol=1e-5)
    np.testing.assert_allclose(var0, var[0], rtol=1e-4, atol=1e-9)
    noise = 1e-3 * (1 + rng.random(k))
    mun, varn = o.integrate(md, hp, a, b, noise)
    K = o.kernel_single(o.SE, hp, x, None, True, 1e-8)
    k1, k2 = o.antideriv(x, hp, a, b), o.antideriv2(hp, a, b)
    for i in range(k):
        Kn = K + noise[i] * np.eye(n)
        assert mun[i] == pytest.approx(float(np.linalg.solve(Kn, y[:, i]) @ k1), rel=1e-6)
        assert varn[i] == pytest.approx(k2 - float(k1 @ np.linalg.solve(Kn, k1)), rel=1e-5, abs=1e-10)
