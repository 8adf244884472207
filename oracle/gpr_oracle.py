"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Op-for-op restatement, in numpy + scipy.linalg (LAPACK dpotrf/dpotrs/dtrsm, BLAS dgemv/ddot/dgemm),
of the hot path of srinix007/GaussianProcessRegression.jl.  Every function cites the reference lines it
follows (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this module, and only as the checker / the reported baseline.

Pinning status: the reference cannot run in this image (no julia binary, un-vendored LazyTensors.jl;
SURVEY.md M7) and it ships no golden vectors -- its tests are identities on random inputs.  This oracle is
therefore pinned against every identity / known-answer test the reference's own suite holds for this path
(tests/test_oracle.py ports test/test_covariance.jl, test_loss.jl, test_models.jl, test_split_kernel.jl),
and frozen golden vectors generated from it live in tests/golden/.  Third-party arithmetic it stands in
for: OpenBLAS 0.3.20 (reference pin, Manifest.toml:363-366; here scipy's OpenBLAS 0.3.30, same LAPACK
routines), LazyTensors.jl 0.1.0 @9d4a87fc (lazy broadcast + sum!(D, expr, 1), Manifest.toml:256-262),
Julia Base.exp (<= 1 ulp).  The Matern-5/2 component does not exist in the reference: parity UNPINNED
for that component, formulas in SURVEY.md A.1.

Array conventions follow Julia: x is (D, N), y is (N,) or (N, ny), K(x, xp) is (N, M); hyper-parameter
index arguments `i` and `train_axis` are 1-based exactly as in the reference.
"""
import math

import numpy as np
import scipy.linalg as sl

SE = "SquaredExp"
NOISE = "WhiteNoise"
MATERN52 = "Matern52"   # extension, not in the reference

_SQRT5 = math.sqrt(5.0)


# --------------------------------------------------------------------------- covariance.jl / compose_covar.jl
def as_list(cov):
    """A covariance is a component tag or a tuple/list of tags (ComposedKernel, compose_covar.jl:1-19)."""
    return [cov] if isinstance(cov, str) else list(cov)


def is_composed(cov):
    return not isinstance(cov, str)


def dim_hp(cov, dim):
    """covariance.jl:27,60 ; compose_covar.jl:26-28"""
    return sum(1 if k == NOISE else dim + 1 for k in as_list(cov))


def split(hp, dims):
    """compose_covar.jl:21-24"""
    cind = np.concatenate([[0], np.cumsum(dims)])
    return [np.asarray(hp[cind[i]:cind[i + 1]], dtype=np.float64) for i in range(len(dims))]


def distance(x, xp, metric="Euclidean"):
    """distance!(::Euclidean...) covariance.jl:72-77 ; SplitDistanceA/C split_kernel.jl:111-123.
    Direct difference, reduced over the coordinate axis (LazyTensors sum!(D, expr, 1))."""
    N, M = x.shape[1], xp.shape[1]
    D = np.zeros((N, M))
    for d in range(x.shape[0]):
        a = x[d, :, None]
        b = xp[d, None, :]
        if metric == "Euclidean":
            D += (a - b) ** 2
        elif metric == "SplitA":      # xe, xq :  xq^2 + 2 xe xq
            D += b ** 2 + 2.0 * a * b
        elif metric == "SplitC":      # xs, xq : -2 xs xq
            D += -2.0 * a * b
        else:
            raise ValueError(metric)
    return D


def _kernel_impl(kind, hp, x, xp, metric="Euclidean"):
    """kernel_impl!(::SquaredExp...) covariance.jl:85-95: xs = x .* l ; K = sigma^2 exp(-d)."""
    ls = np.asarray(hp[1:], dtype=np.float64)
    sig = float(hp[0])
    xs, xps = x * ls[:, None], xp * ls[:, None]
    d = distance(xs, xps, metric)
    if kind == SE:
        return sig ** 2 * np.exp(-1.0 * d)
    if kind == MATERN52:   # extension
        r = np.sqrt(np.maximum(d, 0.0))
        return sig ** 2 * (1.0 + _SQRT5 * r + (5.0 / 3.0) * d) * np.exp(-_SQRT5 * r)
    raise ValueError(kind)


def kernel_single(kind, hp, x, xp=None, same=None, eps=1e-8, metric="Euclidean"):
    """kernel / kernel! for one component, covariance.jl:29-64.  `same` is the reference's x === xp."""
    if xp is None:
        xp, same = x, True
    if same is None:
        same = xp is x
    if kind == NOISE:
        raise ValueError("WhiteNoise has no dense cross-covariance (covariance.jl:61-64)")
    K = _kernel_impl(kind, hp, x, xp, metric)
    if same:
        idx = np.arange(K.shape[0])
        K[idx, idx] += eps
    return K


def rm_noise(cov, hps):
    """compose_covar.jl:30-33"""
    ks = as_list(cov)
    return [k for k in ks if k != NOISE], [h for k, h in zip(ks, hps) if k != NOISE]


def kernel(cov, hp, x, xp=None, same=None, eps=1e-8):
    """kernel(K, hp, x[, xp]): covariance.jl:29-47 (single), compose_covar.jl:35-77 (composed).
    Self form (xp None): sum of components (each adds its own eps) + sigma_n^2 on the diagonal.
    Cross form: no noise; eps per component iff same."""
    if not is_composed(cov):
        if cov == NOISE:
            raise ValueError("bare WhiteNoise model is not a dense kernel")
        return kernel_single(cov, hp, x, xp, same, eps)
    self_form = xp is None
    if self_form:
        xp, same = x, True
    elif same is None:
        same = xp is x
    dim = x.shape[0]
    ks = as_list(cov)
    hps = split(hp, [dim_hp(k, dim) for k in ks])
    Ks, hpn = rm_noise(cov, hps)
    K = kernel_single(Ks[0], hpn[0], x, xp, same, eps)
    for t in range(1, len(Ks)):
        K = K + kernel_single(Ks[t], hpn[t], x, xp, same, eps)
    if self_form and NOISE in ks:    # add_noise!, compose_covar.jl:63-71 (first WhiteNoise only)
        nidx = ks.index(NOISE)
        idx = np.arange(K.shape[0])
        K[idx, idx] += hps[nidx][0] ** 2
    return K


def kernels(cov, hp, x, eps=1e-8):
    """kernels / kernels!: per-component self covariances, noise slot = 1x1 zero (compose_covar.jl:80-107)."""
    if not is_composed(cov):
        return [kernel(cov, hp, x, eps=eps)]
    dim = x.shape[0]
    ks = as_list(cov)
    hps = split(hp, [dim_hp(k, dim) for k in ks])
    return [np.zeros((1, 1)) if k == NOISE else kernel_single(k, h, x, eps=eps) for k, h in zip(ks, hps)]


def find_idx(dims, i):
    """compose_covar.jl:109-115 (1-based i -> (component 1-based, local 1-based))"""
    cd = np.cumsum(dims)
    kidx = int(np.argmax(cd >= i)) + 1 if np.any(cd >= i) else 0
    hpidx = i if kidx == 1 else i - cd[kidx - 2]
    return kidx, int(hpidx)


# --------------------------------------------------------------------------- deriv_covar.jl
def grad_single(kind, i, hp, x, K):
    """grad!(::SquaredExp, DK, i, hp, x, K) deriv_covar.jl:20-29 ; WhiteNoise :31-32 (returns ('I', lambda)).
    i is 1-based; K is the component's self covariance INCLUDING its jitter."""
    if kind == NOISE:
        return ("I", 2.0 * hp[0])
    if kind == SE:
        if i == 1:
            return (2.0 / abs(hp[0])) * K
        d = i - 2
        diff2 = (x[d, :, None] - x[d, None, :]) ** 2
        return -2.0 * hp[i - 1] * K * diff2
    if kind == MATERN52:   # extension
        if i == 1:
            return (2.0 / abs(hp[0])) * K
        ls = np.asarray(hp[1:])
        r = np.sqrt(np.maximum(distance(x * ls[:, None], x * ls[:, None]), 0.0))
        d = i - 2
        diff2 = (x[d, :, None] - x[d, None, :]) ** 2
        return -(5.0 / 3.0) * hp[0] ** 2 * (1.0 + _SQRT5 * r) * np.exp(-_SQRT5 * r) * hp[i - 1] * diff2
    raise ValueError(kind)


def grad_kernel(cov, i, hp, x, Ks=None, eps=1e-8):
    """grad(cov, i, hp, x[, K]) deriv_covar.jl:2-18 ; composed dispatch compose_covar.jl:117-123."""
    if Ks is None:
        Ks = kernels(cov, hp, x, eps)
    if not is_composed(cov):
        return grad_single(cov, i, hp, x, Ks[0])
    dim = x.shape[0]
    ks = as_list(cov)
    dims = [dim_hp(k, dim) for k in ks]
    hps = split(hp, dims)
    kidx, hpidx = find_idx(dims, i)
    return grad_single(ks[kidx - 1], hpidx, hps[kidx - 1], x, Ks[kidx - 1])


# --------------------------------------------------------------------------- models.jl
class GPRModel:
    """models.jl:17-45"""

    def __init__(self, cov, hp, x, y, train_axis=1):
        if len(hp) != dim_hp(cov, x.shape[0]):
            raise ValueError("Parameter size mismatch.")
        if x.shape[1] != y.shape[0]:
            raise ValueError("x and y size mismatch.")
        self.covar, self.params, self.x, self.y, self.train_axis = cov, np.asarray(hp, dtype=np.float64), x, y, train_axis


def get_sample(md):
    return md.y[:, md.train_axis - 1] if md.y.ndim == 2 else md.y


# --------------------------------------------------------------------------- loss_grad.jl / cost.jl
def loss_from_chol(U, y, alpha):
    """loss(::MarginalLikelihood, kchol, y, K^-1 y) loss_grad.jl:39-41"""
    n = U.shape[0]
    logdet = 2.0 * np.sum(np.log(np.diag(U)))
    return 0.5 * (float(np.dot(y, alpha)) + logdet + n * math.log(2.0 * math.pi))


def loss_functional(cov, hp, x, y, eps=1e-8):
    """loss(::MarginalLikelihood, cov, hp, x, y) loss_grad.jl:32-37"""
    K = kernel(cov, hp, x, eps=eps)
    U = sl.cholesky(K, lower=False)
    wt = sl.cho_solve((U, False), y)
    return loss_from_chol(U, y, wt)


def grad_term(dK, alpha, Kinv):
    """grad(::MarginalLikelihood, kchol, dK, alpha, K^-1, tt) loss_grad.jl:43-52"""
    if isinstance(dK, tuple):    # UniformScaling
        return -0.5 * dK[1] * float(np.sum(alpha ** 2 - np.diag(Kinv)))
    tt = dK @ alpha                                    # mul!(tt, dK, alpha)  (dgemv)
    gr = float(np.dot(tt, alpha)) - float(np.vdot(Kinv, dK))   # dot(K^-1, dK) (ddot over N^2)
    return -0.5 * gr


class MllGradCache:
    """caches/cost.jl:21-44 + update_cache! cost.jl:83-111"""

    def __init__(self, md):
        self.md = md
        self.hp = np.array(md.params, dtype=np.float64)
        self.kerns = None
        self.kchol_base = None
        self.alpha = None
        self.Kinv = None
        self.want_inverse = True


class MllLossCache(MllGradCache):
    """caches/cost.jl:6-19 + update_cache! cost.jl:74-81"""

    def __init__(self, md):
        super().__init__(md)
        self.want_inverse = False


def update_cache(tc, hp, md, eps=1e-8):
    tc.hp = np.array(hp, dtype=np.float64)
    tc.kerns = kernels(md.covar, tc.hp, md.x, eps)
    dense = [k for k in tc.kerns if k.shape != (1, 1) or md.x.shape[1] == 1]
    K = dense[0].copy()
    for t in range(1, len(dense)):
        K += dense[t]
    if is_composed(md.covar) and NOISE in as_list(md.covar):   # add_noise!
        ks = as_list(md.covar)
        hps = split(tc.hp, [dim_hp(k, md.x.shape[0]) for k in ks])
        idx = np.arange(K.shape[0])
        K[idx, idx] += hps[ks.index(NOISE)][0] ** 2
    # cholesky!(Hermitian(kchol_base)) = dpotrf('U'): upper <- U, strict lower keeps K (test_loss.jl:46)
    U, info = sl.lapack.dpotrf(np.asfortranarray(K), lower=0, clean=0, overwrite_a=0)
    if info > 0:
        raise np.linalg.LinAlgError(f"PosDefException({info})")
    tc.kchol_base = U
    y = get_sample(md)
    tc.alpha = sl.cho_solve((np.triu(U), False), y)
    if tc.want_inverse:
        n = K.shape[0]
        tc.Kinv, _ = sl.lapack.dpotrs(U, np.eye(n, order="F"), lower=0)   # ldiv!(kchol, I)  cost.jl:90-92
    return tc


def loss(md, tc):
    """loss(::MarginalLikelihood, md, tc) cost.jl:113-117"""
    return loss_from_chol(np.triu(tc.kchol_base), get_sample(md), tc.alpha)


def grad(md, tc):
    """grad!(dL, ::MarginalLikelihood, md, tc) cost.jl:119-127"""
    P = len(tc.hp)
    G = np.zeros(P)
    for i in range(1, P + 1):
        dK = grad_kernel(md.covar, i, tc.hp, md.x, tc.kerns)
        G[i - 1] = grad_term(dK, tc.alpha, tc.Kinv)
    return G


def loss_grad(hp, md, tc=None, want_f=True, want_g=True, eps=1e-8):
    """loss_grad! cost.jl:50-58"""
    tc = tc or (MllGradCache(md) if want_g else MllLossCache(md))
    update_cache(tc, hp, md, eps)
    G = grad(md, tc) if want_g else None
    F = loss(md, tc) if want_f else None
    return F, G


def log_loss_grad(log_hp, md, tc=None, want_f=True, want_g=True, eps=1e-8):
    """log_loss_grad! cost.jl:60-70: hp = exp.(log_hp); G .*= hp"""
    hp = np.exp(np.asarray(log_hp, dtype=np.float64))
    F, G = loss_grad(hp, md, tc, want_f, want_g, eps)
    if G is not None:
        G = G * hp
    return F, G


def islog(md):
    """cost.jl:4-8"""
    return SE in as_list(md.covar) or MATERN52 in as_list(md.covar)


# --------------------------------------------------------------------------- predict.jl
class GPRPredictCache:
    """caches/predict.jl:3-31 + update_cache!(pc, md) predict.jl:29-34"""

    def __init__(self, md, eps=1e-8):
        K = kernel(md.covar, md.params, md.x, eps=eps)
        self.U = sl.cholesky(K, lower=False)
        self.wt = sl.cho_solve((self.U, False), md.y)    # all columns of y
        self.eps = eps


def predict_mean(md, xp, pc=None, same=None):
    """predict_mean! predict.jl:36-40,73-76"""
    pc = pc or GPRPredictCache(md)
    if same is None:
        same = xp is md.x
    Kxp = kernel(md.covar, md.params, xp, md.x, same=same, eps=pc.eps) if is_composed(md.covar) else \
        kernel_single(md.covar, md.params, xp, md.x, same, pc.eps)
    return Kxp @ pc.wt


def prior_diag(md):
    """predict.jl:55-58 (composed: sum over ALL components incl. noise) ; :67 (plain)"""
    if not is_composed(md.covar):
        return md.params[0] ** 2
    dim = md.x.shape[0]
    hps = split(md.params, [dim_hp(k, dim) for k in as_list(md.covar)])
    return sum(h[0] ** 2 for h in hps)


def predict(md, xp, diagonal_var=False, pc=None, same=None):
    """predict / predict! predict.jl:14-25,42-102.  Returns (mean, Sigma) with Sigma (M,M) or the diagonal (M,)."""
    pc = pc or GPRPredictCache(md)
    if same is None:
        same = xp is md.x
    Kxp = kernel(md.covar, md.params, xp, md.x, same=same, eps=pc.eps) if is_composed(md.covar) else \
        kernel_single(md.covar, md.params, xp, md.x, same, pc.eps)
    mu = Kxp @ pc.wt
    V = sl.solve_triangular(pc.U, Kxp.T, trans="T", lower=False).T     # rdiv!(Kxp, U): V = Kxp U^-1
    if diagonal_var:
        return mu, prior_diag(md) - np.sum(V * V, axis=1)
    Sig = kernel(md.covar, md.params, xp, eps=pc.eps)                   # incl. jitter and noise (predict.jl:45)
    return mu, Sig - V @ V.T


# --------------------------------------------------------------------------- split_kernel.jl / split_predict.jl
class Cmap:
    """split_kernel.jl:1-17: lazy op.(xe[:, e], xq[:, q]); flatten order e fastest."""

    def __init__(self, xe, xq):
        self.xe, self.xq = xe, xq

    def flatten(self):
        D, ne, nq = self.xe.shape[0], self.xe.shape[1], self.xq.shape[1]
        out = np.empty((D, ne * nq))
        for q in range(nq):
            out[:, q * ne:(q + 1) * ne] = self.xe + self.xq[:, q:q + 1]
        return out


def split_kernel(cov, hp, xeq, x):
    """kernel!(Kxps::SplitKernel, cov, hp, xp::Cmap, x) split_kernel.jl:137-159 -> A (ne,nq,k), B (ne,N,k), C (N,nq,k)"""
    dim = x.shape[0]
    ks = as_list(cov)
    hps = split(hp, [dim_hp(k, dim) for k in ks])
    Ks, hpn = rm_noise(ks, hps)
    ne, nq, N = xeq.xe.shape[1], xeq.xq.shape[1], x.shape[1]
    A, B, C = np.empty((ne, nq, len(Ks))), np.empty((ne, N, len(Ks))), np.empty((N, nq, len(Ks)))
    for k, (kind, h) in enumerate(zip(Ks, hpn)):
        hs = np.array(h, dtype=np.float64)
        hs[0] = 1.0
        A[:, :, k] = kernel_single(kind, hs, xeq.xe, xeq.xq, same=False, metric="SplitA")
        B[:, :, k] = kernel_single(kind, hs, xeq.xe, x, same=False, metric="Euclidean")
        C[:, :, k] = kernel_single(kind, h, x, xeq.xq, same=False, metric="SplitC")
    return A, B, C


def split_predict(md, xeq, var_range=(1, 3), pc=None):
    """predict!(mu, Sigma::Diagonal, md, xeq::Cmap, pc) predict.jl:51-71 + split_predict.jl:10-19,39-53.
    Returns mu (ne, nq) and the variance vector (ne*nq, q fastest within e)."""
    pc = pc or GPRPredictCache(md)
    A, B, C = split_kernel(md.covar, md.params, xeq, md.x)
    ne, nq = A.shape[0], A.shape[1]
    mu = np.zeros((ne, nq))
    for k in range(A.shape[2]):
        Cw = pc.wt[:, None] * C[:, :, k]       # mul!(Cw, Diagonal(wt), C[:, :, k])
        BCw = B[:, :, k] @ Cw                   # mul!(BCw, B[:, :, k], Cw)
        mu += BCw * A[:, :, k]
    var = np.full(ne * nq, prior_diag(md))
    for e in range(var_range[0], var_range[1] + 1):
        Kxq = np.zeros((nq, md.x.shape[1]))
        for k in range(A.shape[2]):
            Kxq += A[e - 1, :, None, k] * B[e - 1, None, :, k] * C[:, :, k].T
        V = sl.solve_triangular(pc.U, Kxq.T, trans="T", lower=False).T
        var[(e - 1) * nq:e * nq] -= np.sum(V * V, axis=1)
    return mu, var


# =========================================================================== SURVEY.md 8f "next" rows (callers of the hot path)
# --------------------------------------------------------------------------- loss_grad.jl:11-30 (M-estimators used by crossval.jl)
def loss_mse(y, yp, Sp=None):
    """loss(::MSE, y, yp, Sigma_p) loss_grad.jl:11-14"""
    return float(np.sum((y - yp) ** 2) / len(y))


def loss_chisq(y, yp, Sp):
    """loss(::ChiSq, ...) loss_grad.jl:16-22: sum (y_i - yp_i)^2 / Sigma_p[i, i]"""
    return float(np.sum((y - yp) ** 2 / np.diag(Sp)))


def loss_mahalanobis(y, yp, Sp):
    """loss(::Mahalanobis, ...) loss_grad.jl:24-29: |L^-1 (y - yp)|^2 with Sigma_p = L L^T"""
    d = sl.solve_triangular(sl.cholesky(Sp, lower=True), y - yp, lower=True)
    return float(d @ d)


# --------------------------------------------------------------------------- crossval.jl
def kfoldcv(n, k, nb=None, rng=None):
    """kfoldcv(n, k, nb = div(n, k)) crossval.jl:1-12: nb folds of k test points each out of one shuffle of 1:n
    (1-based point labels in the reference; 0-based indices here).  Note the reference's training set of fold i:
    every entry nsh[j] whose POSITION j is not in the fold's position range."""
    nb = n // k if nb is None else nb
    nsh = (rng or np.random.default_rng()).permutation(n)
    trn, tst = [], []
    for i in range(nb):
        idx = np.arange(i * k, (i + 1) * k)
        tst.append(nsh[idx])
        trn.append(np.delete(nsh, idx))
    return trn, tst


def cv_step(md, cost, xtr, ytr, xtst, ytst):
    """cv_step / cv_step! crossval.jl:37-50: refit the cache on the training part, dense predictive covariance on the
    test part, M-estimator loss."""
    mdt = GPRModel(md.covar, md.params, xtr, ytr)
    yp, Sp = predict(mdt, xtst, diagonal_var=False)
    return cost(ytst, yp, Sp)


def cv_batch(md, cost, x, y, cvset):
    """cv_batch crossval.jl:14-35"""
    trn, tst = cvset
    return np.array([cv_step(md, cost, x[:, trn[i]], y[trn[i]], x[:, tst[i]], y[tst[i]]) for i in range(len(trn))])


# --------------------------------------------------------------------------- update_model.jl
def bfgs_hessian(Bi, s, t, rho=None):
    """bfgs_hessian update_model.jl:50-54 (inverse-Hessian BFGS update)"""
    rho = 1.0 / float(s @ t) if rho is None else rho
    n = len(s)
    Bi = np.eye(n) if Bi is None else Bi
    C = np.eye(n) - rho * np.outer(s, t)
    B = C @ Bi @ C.T + rho * np.outer(s, s)
    return 0.5 * (B + B.T)


def bfgs_quad_(theta, JJ, B, gradL, eps, max_iter=100):
    """bfgs_quad! update_model.jl:64-79: quasi-Newton steps with unit step length; theta, JJ, B updated in place"""
    it = 0
    while np.linalg.norm(JJ) > eps and it < max_iter:
        s = theta.copy()
        t = JJ.copy()
        theta -= B @ JJ
        JJ[...] = gradL(theta)
        s = theta - s
        t = JJ - t
        B[...] = bfgs_hessian(B, s, t)
        it += 1
    return it


def bfgs_quad(xx, JJ, HH, jac, eps=1e-5, max_iter=100):
    """bfgs_quad update_model.jl:56-62.  HH: start Hessian (None = identity, the reference passes `I`)."""
    n = len(xx)
    x0, J0 = np.array(xx, dtype=np.float64), np.array(JJ, dtype=np.float64)
    B = np.eye(n) if HH is None else np.linalg.inv(HH)
    it = bfgs_quad_(x0, J0, B, jac, eps, max_iter)
    return x0, J0, np.linalg.inv(B), it


def hessian_fd(gradL, x, eps=1e-6):
    """hessian_fd / hessian_fd! update_model.jl:81-94: forward differences of the gradient, column by column"""
    n = len(x)
    H = np.empty((n, n))
    g0 = gradL(x)
    for i in range(n):
        xe = np.array(x, dtype=np.float64)
        xe[i] += eps
        H[:, i] = (gradL(xe) - g0) / eps
    return H


def update_sample(md, dy, eps_j=1e-3):
    """update_sample!(md, dy, BFGSQuad(), MarginalLikelihood(), eps_J) update_model.jl:6-48: y += dy, gradient and
    finite-difference Hessian of the NLML in log space at the current optimum, quasi-Newton re-optimisation.
    Returns the number of iterations; md.y and md.params are updated in place."""
    md.y += dy
    tc = MllGradCache(md)
    jac = lambda v: log_loss_grad(v, md, tc, want_f=False)[1]
    hp = np.log(md.params)
    J = jac(hp)
    hess = hessian_fd(jac, hp)
    Binv = np.linalg.inv(0.5 * (hess + hess.T))      # inv(Hermitian(hess)): Hermitian() reads the upper triangle
    it = bfgs_quad_(hp, J, Binv, jac, eps_j)
    md.params[...] = np.exp(hp)
    return it


# --------------------------------------------------------------------------- distributions.jl
def sample_mvn(cov, hp, x, z, mu=None, shift=1e-7, eps=1e-8):
    """sample(gp(x, theta)) distributions.jl:20-45: Sigma = kernel(cov, theta, x); cholesky(Sigma .+ 1e-7) -- the shift
    is broadcast over EVERY entry --; s = L z + mu.  z: standard-normal draws (the reference draws them from
    Xoshiro(1); they are an input here)."""
    S = (kernel(cov, hp, x, eps=eps) if is_composed(cov) else kernel_single(cov, hp, x, None, True, eps)) + shift
    L = sl.cholesky(S, lower=True)
    out = L @ z
    return out if mu is None else out + mu


# --------------------------------------------------------------------------- integrate.jl
_RT_PI_BY_2 = math.sqrt(math.pi) * 0.5


def erf2(x, y):
    """SpecialFunctions.erf(x, y) = erf(y) - erf(x), evaluated without cancellation (erfc differences when both
    arguments lie on the same side), as SpecialFunctions.jl does."""
    import scipy.special as sp
    x, y = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    r = 1.0 / math.sqrt(2.0)
    out = sp.erf(y) - sp.erf(x)
    pos = (x >= 0) & (y >= 0) & ~((np.abs(x) <= r) & (np.abs(y) <= r))
    neg = (x <= 0) & (y <= 0) & ~((np.abs(x) <= r) & (np.abs(y) <= r))
    out = np.where(pos, sp.erfc(x) - sp.erfc(y), out)
    out = np.where(neg, sp.erfc(-y) - sp.erfc(-x), out)
    return out


def gauss_integ(xs, w, a, b):
    """gauss_integ(xs, w, a, b) integrate.jl:4-5: integral of exp(-w^2 (x - xs)^2) over [a, b]"""
    return (1.0 / w) * _RT_PI_BY_2 * erf2(w * (a - xs), w * (b - xs))


def erf_integ(w, a, b):
    """erf_integ integrate.jl:6-7: integral over [a, b] of gauss_integ(x, w, a, b)"""
    import scipy.special as sp
    return 1.0 / (w ** 2) * (np.exp(-(w * (b - a)) ** 2) - 1.0) + 2.0 * (_RT_PI_BY_2 / w) * (b - a) * sp.erf(w * (b - a))


def antideriv(xs, hp, a, b):
    """antideriv!(integ, SquaredExp(), xs, hp, a, b) integrate.jl:15-31"""
    nl = xs.shape[0]
    ls = np.asarray(hp[1:nl + 1])
    prefac = hp[0] ** 2 * _RT_PI_BY_2 ** nl * np.prod(1.0 / ls)
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return prefac * np.prod(erf2(ls[:, None] * (a[:, None] - xs), ls[:, None] * (b[:, None] - xs)), axis=0)


def antideriv2(hp, a, b):
    """antideriv2(SquaredExp(), hp, a, b) integrate.jl:33-41"""
    dim = len(a)
    ls = np.asarray(hp[1:dim + 1])
    return float(np.prod(erf_integ(ls, np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))) * hp[0] ** 2)


def integrate(md, hp, a, b, sample_noise=None, eps=1e-8):
    """integrate(md, hp, a, b; sample_noise) integrate.jl:50-62,103-160.  Returns (Iout[ny], var_Iout).
    sample_noise None: Cholesky path, var_Iout[0] only (:131-136).  scalar / vector noise: eigen path (:72-80,145-160)."""
    K = kernel(md.covar, hp, md.x, eps=eps) if is_composed(md.covar) else kernel_single(md.covar, hp, md.x, None, True, eps)
    Y = md.y.reshape(md.y.shape[0], -1)
    k1 = antideriv(md.x, hp, a, b)
    k2 = antideriv2(hp, a, b)
    if sample_noise is None:
        U = sl.cholesky(K, lower=False)
        wt = sl.cho_solve((U, False), Y)
        tt = sl.solve_triangular(U, k1, trans="T", lower=False)
        return wt.T @ k1, np.array([k2 - float(tt @ tt)])
    lam, P = np.linalg.eigh(K)
    noise = np.atleast_1d(np.asarray(sample_noise, dtype=np.float64))
    Dm = 1.0 / (lam[:, None] + noise[None, :])                  # inverse_diagonal_update!  (:82-100)
    wt = P @ (Dm * (P.T @ Y))
    t2 = (P.T @ k1) ** 2
    return wt.T @ k1, k2 - (1.0 / (lam[None, :] + noise[:, None])) @ t2
