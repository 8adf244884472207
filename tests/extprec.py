"""Extended-precision ground truth for the hot path -- TEST INFRASTRUCTURE ONLY.

The same formulas as oracle/gpr_oracle.py (hence the same reference lines: src/covariance.jl:85-95,
src/compose_covar.jl:47-77, src/deriv_covar.jl:20-32, src/loss_grad.jl:39-52, src/predict.jl:73-95) evaluated in
numpy.longdouble (x87 80-bit: 64-bit significand, eps = 1.08e-19) with hand-written Cholesky / triangular solves --
no BLAS, no LAPACK, nothing shared with either the oracle's OpenBLAS or the CUDA library.  With cond(K) up to ~1e10
its own error is ~cond * 1e-19 = 1e-9 relative, three orders of magnitude below what two FP64 factorizations can
agree on (cond * 2.2e-16), so it can referee between the oracle and the GPU.  N <= 512 (O(N^3) numpy loops).
"""
import numpy as np

import gpr_oracle as o

LD = np.longdouble


def _kernel_ld(kind, hp, x, xp):
    ls = hp[1:].astype(LD)
    xs, xps = x.astype(LD) * ls[:, None], xp.astype(LD) * ls[:, None]
    d = np.zeros((x.shape[1], xp.shape[1]), dtype=LD)
    for dd in range(x.shape[0]):
        d += (xs[dd, :, None] - xps[dd, None, :]) ** 2
    if kind == o.SE:
        return LD(hp[0]) ** 2 * np.exp(-d)
    raise ValueError(kind)


def chol_upper(K):
    """U upper with U^T U = K (right-looking, row by row)."""
    A = K.copy()
    n = A.shape[0]
    for k in range(n):
        if not A[k, k] > 0:
            raise np.linalg.LinAlgError(f"PosDefException({k + 1})")
        A[k, k] = np.sqrt(A[k, k])
        A[k, k + 1:] /= A[k, k]
        A[k + 1:, k + 1:] -= np.outer(A[k, k + 1:], A[k, k + 1:])
    return np.triu(A)


def solve_ut(U, B):
    """U^-T B (forward substitution)."""
    X = np.array(B, dtype=LD, copy=True)
    n = U.shape[0]
    for k in range(n):
        X[k] = X[k] / U[k, k]
        if k + 1 < n:
            X[k + 1:] -= np.multiply.outer(U[k, k + 1:], X[k]) if X.ndim == 2 else U[k, k + 1:] * X[k]
    return X


def solve_u(U, B):
    """U^-1 B (back substitution)."""
    X = np.array(B, dtype=LD, copy=True)
    n = U.shape[0]
    for k in range(n - 1, -1, -1):
        X[k] = X[k] / U[k, k]
        if k > 0:
            X[:k] -= np.multiply.outer(U[:k, k], X[k]) if X.ndim == 2 else U[:k, k] * X[k]
    return X


def truth(cov, hp, x, y, xp=None, eps=1e-8):
    """dict with F, G, alpha, Kinv, (pred_mean, pred_var) in longdouble."""
    hp = np.asarray(hp, dtype=np.float64)
    dim, n = x.shape
    ks = o.as_list(cov)
    dims = [o.dim_hp(k, dim) for k in ks]
    hps = o.split(hp, dims)
    comps = []
    K = np.zeros((n, n), dtype=LD)
    idx = np.arange(n)
    for kind, h in zip(ks, hps):
        if kind == o.NOISE:
            comps.append(None)
            continue
        Kc = _kernel_ld(kind, h, x, x)
        Kc[idx, idx] += LD(eps)
        comps.append(Kc)
        K += Kc
    if o.NOISE in ks and len(ks) > 1:
        K[idx, idx] += LD(hps[ks.index(o.NOISE)][0]) ** 2
    U = chol_upper(K)
    yl = y.astype(LD)
    alpha = solve_u(U, solve_ut(U, yl))
    W = solve_u(U, np.eye(n, dtype=LD))              # U^-1
    Kinv = W @ W.T
    F = LD(0.5) * (yl @ alpha + 2 * np.sum(np.log(np.diag(U))) + n * np.log(2 * LD(np.pi)))
    G = []
    M = np.multiply.outer(alpha, alpha) - Kinv
    for c, (kind, h) in enumerate(zip(ks, hps)):
        for li in range(dims[c]):
            if kind == o.NOISE:
                G.append(LD(-0.5) * 2 * LD(h[0]) * np.sum(alpha ** 2 - np.diag(Kinv)))
            elif li == 0:
                G.append(LD(-0.5) * np.sum(M * comps[c]) * (2 / abs(LD(h[0]))))
            else:
                xd = x[li - 1].astype(LD)
                G.append(LD(-0.5) * np.sum(M * comps[c] * (xd[:, None] - xd[None, :]) ** 2) * (-2 * LD(h[li])))
    out = {"F": F, "G": np.array(G, dtype=LD), "alpha": alpha, "Kinv": Kinv, "U": U, "K": K}
    if xp is not None:
        Ks = np.zeros((xp.shape[1], n), dtype=LD)
        prior = LD(0)
        for kind, h in zip(ks, hps):
            prior += LD(h[0]) ** 2
            if kind != o.NOISE:
                Ks += _kernel_ld(kind, h, xp, x)
        V = solve_ut(U, Ks.T)                         # U^-T K*^T
        out["pred_mean"] = Ks @ alpha
        out["pred_var"] = prior - np.sum(V * V, axis=0)
    return out


def cond2(K):
    return float(np.linalg.cond(np.asarray(K, dtype=np.float64)))
