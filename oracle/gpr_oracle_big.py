"""CPU ORACLE, large-N companions -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Two things live here, both restating the same reference lines as oracle/gpr_oracle.py (paths relative to
/root/reference) but organised so that N = 32768 fits a 62 GB host:

* `nlml_grad_lean`  : the oracle's NLML + gradient (src/cost.jl:96-127, src/loss_grad.jl:39-52,
  src/deriv_covar.jl:20-29) with K, U and K^-1 as the only N x N arrays (2 x 8.6 GB at N = 32768).  Same LAPACK
  calls as the reference -- dpotrf('U') in place, dpotrs for alpha, dpotrs on the identity for K^-1 (here over
  column blocks of the identity: LAPACK solves right-hand-side columns independently) -- and the gradient terms
  alpha' dK alpha and <K^-1, dK> accumulated SEPARATELY over column blocks, dK recomputed per block from x exactly
  as grad! writes it.  Used to generate tests/golden/config3_n32768.npz; pinned to gpr_oracle.loss_grad in
  tests/test_oracle_big.py (<= 1e-12 on a well-conditioned model).

* `reference_shaped_eval` : the reference's evaluation AS IT EXECUTES IT: one
  materialised N x N matrix per component, their sum, dpotrf, dpotrs, dpotrs on a materialised identity, then for
  each hyper-parameter a materialised dK, dgemv, two ddots -- (nk + 3) N^2 doubles, nothing restructured, only
  numpy temporaries avoided (out= / in-place forms) so that the working set is what Julia's would be.  This is
  what bench.py's `--impl reference` and `cpu_baseline` time.  Per-stage wall times are returned.
"""
import math
import time

import numpy as np
import scipy.linalg as sl

import gpr_oracle as o

# OpenBLAS's dpotrf mis-factors the N = 32768 covariance of BASELINE.json config 3 in this image: it returns
# info = 16545 although every leading minor up to order 17000 factors cleanly with the same routine -- observed with
# BOTH bundled builds (scipy's LP64 0.3.31.dev and numpy's ILP64 0.3.30, 8 threads, SkylakeX kernels), so it is a
# library defect at that order, not an integer-width problem.  Above BIG_N the factorization and the solves are
# therefore BLOCKED at this level: LAPACK dpotrf / BLAS dtrsm on contiguous blocks of order <= NB_BIG and dgemm
# updates -- the textbook right-looking algorithm dpotrf itself implements -- and every large result is verified
# against K on sampled entries (verify_factor / verify_inverse_columns), which holds whatever the library does.
BIG_N = 16384
NB_BIG = 8192


def _potrf_upper(K):
    """dpotrf('U') in place (strict lower triangle untouched); returns (U, info)."""
    n = K.shape[0]
    if n <= BIG_N:
        return sl.lapack.dpotrf(K, lower=0, clean=0, overwrite_a=1)
    nb = NB_BIG
    for k0 in range(0, n, nb):
        k1 = min(n, k0 + nb)
        Akk = np.array(K[k0:k1, k0:k1], order="F")
        Ukk, info = sl.lapack.dpotrf(Akk, lower=0, clean=0, overwrite_a=1)
        if info > 0:
            return K, k0 + info
        iu = np.triu_indices(k1 - k0)
        K[k0:k1, k0:k1][iu] = Ukk[iu]                       # upper part only: the strict lower triangle keeps K
        if k1 < n:
            P = np.array(K[k0:k1, k1:], order="F")          # row panel, contiguous copy
            P = sl.blas.dtrsm(1.0, Ukk, P, side=0, lower=0, trans_a=1, diag=0, overwrite_b=1)   # U_kk^-T P
            K[k0:k1, k1:] = P
            for j0 in range(k1, n, nb):                     # trailing update, upper block triangle only
                j1 = min(n, j0 + nb)
                K[k1:j1, j0:j1] -= P[:, :j1 - k1].T @ P[:, j0 - k1:j1 - k1]
                # the update also touched the strict lower part of the diagonal block (rows > cols inside [j0, j1)):
                # restore it from the symmetric K so that "strict lower = K" holds (test/test_loss.jl:46)
                blk = K[j0:j1, j0:j1]
                il = np.tril_indices(j1 - j0, -1)
                blk[il] += (P[:, j0 - k1:j1 - k1].T @ P[:, j0 - k1:j1 - k1])[il]
    return K, 0


def _solve_ut(U, B, trans, inplace=False):
    """op(U)^-1 B for the upper-triangular part of U, trans in {'T', 'N'}; blocked above BIG_N (there `inplace` solves in
    B's own storage, which must be a Fortran-ordered float64 array)."""
    n = U.shape[0]
    if n <= BIG_N:
        return sl.solve_triangular(U, B, trans=trans, lower=False)
    nb = NB_BIG
    X = B if inplace else np.array(B, dtype=np.float64, order="F", copy=True)
    X2 = X.reshape(n, -1)
    blocks = [(k0, min(n, k0 + nb)) for k0 in range(0, n, nb)]
    if trans == "T":                                        # forward substitution with U^T
        for (k0, k1) in blocks:
            for (i0, i1) in blocks:
                if i0 >= k0:
                    break
                X2[k0:k1] -= U[i0:i1, k0:k1].T @ X2[i0:i1]
            Ukk = np.array(U[k0:k1, k0:k1], order="F")
            X2[k0:k1] = sl.blas.dtrsm(1.0, Ukk, np.array(X2[k0:k1], order="F"), side=0, lower=0, trans_a=1, diag=0)
    else:                                                   # back substitution with U
        for (k0, k1) in reversed(blocks):
            for (j0, j1) in blocks:
                if j0 <= k0:
                    continue
                X2[k0:k1] -= U[k0:k1, j0:j1] @ X2[j0:j1]
            Ukk = np.array(U[k0:k1, k0:k1], order="F")
            X2[k0:k1] = sl.blas.dtrsm(1.0, Ukk, np.array(X2[k0:k1], order="F"), side=0, lower=0, trans_a=0, diag=0)
    return X


def _potrs_upper(U, B):
    """dpotrs('U'): (U^T U)^-1 B."""
    if U.shape[0] <= BIG_N:
        X, info = sl.lapack.dpotrs(U, B, lower=0, overwrite_b=1)
        assert info == 0
        return X
    X = np.asfortranarray(B, dtype=np.float64)          # no copy for the Fortran-ordered workspaces the callers pass (dpotrs overwrites b too)
    return _solve_ut(U, _solve_ut(U, X, "T", inplace=True), "N", inplace=True)


def _columns_of_K(cov, hp, x, cols, eps=1e-8):
    """K[:, cols] of the self covariance (jitter and noise on the diagonal entries), via the plain oracle."""
    cols = np.asarray(cols)
    Kc = o.kernel(cov, hp, x, x[:, cols], same=False) if o.is_composed(cov) else o.kernel_single(cov, hp, x, x[:, cols], False)
    ks = o.as_list(cov)
    nse = sum(1 for k in ks if k != o.NOISE)
    add = nse * eps
    if o.NOISE in ks and o.is_composed(cov):
        dims = [o.dim_hp(k, x.shape[0]) for k in ks]
        add += o.split(hp, dims)[ks.index(o.NOISE)][0] ** 2
    Kc[cols, np.arange(len(cols))] += add
    return Kc


def verify_factor(U, cov, hp, x, ncols=48, seed=0):
    """max |(U^T U)[:, c] - K[:, c]| / max|K| over sampled columns c; only the upper triangle of U is referenced
    (the strict lower triangle of the buffer holds K)."""
    n = U.shape[0]
    cols = np.sort(np.random.default_rng(seed).choice(n, min(ncols, n), replace=False))
    Kc = _columns_of_K(cov, hp, x, cols)
    Uc = np.array(U[:, cols], order="F")
    Uc[np.arange(n)[:, None] > cols[None, :]] = 0.0          # column c of the factor: rows <= c
    acc = np.zeros((n, len(cols)))
    cidx = np.arange(n)[None, :]
    for k0 in range(0, n, 2048):                            # (U^T U)[:, c] = sum over row blocks of W^T U[rows, c]
        k1 = min(n, k0 + 2048)
        W = np.array(U[k0:k1, :])
        W[cidx < np.arange(k0, k1)[:, None]] = 0.0
        acc += W.T @ Uc[k0:k1]
    return float(np.abs(acc - Kc).max() / np.abs(Kc).max())


def verify_inverse_columns(Kinv, cov, hp, x, ncols=32, seed=1):
    """max |K[:, c]^T Kinv - e_c^T| over sampled columns c (uses the symmetry of K)."""
    n = Kinv.shape[0]
    cols = np.sort(np.random.default_rng(seed).choice(n, min(ncols, n), replace=False))
    Kc = _columns_of_K(cov, hp, x, cols)
    E = Kc.T @ Kinv
    E[np.arange(len(cols)), cols] -= 1.0
    return float(np.abs(E).max())


def _component_block(kind, hp, x, j0, j1, eps):
    """Columns j0:j1 of one component's self covariance, jitter on the diagonal entries (covariance.jl:49-58,85-95)."""
    ls = np.asarray(hp[1:], dtype=np.float64)
    xs = x * ls[:, None]
    Kb = o._kernel_impl(kind, hp, x, x[:, j0:j1]) if kind != o.SE else None
    if Kb is None:
        d = np.zeros((x.shape[1], j1 - j0))
        tmp = np.empty_like(d)
        for dd in range(x.shape[0]):
            np.subtract(xs[dd, :, None], xs[dd, None, j0:j1], out=tmp)
            np.multiply(tmp, tmp, out=tmp)
            d += tmp
        np.multiply(d, -1.0, out=d)
        np.exp(d, out=d)
        np.multiply(d, float(hp[0]) ** 2, out=d)
        Kb = d
    idx = np.arange(j0, j1)
    Kb[idx, idx - j0] += eps
    return Kb


def build_K(cov, hp, x, eps=1e-8, blk=1024, out=None):
    """kernel(cov, hp, x) (self form, compose_covar.jl:47-77) into a Fortran-ordered N x N array, column blocks."""
    N, dim = x.shape[1], x.shape[0]
    ks = o.as_list(cov)
    hps = o.split(hp, [o.dim_hp(k, dim) for k in ks])
    K = np.empty((N, N), order="F") if out is None else out
    for j0 in range(0, N, blk):
        j1 = min(N, j0 + blk)
        acc = None
        for kind, h in zip(ks, hps):
            if kind == o.NOISE:
                continue
            Kb = _component_block(kind, h, x, j0, j1, eps)
            acc = Kb if acc is None else acc + Kb
        if o.NOISE in ks:
            idx = np.arange(j0, j1)
            acc[idx, idx - j0] += hps[ks.index(o.NOISE)][0] ** 2
        K[:, j0:j1] = acc
    return K


def nlml_grad_lean(cov, hp, x, y, eps=1e-8, blk=1024, log=None):
    """(F, G, alpha, U, Kinv) with U = dpotrf('U') output (strict lower keeps K) and Kinv the full dpotrs(I) result."""
    say = log or (lambda *a: None)
    N, dim = x.shape[1], x.shape[0]
    hp = np.asarray(hp, dtype=np.float64)
    ks = o.as_list(cov)
    dims = [o.dim_hp(k, dim) for k in ks]
    hps = o.split(hp, dims)
    t0 = time.perf_counter()
    K = build_K(cov, hp, x, eps, blk)
    say(f"K built {time.perf_counter() - t0:.1f}s")
    t0 = time.perf_counter()
    U, info = _potrf_upper(K)                                                  # cost.jl:104
    if info > 0:
        raise np.linalg.LinAlgError(f"PosDefException({info})")
    say(f"dpotrf {time.perf_counter() - t0:.1f}s")
    alpha = _potrs_upper(U, np.array(y, dtype=np.float64, order="F")).reshape(y.shape)   # cost.jl:106
    F = 0.5 * (float(np.dot(y, alpha)) + 2.0 * float(np.sum(np.log(np.diag(U)))) + N * math.log(2.0 * math.pi))
    t0 = time.perf_counter()
    Kinv = np.empty((N, N), order="F")
    for j0 in range(0, N, 4 * blk):                                            # cost.jl:107-109, identity by column blocks
        j1 = min(N, j0 + 4 * blk)
        B = np.zeros((N, j1 - j0), order="F")
        B[np.arange(j0, j1), np.arange(j1 - j0)] = 1.0
        Kinv[:, j0:j1] = _potrs_upper(U, B)
    say(f"dpotrs(I) {time.perf_counter() - t0:.1f}s")
    t0 = time.perf_counter()
    P = len(hp)
    t1, t2 = np.zeros(P), np.zeros(P)              # alpha' dK alpha and <K^-1, dK>, loss_grad.jl:44-45
    off = np.concatenate([[0], np.cumsum(dims)])
    for j0 in range(0, N, blk):
        j1 = min(N, j0 + blk)
        aa = alpha[:, None] * alpha[None, j0:j1]
        Kib = Kinv[:, j0:j1]
        for c, (kind, h) in enumerate(zip(ks, hps)):
            if kind == o.NOISE:
                continue
            Kb = _component_block(kind, h, x, j0, j1, eps)
            for li in range(dims[c]):
                if li == 0:
                    dK = (2.0 / abs(h[0])) * Kb                                # deriv_covar.jl:23
                elif kind == o.SE:
                    diff2 = (x[li - 1, :, None] - x[li - 1, None, j0:j1]) ** 2
                    dK = -2.0 * h[li] * Kb * diff2                             # deriv_covar.jl:26
                else:
                    raise NotImplementedError("lean gradient: SquaredExp components only")
                t1[off[c] + li] += float(np.sum(aa * dK))
                t2[off[c] + li] += float(np.sum(Kib * dK))
    G = -0.5 * (t1 - t2)
    if o.NOISE in ks:                                                          # loss_grad.jl:49-52
        c = ks.index(o.NOISE)
        G[off[c]] = -0.5 * 2.0 * hps[c][0] * float(np.sum(alpha ** 2 - np.diag(Kinv)))
        for c2 in range(c + 1, len(ks)):
            if ks[c2] == o.NOISE:                                              # later noise terms: same UniformScaling form
                G[off[c2]] = -0.5 * 2.0 * hps[c2][0] * float(np.sum(alpha ** 2 - np.diag(Kinv)))
    say(f"gradient {time.perf_counter() - t0:.1f}s")
    return F, G, alpha, U, Kinv


class FactorCache:
    """Stand-in for gpr_oracle.GPRPredictCache built from an existing factor (same fields: U, wt, eps)."""

    def __init__(self, U, wt, eps=1e-8):
        self.U, self.wt, self.eps = U, wt, eps


# ----------------------------------------------------------------------------------------------------------------
def reference_shaped_eval(cov, log_hp, x, y, eps=1e-8, ws=None):
    """One log_loss_grad! (src/cost.jl:60-70) executed the way the reference executes it; returns (F, G, stage seconds,
    workspace).  Workspace (reused between calls like MllGradCache, caches/cost.jl:21-44): kerns[nk] N x N,
    kchol_base, dK, Kinv, tt."""
    N, dim = x.shape[1], x.shape[0]
    hp = np.exp(np.asarray(log_hp, dtype=np.float64))
    ks = o.as_list(cov)
    dims = [o.dim_hp(k, dim) for k in ks]
    hps = o.split(hp, dims)
    st = {}
    if ws is None:
        ws = {"kerns": [None if k == o.NOISE else np.empty((N, N), order="F") for k in ks],
              "kchol": np.empty((N, N), order="F"), "dK": np.empty((N, N), order="F"), "Kinv": np.empty((N, N), order="F"),
              "tt": np.empty(N)}
    kerns, kchol, dK, Kinv, tt = ws["kerns"], ws["kchol"], ws["dK"], ws["Kinv"], ws["tt"]
    row = np.empty(N)
    t0 = time.perf_counter()
    for c, (kind, h) in enumerate(zip(ks, hps)):                               # kernels!  compose_covar.jl:102-107
        if kind == o.NOISE:
            continue
        Kc = kerns[c]
        xs = x * h[1:, None]
        Kc[...] = 0.0
        for j in range(N):                                                     # one column at a time: fused distance + exp
            col = Kc[:, j]
            for dd in range(dim):
                np.subtract(xs[dd], xs[dd, j], out=row)
                np.multiply(row, row, out=row)
                col += row
            np.negative(col, out=col)
            np.exp(col, out=col)
            np.multiply(col, h[0] ** 2, out=col)
            col[j] += eps
    st["kbuild"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    first = True
    for c, kind in enumerate(ks):                                              # cost.jl:99-103
        if kind == o.NOISE:
            continue
        if first:
            np.copyto(kchol, kerns[c]); first = False
        else:
            np.add(kchol, kerns[c], out=kchol)
    if o.NOISE in ks:
        idx = np.arange(N)
        kchol[idx, idx] += hps[ks.index(o.NOISE)][0] ** 2
    st["sum"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    U, info = _potrf_upper(kchol)                                              # cost.jl:104
    if info > 0:
        raise np.linalg.LinAlgError(f"PosDefException({info})")
    st["potrf"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    alpha = _potrs_upper(U, np.array(y, dtype=np.float64, order="F"))          # cost.jl:106
    Kinv[...] = 0.0
    idx = np.arange(N)
    Kinv[idx, idx] = 1.0                                                       # cost.jl:107-108
    Ki = _potrs_upper(U, Kinv)                                                 # cost.jl:109 (N right-hand sides)
    st["potrs_identity"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    P = len(hp)
    G = np.zeros(P)
    off = np.concatenate([[0], np.cumsum(dims)])
    for c, (kind, h) in enumerate(zip(ks, hps)):                               # cost.jl:119-127
        for li in range(dims[c]):
            if kind == o.NOISE:
                G[off[c] + li] = -0.5 * 2.0 * h[0] * float(np.sum(alpha ** 2 - np.diag(Ki)))
                continue
            Kc = kerns[c]
            if li == 0:
                np.multiply(Kc, 2.0 / abs(h[0]), out=dK)                       # deriv_covar.jl:23
            else:
                xd = x[li - 1]
                for j in range(N):                                             # deriv_covar.jl:26, fused per column
                    np.subtract(xd, xd[j], out=row)
                    np.multiply(row, row, out=row)
                    np.multiply(row, -2.0 * h[li], out=row)
                    np.multiply(Kc[:, j], row, out=dK[:, j])
            np.dot(dK, alpha, out=tt)                                          # mul!(tt, dK, alpha)   dgemv
            # dot(K^-1, dK): ddot over the N^2 entries (both Fortran ordered: the flat views alias the buffers)
            G[off[c] + li] = -0.5 * (float(np.dot(tt, alpha)) - float(np.dot(Ki.reshape(-1, order="F"), dK.reshape(-1, order="F"))))
    st["gradient"] = time.perf_counter() - t0
    G *= hp                                                                    # cost.jl:65
    F = 0.5 * (float(np.dot(y, alpha)) + 2.0 * float(np.sum(np.log(np.diag(U)))) + N * math.log(2.0 * math.pi))
    return F, G, st, ws
