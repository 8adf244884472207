"""Developer GPU check: runs each section in its own process and prints detailed errors.
Usage (on a GPU box):  python tools/gpu_check.py [section ...]
Not part of the product or the test-suite; tests/ holds the real parity tests."""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def sec_gemm():
    from gpr_sm100a import _ffi
    ctx = _ffi.get_context()
    rng = np.random.default_rng(0)
    for (tA, tB) in (("T", "N"), ("N", "N"), ("N", "T")):
        for (M, N, K) in ((128, 128, 16), (256, 384, 128), (384, 256, 272)):
            A = rng.standard_normal((K, M) if tA == "T" else (M, K))
            B = rng.standard_normal((N, K) if tB == "T" else (K, N))
            C0 = rng.standard_normal((M, N))
            opA = A.T if tA == "T" else A
            opB = B.T if tB == "T" else B
            for (alpha, beta) in ((1.0, 0.0), (-1.0, 1.0), (0.5, 2.0)):
                C, _ = _ffi.dbg_dgemm(ctx, tA, tB, alpha, A, B, beta, C0)
                ref = alpha * opA @ opB + beta * C0
                e = relerr(C, ref)
                flag = "OK " if e < 1e-13 else "BAD"
                print(f"gemm {tA}{tB} {M}x{N}x{K} a={alpha} b={beta}: relerr {e:.2e} {flag}")
                if e >= 1e-13:
                    bad = np.argwhere(np.abs(C - ref) > 1e-10 * np.abs(ref).max())
                    print("   first bad idx:", bad[:8].tolist(), "count", len(bad), "of", M * N)
        # upper-only
        M = N = 384
        K = 128
        A = rng.standard_normal((K, M) if tA == "T" else (M, K))
        B = rng.standard_normal((N, K) if tB == "T" else (K, N))
        C0 = rng.standard_normal((M, N))
        opA = A.T if tA == "T" else A
        opB = B.T if tB == "T" else B
        C, _ = _ffi.dbg_dgemm(ctx, tA, tB, -1.0, A, B, 1.0, C0, flags=1)
        ref = C0 - opA @ opB
        eu = relerr(np.triu(C), np.triu(ref))
        el = float(np.abs(np.tril(C, -1) - np.tril(C0, -1)).max())
        print(f"gemm {tA}{tB} upper-only: upper relerr {eu:.2e}, lower untouched diff {el:.2e}", "OK" if eu < 1e-13 and el == 0 else "BAD")
        if not (eu < 1e-13 and el == 0):
            badu = np.argwhere(np.abs(np.triu(C) - np.triu(ref)) > 1e-10)
            badl = np.argwhere(np.abs(np.tril(C, -1) - np.tril(C0, -1)) > 0)
            print("   bad upper:", len(badu), badu[:6].tolist(), " bad lower:", len(badl), badl[:6].tolist())
            if len(badu):
                i, j = badu[0]
                print("   C, ref, C0, full-update at first bad upper:", C[i, j], ref[i, j], C0[i, j])


def sec_gemm_cfgs():
    """Correctness + throughput of the three GEMM tile configurations (option gemm_cfg)."""
    from gpr_sm100a import _ffi
    ctx = _ffi.get_context()
    rng = np.random.default_rng(1)
    n = 4096
    Abig = np.asfortranarray(rng.standard_normal((n, n)))
    Bbig = np.asfortranarray(rng.standard_normal((n, n)))
    Czero = np.zeros((n, n), order="F")
    for cfg in (1, 2, 3):
        ctx.set_option("gemm_cfg", cfg)
        worst = 0.0
        for (tA, tB) in (("T", "N"), ("N", "N"), ("N", "T")):
            M, N, K = 384, 256, 272
            A = rng.standard_normal((K, M) if tA == "T" else (M, K))
            B = rng.standard_normal((N, K) if tB == "T" else (K, N))
            C0 = rng.standard_normal((M, N))
            opA = A.T if tA == "T" else A
            opB = B.T if tB == "T" else B
            C, _ = _ffi.dbg_dgemm(ctx, tA, tB, 0.5, A, B, 2.0, C0)
            worst = max(worst, relerr(C, 0.5 * opA @ opB + 2.0 * C0))
            # upper-only on a square diagonal-anchored C
            M = N = 384
            A = rng.standard_normal((K, M) if tA == "T" else (M, K))
            B = rng.standard_normal((N, K) if tB == "T" else (K, N))
            C0 = rng.standard_normal((M, N))
            opA = A.T if tA == "T" else A
            opB = B.T if tB == "T" else B
            C, _ = _ffi.dbg_dgemm(ctx, tA, tB, -1.0, A, B, 1.0, C0, flags=1)
            ref = C0 - opA @ opB
            worst = max(worst, relerr(np.triu(C), np.triu(ref)), float(np.abs(np.tril(C, -1) - np.tril(C0, -1)).max()))
        # triangular product W W^T (UPPER_ONLY | K_FROM_N), W upper triangular with zero lower part
        m = 512
        W = np.triu(rng.standard_normal((m, m)))
        C0 = rng.standard_normal((m, m))
        C, _ = _ffi.dbg_dgemm(ctx, "N", "T", 1.0, W, W, 0.0, C0, flags=3)
        worst = max(worst, relerr(np.triu(C), np.triu(W @ W.T)), float(np.abs(np.tril(C, -1) - np.tril(C0, -1)).max()))
        print(f"gemm cfg {cfg}: worst error over forms/flags {worst:.2e}", "OK" if worst < 1e-12 else "BAD")
        for (tA, tB) in (("T", "N"), ("N", "N"), ("N", "T")):
            _, ms = _ffi.dbg_dgemm(ctx, tA, tB, 1.0, Abig, Bbig, 1.0, Czero, reps=5)
            print(f"   cfg {cfg} {tA}{tB} {n}^3: {ms:.3f} ms {2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
        for k in (128, 512):
            _, ms = _ffi.dbg_dgemm(ctx, "T", "N", -1.0, np.asfortranarray(Abig[:k, :]), np.asfortranarray(Bbig[:k, :]), 1.0, Czero, reps=5)
            print(f"   cfg {cfg} TN {n}x{n}x{k}: {ms:.3f} ms {2 * n * n * k / ms / 1e9:.2f} TFLOP/s")
        _, ms = _ffi.dbg_dgemm(ctx, "T", "N", 1.0, np.asfortranarray(Abig[:128, :128]), np.asfortranarray(Bbig[:128, :2048]), 0.0, np.zeros((128, 2048), order="F"), reps=5)
        print(f"   cfg {cfg} leaf 128x2048x128: {ms * 1e3:.1f} us")
        K = rng.standard_normal((n, 64))
        K = K @ K.T / 64 + np.eye(n)
        for mode in (0, 3):
            A, ms = _ffi.dbg_factor(ctx, K, mode)
            print(f"   cfg {cfg} factor n={n} mode={mode}: {ms:.1f} ms")
    ctx.set_option("gemm_cfg", 0)


def sec_gemm_perf():
    from gpr_sm100a import _ffi
    import torch
    ctx = _ffi.get_context()
    rng = np.random.default_rng(0)
    n = 8192
    A = rng.standard_normal((n, n))
    B = rng.standard_normal((n, n))
    C0 = np.zeros((n, n))
    for (tA, tB) in (("T", "N"), ("N", "N"), ("N", "T")):
        C, ms = _ffi.dbg_dgemm(ctx, tA, tB, 1.0, A, B, 0.0, C0, reps=6)
        print(f"dgemm128 {tA}{tB} {n}^3: {ms:.2f} ms  {2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
    for k in (128, 256, 512, 1024):
        Ak = np.ascontiguousarray(A[:k, :]) if True else None
        C, ms = _ffi.dbg_dgemm(ctx, "T", "N", -1.0, np.asfortranarray(A[:k, :]), np.asfortranarray(B[:k, :]), 1.0, C0, reps=6)
        print(f"dgemm128 TN {n}x{n}x{k} (rank-k update): {ms:.3f} ms  {2 * n * n * k / ms / 1e9:.2f} TFLOP/s")
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        c = a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS dgemm (torch.matmul fp64) {n}^3 burst: {best:.2f} ms {2 * n ** 3 / best / 1e9:.2f} TFLOP/s")
    t0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    it = 0
    while it < 60:
        c = a @ b
        it += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    print(f"cuBLAS dgemm sustained ({it} back-to-back): {ms:.2f} ms {2 * n ** 3 / ms / 1e9:.2f} TFLOP/s")
    import torch.linalg
    for nn in (8192, 16384):
        X = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
        S = X @ X.T + nn * torch.eye(nn, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        for _ in range(2):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            L = torch.linalg.cholesky(S)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"cuSOLVER potrf (torch.linalg.cholesky) N={nn}: {ms:.1f} ms {nn ** 3 / 3 / ms / 1e9:.2f} TFLOP/s  [context only]")
        del X, S, L


def sec_factor():
    from gpr_sm100a import _ffi
    import scipy.linalg as sl
    ctx = _ffi.get_context()
    rng = np.random.default_rng(0)
    for n in (128, 256, 384, 1000, 1537):
        X = rng.standard_normal((n, n))
        K = X @ X.T / n + np.eye(n)
        U = sl.cholesky(K, lower=False)
        refs = [U, np.linalg.inv(U), np.linalg.inv(K), np.linalg.inv(K)]
        for mode in (0, 1, 2, 3):
            A, ms = _ffi.dbg_factor(ctx, K, mode)
            e = relerr(np.triu(A), np.triu(refs[mode]))
            low = float(np.abs(np.tril(A, -1) - np.tril(K, -1)).max())
            print(f"factor n={n} mode={mode}: upper relerr {e:.2e} lower-untouched {low:.1e} ({ms:.2f} ms)", "OK" if e < 1e-11 and low == 0 else "BAD")
    # not positive definite -> info
    K = np.eye(300)
    K[200, 200] = -1.0
    try:
        _ffi.dbg_factor(ctx, K, 0)
        print("posdef check: BAD (no exception)")
    except _ffi.PosDefException as ex:
        print("posdef check: info =", ex.info, "OK" if ex.info == 201 else "BAD")


def sec_factor_perf():
    from gpr_sm100a import _ffi
    ctx = _ffi.get_context()
    rng = np.random.default_rng(0)
    for n in (4096, 8192, 16384):
        X = rng.standard_normal((n, 64))
        K = X @ X.T / 64 + np.eye(n)
        for mode in (0, 2, 3):
            c0 = ctx.launch_count()
            A, ms = _ffi.dbg_factor(ctx, K, mode)
            fl = n ** 3 / 3 if mode == 0 else n ** 3
            print(f"factor perf n={n} mode={mode}: {ms:.1f} ms  {fl / ms / 1e9:.2f} TFLOP/s  launches {ctx.launch_count() - c0}")
        if n == 4096:
            r = np.abs(np.triu(A) @ K - 0).max()
            Ai = np.triu(A) + np.triu(A, 1).T
            print("   inverse check |Ainv K - I| =", float(np.abs(Ai @ K - np.eye(n)).max()))


def _model(cov_o, N, D, seed, ny=1, noise=0.1):
    import gpr_oracle as o
    rng = np.random.default_rng(seed)
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + noise * rng.standard_normal(N)
    if ny > 1:
        y = np.stack([y * (0.5 + 0.3 * k) for k in range(ny)], axis=1)
    return x, y


TYPES = {"SquaredExp": 1, "WhiteNoise": 2, "Matern52": 3}


def sec_kernel():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    ctx = _ffi.get_context()
    rng = np.random.default_rng(3)
    for cov in ((o.SE,), (o.SE, o.NOISE), (o.SE, o.SE, o.NOISE), (o.NOISE, o.SE), (o.SE, o.NOISE, o.SE), (o.SE, o.MATERN52, o.NOISE)):
        for (D, N, M) in ((1, 100, 200), (3, 77, 130), (8, 300, 65)):
            x = rng.random((D, N))
            xp = rng.random((D, M))
            hp = 0.2 + rng.random(o.dim_hp(cov, D))
            types = [TYPES[c] for c in cov]
            K = _ffi.kernel_matrix(ctx, types, D, hp, x, x, True, 1e-8, True)
            Ko = o.kernel(cov, hp, x) if len(cov) > 1 else o.kernel(cov[0], hp, x)
            Kc = _ffi.kernel_matrix(ctx, types, D, hp, x, xp, False, 1e-8, False)
            Kco = o.kernel(cov, hp, x, xp) if len(cov) > 1 else o.kernel(cov[0], hp, x, xp)
            e1 = float(np.abs(K / Ko - 1).max())
            e2 = float(np.abs(Kc / Kco - 1).max())
            print(f"kernel {cov} D={D} N={N} M={M}: self max rel {e1:.2e} cross max rel {e2:.2e}", "OK" if max(e1, e2) < 1e-10 else "BAD")


def sec_nlml():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    ctx = _ffi.get_context()
    for cov, N, D, ny in (((o.SE, o.NOISE), 100, 2, 1), ((o.SE, o.NOISE), 300, 5, 1), ((o.SE, o.SE, o.NOISE), 1000, 8, 1),
                          ((o.SE, o.NOISE), 257, 3, 4), ((o.SE, o.MATERN52, o.NOISE), 500, 4, 1), ((o.SE,), 200, 2, 1),
                          ((o.SE, o.SE, o.NOISE), 2048, 8, 1)):
        x, y = _model(cov, N, D, 11, ny)
        rng = np.random.default_rng(5)
        hp = 0.3 + rng.random(o.dim_hp(cov, D))
        if o.NOISE in cov:
            hp[-1] = 0.1 if cov[-1] == o.NOISE else hp[-1]
        ta = 2 if ny > 1 else 1
        covo = cov if len(cov) > 1 else cov[0]
        md = o.GPRModel(covo, hp, x, y, train_axis=ta)
        tc = o.MllGradCache(md)
        Fo, Go = o.loss_grad(hp, md, tc)
        types = [TYPES[c] for c in cov]
        mh = _ffi.ModelHandle(ctx, types, D, x, y, train_axis=ta)
        F, G = mh.nlml_grad(hp)
        eF = abs(F - Fo) / abs(Fo)
        eG = float(np.abs(G - Go).max() / max(np.abs(Go).max(), 1e-300))
        eGi = float((np.abs(G - Go) / np.maximum(np.abs(Go), 1e-8 * np.linalg.norm(Go))).max())
        U = mh.fetch(_ffi.FETCH_U)
        al = mh.fetch(_ffi.FETCH_ALPHA)
        Ki = mh.fetch(_ffi.FETCH_KINV)
        eU = relerr(U, tc.kchol_base)
        ea = relerr(al, tc.alpha)
        eK = relerr(Ki, tc.Kinv)
        cond = np.linalg.cond(np.triu(tc.kchol_base)) ** 2
        Fl, Gl = mh.nlml_grad(np.log(hp), log_scale=True)
        eL = float(np.abs(Gl - Go * hp).max() / np.abs(Go * hp).max())
        ok = eF < 1e-8 and eGi < 1e-8 and eU < 1e-9 and ea < 1e-8 and eK < 1e-8
        print(f"nlml {cov} N={N} D={D} ny={ny}: relF {eF:.1e} relG(max) {eG:.1e} relG(comp) {eGi:.1e} U {eU:.1e} alpha {ea:.1e} Kinv {eK:.1e} logG {eL:.1e} cond {cond:.1e}", "OK" if ok else "BAD")
        mh.close()


def sec_predict():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    ctx = _ffi.get_context()
    ctx.set_option("predict_tile", 256)
    for cov, N, D, M, ny in (((o.SE, o.NOISE), 200, 2, 150, 1), ((o.SE, o.SE, o.NOISE), 500, 5, 700, 1), ((o.SE,), 300, 3, 100, 1),
                             ((o.SE, o.NOISE), 300, 3, 333, 3)):
        x, y = _model(cov, N, D, 21, ny)
        rng = np.random.default_rng(6)
        xp = rng.random((D, M))
        hp = 0.3 + rng.random(o.dim_hp(cov, D))
        covo = cov if len(cov) > 1 else cov[0]
        md = o.GPRModel(covo, hp, x, y)
        pc = o.GPRPredictCache(md)
        mu_o, var_o = o.predict(md, xp, diagonal_var=True, pc=pc)
        _, cov_o = o.predict(md, xp, diagonal_var=False, pc=pc)
        types = [TYPES[c] for c in cov]
        mh = _ffi.ModelHandle(ctx, types, D, x, y)
        mh.update_cache(hp)
        mu, var, _ = mh.predict(xp, want_var=True)
        mu2, var2, cv = mh.predict(xp, want_var=True, want_cov=True)
        mu_o2 = mu_o.reshape(M, -1)
        e1 = relerr(mu, mu_o2)
        e2 = float(np.abs(var - var_o).max())
        e3 = float(np.abs(cv - cov_o).max())
        e4 = relerr(mu2, mu_o2)
        # same_x path
        mus, _, _ = mh.predict(x, same_x=True)
        mus_o = o.predict_mean(md, x, pc=pc, same=True).reshape(N, -1)
        e5 = relerr(mus, mus_o)
        print(f"predict {cov} N={N} M={M} ny={ny}: mean {e1:.1e} var(abs) {e2:.1e} cov(abs) {e3:.1e} mean(cov path) {e4:.1e} same_x mean {e5:.1e}",
              "OK" if max(e1, e4, e5) < 1e-8 and max(e2, e3) < 1e-8 else "BAD")
        mh.close()


def sec_split():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    ctx = _ffi.get_context()
    ctx.set_option("predict_tile", 256)
    rng = np.random.default_rng(8)
    for cov, N, D, ne, nq in (((o.SE, o.NOISE), 200, 2, 20, 30), ((o.SE, o.SE, o.NOISE), 300, 5, 50, 10), ((o.SE,), 100, 3, 10, 130)):
        x, y = _model(cov, N, D, 31)
        xe = 0.5 * rng.random((D, ne))
        xq = 0.5 * rng.random((D, nq))
        hp = 0.3 + rng.random(o.dim_hp(cov, D))
        covo = cov if len(cov) > 1 else cov[0]
        md = o.GPRModel(covo, hp, x, y)
        cm = o.Cmap(xe, xq)
        A_o, B_o, C_o = o.split_kernel(covo, hp, cm, x)
        types = [TYPES[c] for c in cov]
        A, B, C = _ffi.split_kernel_arrays(ctx, types, D, hp, xe, xq, x)
        eA, eB, eC = relerr(A, A_o), relerr(B, B_o), relerr(C, C_o)
        mu_o, var_o = o.split_predict(md, cm, var_range=(1, 3))
        mh = _ffi.ModelHandle(ctx, types, D, x, y)
        mh.update_cache(hp)
        mu, var = mh.split_predict(xe, xq, var_range=(1, 3))
        e1 = relerr(mu, mu_o)
        e2 = float(np.abs(var - var_o).max())
        mu_f, var_f = mh.split_predict(xe, xq, var_range=(1, ne))
        mu_of, var_of = o.split_predict(md, cm, var_range=(1, ne))
        e3 = float(np.abs(var_f - var_of).max())
        print(f"split {cov} N={N} ne={ne} nq={nq}: A {eA:.1e} B {eB:.1e} C {eC:.1e} mean {e1:.1e} var(1:3) {e2:.1e} var(all) {e3:.1e}",
              "OK" if max(eA, eB, eC, e1) < 1e-9 and max(e2, e3) < 1e-8 else "BAD")
        mh.close()


def sec_perf():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    ctx = _ffi.get_context()
    for N in (8192, 32768):
        D = 8
        cov = (o.SE, o.SE, o.NOISE)
        rng = np.random.default_rng(3003)
        x = rng.random((D, N))
        y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
        hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
        mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
        for it in range(3):
            hp_it = hp * (1 + 0.01 * it)
            c0 = ctx.launch_count()
            t0 = time.time()
            F, G = mh.nlml_grad(hp_it)
            dt = time.time() - t0
            tm = mh.timings()
            print(f"perf N={N} it={it}: wall {dt * 1e3:.1f} ms F={F:.6f} |G|={np.linalg.norm(G):.4e} launches {ctx.launch_count() - c0}")
            print("    ", {k: round(v, 2) for k, v in tm.items() if v > 0})
            print(f"     potrf {N ** 3 / 3 / tm['potrf'] / 1e9:.2f} TF/s  inverse {2 * N ** 3 / 3 / (tm['trtri'] + tm['lauum']) / 1e9:.2f} TF/s  eval {N ** 3 / (dt * 1e3) / 1e9:.2f} TF/s")
        M = 16384
        xp = rng.random((D, M))
        mh.update_cache(hp)
        t0 = time.time()
        mu, var, _ = mh.predict(xp, want_var=True)
        dt = time.time() - t0
        tm = mh.timings()
        print(f"perf predict N={N} M={M}: wall {dt * 1e3:.1f} ms  {M / dt:.0f} pts/s  {M * N * N / dt / 1e12:.2f} TF/s")
        print("    ", {k: round(v, 2) for k, v in tm.items() if v > 0 and k.startswith('pred')})
        mh.close()


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, n))
    return X @ X.T / n + np.eye(n)


def sec_mgpu_factor():
    """block-cyclic drivers with virtual ranks on the visible device(s): device list cycles over what exists"""
    import scipy.linalg as sl
    from gpr_sm100a import _ffi
    import torch
    ndev = torch.cuda.device_count()
    for (n, nb, G) in ((512, 128, 1), (512, 128, 2), (768, 128, 3), (1024, 256, 2), (1000, 256, 3), (1536, 512, 2), (2048, 256, 4), (2048, 512, 8)):
        K = _spd(n, n + G)
        rng = np.random.default_rng(n)
        Y0 = rng.standard_normal((n, 3))
        U = sl.cholesky(K, lower=False)
        for mode in (0, 1, 2):
            mc = _ffi.MultiContext([r % ndev for r in range(G)], nb=nb)
            A, Y, ms = mc.dbg_factor(np.triu(K), Y0, mode)
            ref = [U, np.linalg.inv(U), np.linalg.inv(K)][mode]
            yref = sl.solve_triangular(U, Y0, trans="T") if mode == 0 else -np.linalg.solve(K, Y0)
            e = relerr(np.triu(A), np.triu(ref))
            ey = relerr(Y, yref)
            low = float(np.abs(np.tril(A, -1)).max())
            ok = e < 1e-11 and ey < 1e-11 and low == 0
            print(f"mgpu factor n={n} nb={nb} G={G} mode={mode}: relerr {e:.2e} rhs {ey:.2e} lower {low:.1e} launches {mc.launch_count()} ms {np.round(ms, 2).tolist()}", "OK" if ok else "BAD")
            mc.close()


def sec_mgpu_nlml():
    from gpr_sm100a import _ffi
    import gpr_oracle as o
    import torch
    ndev = torch.cuda.device_count()
    ctx = _ffi.get_context()
    for cov, N, D, ny, nb, G in (((o.SE, o.NOISE), 300, 5, 1, 128, 2), ((o.SE, o.SE, o.NOISE), 1000, 8, 1, 256, 3),
                                 ((o.SE, o.NOISE), 257, 3, 4, 128, 4), ((o.SE, o.MATERN52, o.NOISE), 700, 4, 1, 128, 2),
                                 ((o.SE, o.SE, o.NOISE), 2048, 8, 1, 512, 4), ((o.SE, o.NOISE), 2000, 16, 1, 256, 1)):
        x, y = _model(cov, N, D, 11, ny)
        rng = np.random.default_rng(5)
        hp = 0.3 + rng.random(o.dim_hp(cov, D))
        hp[-1] = 0.1
        ta = 2 if ny > 1 else 1
        md = o.GPRModel(cov, hp, x, y, train_axis=ta)
        tc = o.MllGradCache(md)
        Fo, Go = o.loss_grad(hp, md, tc)
        types = [TYPES[c] for c in cov]
        mc = _ffi.MultiContext([r % ndev for r in range(G)], nb=nb)
        mm = _ffi.MultiModelHandle(mc, types, D, x, y, train_axis=ta)
        F, Gd = mm.nlml_grad(hp)
        eF = abs(F - Fo) / abs(Fo)
        eGi = float((np.abs(Gd - Go) / np.maximum(np.abs(Go), 1e-8 * np.linalg.norm(Go))).max())
        ea = relerr(mm.fetch(_ffi.FETCH_ALPHA), tc.alpha)
        eK = relerr(mm.fetch(_ffi.FETCH_KINV), tc.Kinv)
        Fl, Gl = mm.nlml_grad(np.log(hp), log_scale=True)
        eL = float(np.abs(Gl - Go * hp).max() / np.abs(Go * hp).max())
        # against the single-GPU path of the same library
        mh = _ffi.ModelHandle(ctx, types, D, x, y, train_axis=ta)
        F1, G1 = mh.nlml_grad(hp)
        e1 = max(abs(F - F1) / abs(F1), float(np.abs(Gd - G1).max() / np.abs(G1).max()))
        mh.close()
        ok = eF < 1e-8 and eGi < 1e-8 and ea < 1e-8 and eK < 1e-8 and eL < 1e-8
        print(f"mgpu nlml {cov} N={N} D={D} ny={ny} nb={nb} G={G}: relF {eF:.1e} relG(comp) {eGi:.1e} alpha {ea:.1e} Kinv {eK:.1e} logG {eL:.1e} vs1gpu {e1:.1e}", "OK" if ok else "BAD")
        mm.close(); mc.close()


def sec_mgpu_perf():
    """timing of the block-cyclic path (ranks cycle over the visible devices)"""
    from gpr_sm100a import _ffi
    import torch
    ndev = torch.cuda.device_count()
    N, D = int(os.environ.get("MGPU_N", "16384")), 8
    rng = np.random.default_rng(3003)
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.1]])
    ctx = _ffi.get_context()
    mh = _ffi.ModelHandle(ctx, [1, 2], D, x, y)
    mh.nlml_grad(hp)
    F1, G1 = mh.nlml_grad(hp * 1.01)
    t1 = mh.timings()
    print(f"1gpu recursive N={N}:", {k: round(v, 1) for k, v in t1.items() if v > 0})
    mh.close()
    for G in sorted(set([1, ndev, 2 * ndev])):
        for nb in (512, 1024):
            mc = _ffi.MultiContext([r % ndev for r in range(G)], nb=nb)
            mm = _ffi.MultiModelHandle(mc, [1, 2], D, x, y)
            mm.nlml_grad(hp)
            t0 = time.perf_counter()
            F, Gd = mm.nlml_grad(hp * 1.01)
            dt = time.perf_counter() - t0
            tm = mm.timings()
            e = max(abs(F - F1) / abs(F1), float(np.abs(Gd - G1).max() / np.abs(G1).max()))
            dense = tm["potrf"] + tm["trtri"] + tm["lauum"]
            print(f"mgpu N={N} G={G} (devices {ndev}) nb={nb}: wall {dt*1e3:.1f} ms, dense {dense:.1f} ms = {N**3 / dense / 1e9:.2f} TFLOP/s aggregate, vs 1gpu {e:.1e}",
                  {k: round(v, 1) for k, v in tm.items() if v > 0})
            mm.close(); mc.close()


SECTIONS = {"gemm_cfgs": sec_gemm_cfgs, "gemm": sec_gemm, "factor": sec_factor, "kernel": sec_kernel, "nlml": sec_nlml, "predict": sec_predict,
            "split": sec_split, "gemm_perf": sec_gemm_perf, "factor_perf": sec_factor_perf, "perf": sec_perf,
            "mgpu_factor": sec_mgpu_factor, "mgpu_nlml": sec_mgpu_nlml, "mgpu_perf": sec_mgpu_perf}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        try:
            SECTIONS[sys.argv[2]]()
        except Exception:
            traceback.print_exc()
            sys.exit(1)
        sys.exit(0)
    names = sys.argv[1:] or list(SECTIONS)
    for nme in names:
        print(f"===== {nme} =====", flush=True)
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", nme], timeout=900)
        print(f"===== {nme}: exit {r.returncode} ({time.time() - t0:.1f} s) =====", flush=True)
