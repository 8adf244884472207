"""CPU check of the blocked-algorithm drivers (csrc/blocked.hpp) with the plain-loop test backend
(tests/hostlogic/hostlogic.cpp).  The CUDA library instantiates the very same template."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hl(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hl") / "libhostlogic_test.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "hostlogic", "hostlogic.cpp")])
    L = ctypes.CDLL(out)
    for f in (L.hl_factor, L.hl_potrs, L.hl_trsm_run, L.hl_potrsv, L.hl_dist_factor):
        f.restype = ctypes.c_longlong
    return L


def dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def spd(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, n))
    return X @ X.T / n + np.eye(n)


@pytest.mark.parametrize("n", [128, 256, 384, 640, 1024])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 16, 20, 21])
def test_blocked_factor(hl, n, mode):
    K = spd(n, n)
    A = np.asfortranarray(K.copy())
    calls = ctypes.c_longlong(0)
    info = hl.hl_factor(dp(A), ctypes.c_int64(n), mode, ctypes.byref(calls))
    mode &= 15                                   # bit 4: leaf look-ahead ordering of the update
    assert info == 0
    U = sl.cholesky(K, lower=False)
    ref = [U, np.linalg.inv(U), np.linalg.inv(K), np.linalg.inv(K), np.linalg.inv(K), np.linalg.inv(K)][mode]   # modes 3-5: out-of-place inverse
    np.testing.assert_allclose(np.triu(A), np.triu(ref), rtol=0, atol=1e-12 * np.abs(ref).max())
    # dpotrf('U') semantics: the strict lower triangle keeps K (test/test_loss.jl:46)
    assert np.array_equal(np.tril(A, -1), np.tril(K, -1))


def test_blocked_not_posdef(hl):
    K = np.eye(384)
    K[300, 300] = -2.0
    A = np.asfortranarray(K)
    assert hl.hl_factor(dp(A), ctypes.c_int64(384), 0, None) == 301     # LAPACK info: 1-based failing pivot


@pytest.mark.parametrize("n", [128, 512])
def test_blocked_solves(hl, n):
    K = spd(n, 7 * n)
    rng = np.random.default_rng(n)
    B = np.asfortranarray(rng.standard_normal((n, 128)))
    B0 = B.copy()
    A = np.asfortranarray(K.copy())
    assert hl.hl_potrs(dp(A), ctypes.c_int64(n), dp(B), ctypes.c_int64(128)) == 0
    np.testing.assert_allclose(B, np.linalg.solve(K, B0), atol=1e-12)
    v = rng.standard_normal(n)
    v0 = v.copy()
    A = np.asfortranarray(K.copy())
    assert hl.hl_potrsv(dp(A), ctypes.c_int64(n), dp(v)) == 0
    np.testing.assert_allclose(v, np.linalg.solve(K, v0), atol=1e-12)
    R = np.asfortranarray(rng.standard_normal((256, n)))
    R0 = R.copy()
    A = np.asfortranarray(K.copy())
    assert hl.hl_trsm_run(dp(A), ctypes.c_int64(n), dp(R), ctypes.c_int64(256)) == 0
    np.testing.assert_allclose(R, R0 @ np.linalg.inv(sl.cholesky(K, lower=False)), atol=1e-12)


# ---- block-cyclic multi-rank drivers (csrc/dist_blocked.hpp), G simulated ranks in one process
@pytest.mark.parametrize("n,nb,G", [(512, 128, 1), (512, 128, 2), (768, 128, 3), (1024, 256, 2), (1024, 256, 3),
                                    (1024, 128, 4), (768, 384, 2), (1280, 256, 4), (512, 256, 4)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_dist_factor(hl, n, nb, G, mode):
    K = spd(n, n + G)
    rng = np.random.default_rng(n + nb + G)
    Y0 = rng.standard_normal((n, 128))
    A = np.asfortranarray(np.triu(K))
    Y = np.asfortranarray(Y0.copy())
    info = hl.hl_dist_factor(dp(A), ctypes.c_int64(n), ctypes.c_int64(nb), G, dp(Y), ctypes.c_int64(128), mode)
    assert info == 0
    U = sl.cholesky(K, lower=False)
    ref = [U, np.linalg.inv(U), np.linalg.inv(K)][mode]
    np.testing.assert_allclose(np.triu(A), np.triu(ref), rtol=0, atol=1e-11 * np.abs(ref).max())
    assert not np.tril(A, -1).any()
    yref = sl.solve_triangular(U, Y0, trans="T") if mode == 0 else -np.linalg.solve(K, Y0)
    np.testing.assert_allclose(Y, yref, rtol=0, atol=1e-11 * np.abs(yref).max())


def test_dist_not_posdef(hl):
    K = np.eye(512)
    K[300, 300] = -2.0
    A = np.asfortranarray(K)
    Y = np.zeros((512, 128), order="F")
    assert hl.hl_dist_factor(dp(A), ctypes.c_int64(512), ctypes.c_int64(128), 2, dp(Y), ctypes.c_int64(128), 0) == 301


def test_unshipped_epilogue_patch_still_applies():
    """profiles/ozaki_batched_epilogue_r2ar.patch (the batched epilogue of the INT8 kernels: measured, not shipped -- DESIGN.md section 9)
    must keep applying to the kernel sources, or the record stops being something a maintainer can act on."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    patch = os.path.join(root, "profiles", "ozaki_batched_epilogue_r2ar.patch")
    assert os.path.exists(patch)
    if shutil.which("git") is None:
        pytest.skip("git not available")
    r = subprocess.run(["git", "apply", "--check", patch], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
