"""Host-side logic of the mirrored API (no device calls): kernel algebra, hp bookkeeping, Cmap, sharding."""
import numpy as np
import pytest


def test_compose_and_dim_hp(gpr):
    k = gpr.SquaredExp() + gpr.WhiteNoise()
    assert isinstance(k, gpr.ComposedKernel) and len(k.kernels) == 2
    k3 = k + gpr.SquaredExp()
    assert [type(c).__name__ for c in k3.kernels] == ["SquaredExp", "WhiteNoise", "SquaredExp"]
    k4 = k + k
    assert len(k4.kernels) == 4
    assert gpr.dim_hp(gpr.SquaredExp(), 5) == 6 and gpr.dim_hp(gpr.WhiteNoise(), 5) == 1
    assert gpr.dim_hp(k3, 2) == 7
    assert gpr.SquaredExp() == gpr.SquaredExp() and gpr.SquaredExp() != gpr.WhiteNoise()


def test_split_find_idx_rm_noise(gpr):
    hp = np.arange(1.0, 8.0)
    parts = gpr.split(hp, [3, 1, 3])
    assert [list(p) for p in parts] == [[1, 2, 3], [4], [5, 6, 7]]
    assert gpr.find_idx([3, 1, 3], 1) == (1, 1)
    assert gpr.find_idx([3, 1, 3], 3) == (1, 3)
    assert gpr.find_idx([3, 1, 3], 4) == (2, 1)
    assert gpr.find_idx([3, 1, 3], 7) == (3, 3)
    ks, hs = gpr.rm_noise(gpr.SquaredExp() + gpr.WhiteNoise() + gpr.SquaredExp(), parts)
    assert len(ks) == 2 and list(hs[1]) == [5, 6, 7]


def test_model_checks_mirror_reference_errors(gpr):
    x, y = np.random.rand(2, 10), np.random.rand(10)
    with pytest.raises(gpr.GPRError, match="Parameter size mismatch"):
        gpr.GPRModel(gpr.SquaredExp(), np.ones(2), x, y)
    with pytest.raises(gpr.GPRError, match="x and y size mismatch"):
        gpr.GPRModel(gpr.SquaredExp(), np.ones(3), x, np.random.rand(9))
    md = gpr.GPRModel(gpr.SquaredExp() + gpr.WhiteNoise(), x, np.random.rand(10, 4), train_axis=3)
    assert md.params.shape == (4,) and gpr.get_sample(md).shape == (10,)
    assert np.array_equal(gpr.get_sample(md), md.y[:, 2])


def test_islog_and_noise_pieces(gpr):
    x, y = np.random.rand(2, 10), np.random.rand(10)
    ll = gpr.MarginalLikelihood()
    assert isinstance(gpr.islog(ll, gpr.GPRModel(gpr.SquaredExp(), x, y)), gpr.LogScale)
    assert isinstance(gpr.islog(ll, gpr.GPRModel(gpr.SquaredExp() + gpr.WhiteNoise(), x, y)), gpr.LogScale)
    u = gpr.kernel(gpr.WhiteNoise(), np.array([0.3]), x)
    assert isinstance(u, gpr.UniformScaling) and u.lam == pytest.approx(0.09)
    assert gpr.kernel(gpr.WhiteNoise(), np.array([0.3]), x, np.random.rand(2, 4)) == 0.0
    K = np.zeros((10, 10))
    gpr.add_noise_(K, gpr.SquaredExp() + gpr.WhiteNoise(), np.array([1.0, 1.0, 1.0, 0.5]), x)
    assert np.allclose(np.diag(K), 0.25)
    assert np.all(gpr.init_params(ll, gpr.GPRModel(gpr.SquaredExp(), x, y)) == 1.0)


def test_cmap_indexing(gpr):
    """test/test_split_kernel.jl:10-22"""
    rng = np.random.default_rng(0)
    xe, xq = rng.random((3, 5)), rng.random((3, 4))
    cm = gpr.Cmap(np.add, xe, xq)
    assert cm.shape == (3, 5, 4)
    np.testing.assert_allclose(cm[2, 3].ravel(), xe[:, 2] + xq[:, 3])
    flat = cm[:, :]
    ref = np.stack([xe[:, i] + xq[:, j] for j in range(4) for i in range(5)], axis=1)
    np.testing.assert_allclose(flat, ref)
    np.testing.assert_allclose(cm[:, 1], np.stack([xe[:, i] + xq[:, 1] for i in range(5)], axis=1))


def test_block_ranges_cover(gpr):
    from gpr_sm100a.shard import block_range, replica_indices, sharded_split_rows
    for total in (0, 1, 7, 16, 16777216):
        for world in (1, 2, 3, 8):
            r = [block_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    assert sorted(sum((replica_indices(10, k, 4) for k in range(4)), [])) == list(range(10))
    assert sharded_split_rows(4096, 0, 8) == (1, 512) and sharded_split_rows(4096, 7, 8) == (3585, 4096)


def test_inplace_result_arrays(gpr):
    """predict!(mu, Sigma, ...) writes in place (src/predict.jl:36-71): the mirror hands the caller's array to the C ABI only when
    the layout allows it and falls back to a copy otherwise."""
    from gpr_sm100a import api
    mu = np.zeros((4, 6), order="F")
    assert api._inplace(mu, (4, 6), "F") is mu
    assert api._inplace(mu, (6, 4), "F") is None                               # wrong shape
    assert api._inplace(np.zeros((4, 6)), (4, 6), "F") is None                 # C order
    assert api._inplace(np.zeros((4, 6), dtype=np.float32, order="F"), (4, 6), "F") is None
    assert api._inplace(mu[:, ::2], (4, 3), "F") is None                       # strided view
    v = np.zeros(24)
    assert api._inplace(v, (24,), "C") is v
    ro = np.zeros(24)
    ro.setflags(write=False)
    assert api._inplace(ro, (24,), "C") is None
    assert api._inplace([0.0] * 24, (24,), "C") is None


def test_bench_stage_wise_scaling():
    """bench.py's CPU baseline: one measured evaluation is scaled to N = 32768 stage by stage -- the factorization and the dpotrs on
    the identity by (32768/N)^3, everything else by (32768/N)^2 (reference stages: src/cost.jl:96-127)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    st = {"kbuild": 4.0, "sum": 0.7, "potrf": 1.2, "potrs_identity": 6.6, "gradient": 10.3}
    dt = 23.0                                                                  # 0.2 s outside the named stages: counted as N^2
    t = bench.scale_by_stage(dt, st, 16384)
    assert abs(t - ((1.2 + 6.6) * 8 + (dt - 7.8) * 4)) < 1e-12
    assert abs(bench.scale_by_stage(dt, st, 32768) - dt) < 1e-12               # already at full size
    assert bench.scale_by_stage(5.0, {"potrf": 4.0, "potrs_identity": 4.0}, 16384) == 5.0 * 8      # stages cannot exceed the total
    a, b = bench.fit_cost_model([2048, 4096, 16384], [0.33, 1.31, 22.9])
    assert a >= 0 and b >= 0
