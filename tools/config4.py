"""BASELINE.json config 4: posterior mean + variance on the sum-grid xe[:,e] + xq[:,q] via split predict,
N=32768 training points, ne = nq = 4096 (M = 16 777 216 test points), `e` rows sharded over the ranks.
Mean for all M points; variance for the first `nvar` e rows of every rank's block (per-row work is uniform:
nq * N^2 flop), reported as points/s.  Cross-checked against the dense predict path of the library on explicit
test points (the oracle pins both paths at small sizes in tests/).  Run alone or under torchrun."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi, shard  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ne = nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nvar = int(sys.argv[3]) if len(sys.argv) > 3 else 64
D = 8
rng = np.random.default_rng(4004)
x = rng.random((D, N))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
xe, xq = 0.5 * rng.random((D, ne)), 0.5 * rng.random((D, nq))
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
ctx = _ffi.Context(local)
ctx.set_option("predict_tile", 16384)
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
t0 = time.perf_counter()
mh.update_cache(hp)
t_factor = time.perf_counter() - t0
lo, hi = shard.block_range(ne, rank, world)
xe_blk = np.asfortranarray(xe[:, lo:hi])
nv = min(nvar, hi - lo)


def sync():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


sync()
t0 = time.perf_counter()
mean, _ = mh.split_predict(xe_blk, xq, var_range=None, want_var=False)
sync()
t_mean = time.perf_counter() - t0
t0 = time.perf_counter()
mean2, var = mh.split_predict(xe_blk, xq, var_range=(1, nv), want_var=True)
sync()
t_var = time.perf_counter() - t0 - t_mean          # the second call recomputes the mean as well
if dist is not None:
    tt = torch.tensor([t_mean, t_var], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_mean, t_var = float(tt[0]), float(tt[1])
if rank == 0:
    # cross-check a few rows against the dense path on explicit points
    es = [0, nv // 2, nv - 1]
    worst_mu = worst_var = 0.0
    for e in es:
        pts = np.asfortranarray(xe_blk[:, e:e + 1] + xq[:, :512])
        mu_d, var_d, _ = mh.predict(pts, want_var=True)
        worst_mu = max(worst_mu, float(np.abs(mu_d[:, 0] - mean[e, :512]).max()))
        worst_var = max(worst_var, float(np.abs(var_d - var[e * nq:e * nq + 512]).max()))
    Mtot = ne * nq
    print(f"config4 world={world} N={N} ne=nq={ne}: factor {t_factor * 1e3:.0f} ms; split mean for all {Mtot} points: {t_mean * 1e3:.1f} ms "
          f"({Mtot / t_mean / 1e6:.1f} Mpts/s, {2 * ne * nq * N * 2 / t_mean / 1e12:.1f} TFLOP/s); variance of {world * nv} e-rows "
          f"({world * nv * nq} points): {t_var:.2f} s = {world * nv * nq / t_var:.0f} pts/s ({world * nv * nq * float(N) ** 2 / t_var / 1e12:.1f} TFLOP/s aggregate); "
          f"split vs dense: max |mean diff| {worst_mu:.2e}, max |var diff| {worst_var:.2e}", "OK" if worst_mu < 1e-9 and worst_var < 1e-9 else "BAD")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
