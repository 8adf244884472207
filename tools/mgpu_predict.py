"""Multi-GPU check of the sharded prediction path (run under torchrun, one rank per GPU):
every rank factors its replica of the model, predicts its block of test points (general and split),
the outputs are all-gathered over NCCL and rank 0 compares them with a single-GPU evaluation of all points."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi, shard  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
D = 8
rng = np.random.default_rng(4004)
x = rng.random((D, N))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
xp = rng.random((D, M))
hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
ctx = _ffi.Context(local)
mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x, y)
mh.update_cache(hp)


def fn(blk):
    mu, var, _ = mh.predict(blk, want_var=True)
    return mu, var


fn(xp[:, :256])
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
mean, var = shard.sharded_predict(fn, xp, ny=1, want_var=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    mu1, var1, _ = mh.predict(xp, want_var=True)
    print(f"sharded predict world={world} N={N} M={M}: {M / dt:.0f} pts/s; max |mean diff| {np.abs(mean - mu1).max():.2e}, "
          f"max |var diff| {np.abs(var - var1).max():.2e}", "OK" if np.abs(mean - mu1).max() < 1e-12 and np.abs(var - var1).max() < 1e-12 else "BAD")
# hyper-parameter replicas
sets = [hp * (1 + 0.01 * k) for k in range(2 * world)]
F, G = shard.replicated_nlml_grad(lambda h: mh.nlml_grad(h), sets)
if rank == 0:
    F0, G0 = mh.nlml_grad(sets[-1])
    print(f"replicas world={world}: F[-1] {F[-1]:.10f} vs local {F0:.10f}; max |G diff| {np.abs(G[-1] - G0).max():.2e}",
          "OK" if abs(F[-1] - F0) < 1e-9 * abs(F0) else "BAD")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
