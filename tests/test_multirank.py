"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: test-point sharding of prediction and
hyper-parameter replicas of NLML+grad.  The device predictor is replaced by the oracle here
(checker only); on the GPU box bench.py --gpus N exercises the same functions over nccl."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpr_oracle as o
    from gpr_sm100a import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    D, N, M = 3, 80, 37
    x, xp = rng.random((D, N)), rng.random((D, M))
    y = np.sin(3 * x).sum(0)
    cov = (o.SE, o.NOISE)
    hp = np.array([1.0, 0.7, 0.9, 1.1, 0.1])
    md = o.GPRModel(cov, hp, x, y)
    pc = o.GPRPredictCache(md)

    def predict_fn(blk):
        mu, var = o.predict(md, blk, diagonal_var=True, pc=pc, same=False)
        return mu.reshape(-1, 1), var

    mean, var = shard.sharded_predict(predict_fn, xp, ny=1, want_var=True)
    mu_ref, var_ref = o.predict(md, xp, diagonal_var=True, pc=pc, same=False)
    ok1 = np.allclose(mean[:, 0], mu_ref, rtol=0, atol=1e-12) and np.allclose(var, var_ref, rtol=0, atol=1e-12)

    hp_sets = [hp * (1 + 0.05 * k) for k in range(5)]
    F, G = shard.replicated_nlml_grad(lambda h: o.loss_grad(h, md), hp_sets)
    ref = [o.loss_grad(h, md) for h in hp_sets]
    ok2 = all(abs(F[i] - ref[i][0]) < 1e-10 and np.allclose(G[i], ref[i][1], atol=1e-9) for i in range(5))
    q.put((rank, bool(ok1), bool(ok2), shard.block_range(M, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_predict_and_replicas_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == (0, 19) and res[1][3] == (19, 37)
