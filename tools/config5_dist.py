"""Config 5 with ONE PROCESS PER GPU (torchrun) and NCCL as the transport of the block-cyclic factorization:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/config5_dist.py [N [NB [check]]]          (positional: torchrun would try to parse --n... itself)

Every rank evaluates NLML + gradient of the same model; rank 0 prints one JSON line.  --check also runs the
single-process path (gpr_mgpu_*, peer-memory transport, ranks cycled over the visible devices) on rank 0 and compares."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))
from gpr_sm100a import _ffi, shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int, nargs="?", default=32768)
ap.add_argument("nb", type=int, nargs="?", default=1024)
ap.add_argument("check", nargs="?", default="")
args = ap.parse_args()
args.d = 16
args.check = args.check == "check"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, D = args.n, args.d
rng = np.random.default_rng(5005)
x = np.asfortranarray(rng.random((D, N)))
y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
hp = np.concatenate([[1.0], 0.4 * np.ones(D), [0.1]])
dc = shard.dist_context(local, nb=args.nb)
mm = _ffi.MultiModelHandle(dc, [1, 2], D, x, y)
mm.nlml_grad(hp * 0.99)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
F, G = mm.nlml_grad(hp)
dt = time.perf_counter() - t0
Fl, _ = mm.nlml_grad(hp, want_g=False)
tm = mm.timings()
tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
Fs = torch.tensor([F], dtype=torch.float64, device="cuda")
Fall = [torch.zeros_like(Fs) for _ in range(world)]
dist.all_gather(Fall, Fs)
if rank == 0:
    out = {"config": "5-dist (one process per GPU, NCCL)", "N": N, "D": D, "world": world, "nb": args.nb, "s_per_eval": float(tt.item()),
           "F": F, "F_loss_only": Fl, "F_equal_on_all_ranks": bool(all(float(f.item()) == F for f in Fall)), "G_norm": float(np.linalg.norm(G))}
    if args.check:
        nd = _ffi.device_count()
        mc = _ffi.MultiContext([0], nb=args.nb)      # single-rank reference on this process's device
        m1 = _ffi.MultiModelHandle(mc, [1, 2], D, x, y)
        F1, G1 = m1.nlml_grad(hp)
        out["relF_vs_single_process"] = abs(F - F1) / abs(F1)
        out["relG_vs_single_process"] = float((np.abs(G - G1) / np.maximum(np.abs(G1), 1e-8 * np.linalg.norm(G1))).max())
        m1.close(); mc.close()
    out["phase_ms"] = {k: round(v, 1) for k, v in tm.items() if v > 0}
    print(json.dumps(out), flush=True)
dist.barrier()
mm.close(); dc.close()
dist.destroy_process_group()
