"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

SURVEY.md 8e: the path shards without any data-path collective where its units are independent --
  * prediction: contiguous blocks of test points (split predict: blocks of `e` rows) per rank; x, y, hp are
    replicated and every rank factors K itself (N^3/3 once, no exchange);
  * NLML + gradient at N <= 32768: "replicas only" -- different hyper-parameter sets (line-search points,
    restarts, the 2P finite-difference points of update_sample!, src/update_model.jl:87-94) per rank.
Results are combined with one all_gather of the outputs (gloo on CPU, nccl on GPU); nothing is exchanged
inside the timed data path.
"""
import numpy as np


def block_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replica_indices(n_items, rank, world):
    """Round-robin assignment of independent evaluations (hyper-parameter sets) to ranks."""
    return list(range(rank, int(n_items), int(world)))


def _dist():
    import torch.distributed as dist
    return dist


def sharded_predict(predict_fn, xp, ny=1, want_var=True, group=None):
    """Every rank calls predict_fn(xp[:, lo:hi]) -> (mean (m, ny), var (m,) or None) on its block of test
    points and the blocks are all-gathered.  predict_fn is the device path in production
    (`lambda blk: handle.predict(blk, want_var=True)[:2]`); the CPU test-suite injects a stand-in."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    M = xp.shape[1]
    lo, hi = block_range(M, rank, world)
    mean, var = predict_fn(xp[:, lo:hi]) if hi > lo else (np.zeros((0, ny)), np.zeros(0))
    mean = np.asarray(mean).reshape(hi - lo, ny)
    if world == 1:
        return mean, (np.asarray(var) if want_var else None)
    sizes = [block_range(M, r, world) for r in range(world)]
    maxlen = max(h - l for l, h in sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros((maxlen, ny + 1), dtype=torch.float64, device=dev)
    buf[:hi - lo, :ny] = torch.from_numpy(np.ascontiguousarray(mean)).to(dev)
    if want_var:
        buf[:hi - lo, ny] = torch.from_numpy(np.ascontiguousarray(var)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = [o[:h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)]
    full = np.concatenate(parts, axis=0)
    return full[:, :ny], (full[:, ny] if want_var else None)


def sharded_split_rows(ne, rank, world):
    """Block of `e` rows (1-based inclusive var_range) of a split prediction owned by `rank`."""
    lo, hi = block_range(ne, rank, world)
    return lo + 1, hi


def replicated_nlml_grad(eval_fn, hp_sets, group=None):
    """Each rank evaluates its round-robin share of independent hyper-parameter sets with eval_fn(hp) ->
    (F, G) and the results are all-gathered: returns (F[n], G[n, P]) on every rank."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(hp_sets)
    P = len(hp_sets[0])
    mine = replica_indices(n, rank, world)
    local = np.zeros((n, P + 1))
    for i in mine:
        F, G = eval_fn(np.asarray(hp_sets[i], dtype=np.float64))
        local[i, 0] = F
        local[i, 1:] = G
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(local).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)   # disjoint supports: a gather expressed as a sum
        local = t.cpu().numpy()
    return local[:, 0], local[:, 1:]


def dist_context(device=None, nb=1024, group=None):
    """DistContext for this torch.distributed rank: rank 0 draws the NCCL unique id of the library's own communicator
    and broadcasts it through the already initialised process group (any backend)."""
    import torch
    from . import _ffi
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = [_ffi.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    return _ffi.DistContext(device, rank, world, ids[0], nb=nb)
