#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP hot path on B200 (contract: task statement (4)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], "3-ref" of SURVEY.md 8d): SquaredExp()+SquaredExp()+WhiteNoise() (P = 19),
N = 32768, D = 8, FP64.  One step = one NLML + gradient evaluation (log_loss_grad!, src/cost.jl:60-70) at the
next point of a fixed hyper-parameter trajectory (so nothing can be cached between steps).
  value : evaluations/s with x, y resident in HBM (only hp goes in and F, G come out per step)
  e2e   : the same through the C ABI with HOST buffers: every step re-uploads x, y from pinned host memory
          (gpr_model_set_x/_y), evaluates, and reads F, G back
  N > 1 : replicas only: every rank evaluates its own hyper-parameter set; no data-path collective
  --impl reference : the reference-shaped CPU path (oracle/gpr_oracle.py: materialised K per component, dpotrf,
          dpotrs on the identity, per-hyper-parameter dK + dgemv + ddot) on the host cores, on a bounded sample
          (smaller N, scaled by N^3 to the metric's configuration; flagged in cpu_baseline.sample)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocessregression.jl_b200"))

METRIC = "NLML+grad evals/s (N=32k,D=8,fp64)"
UNIT = "evals/s"
N_FULL, D_FULL = 32768, 8
WORKLOAD = "3-ref: SquaredExp()+SquaredExp()+WhiteNoise() P=19, N=32768, D=8, one log_loss_grad! per step"


def make_problem(N, D, seed=3003):
    rng = np.random.default_rng(seed)
    x = rng.random((D, N))
    y = np.sin(3 * x).sum(0) + 0.1 * rng.standard_normal(N)
    hp = np.concatenate([[1.0], 0.5 * np.ones(D), [0.5], 2.0 * np.ones(D), [0.1]])
    return x, y, hp


def hp_at(hp0, step, rank=0):
    """Fixed, pre-determined hyper-parameter trajectory (identical work for CPU and GPU arms)."""
    return hp0 * (1.0 + 0.004 * step + 0.01 * rank)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle; checker/baseline only)
def cpu_eval_seconds(N, D, hp0, steps, warmup, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpr_oracle as o
    x, y, _ = make_problem(N, D)
    md = o.GPRModel((o.SE, o.SE, o.NOISE), hp0, x, y)
    tc = o.MllGradCache(md)
    ts = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        o.log_loss_grad(np.log(hp_at(hp0, s)), md, tc)
        dt = time.perf_counter() - t0
        if s >= warmup:
            ts.append(dt)
    return ts


def cpu_baseline(hp0, budget_s=25.0, steps=1, warmup=0):
    """Reference-shaped CPU path timed on the host cores on a bounded sample and extrapolated to N = 32768 with the
    cost model t(N) = a N^3 + b N^2 (factor / inverse vs the P per-hyper-parameter sweeps), a and b fitted to the
    measured times at N/2 and N of the sample (plain N^3 scaling would overstate the CPU time: at sample sizes the
    N^2 P term is a large share)."""
    cores = os.cpu_count() or 1
    t_probe = cpu_eval_seconds(2048, D_FULL, hp0, 1, 1, cores)[0]
    Ns = 4096
    for cand in (8192,):
        if t_probe * (cand / 2048) ** 3 * (steps + warmup) <= budget_s:
            Ns = cand
    ts = cpu_eval_seconds(Ns, D_FULL, hp0, steps, warmup, cores)
    t = float(np.mean(ts))
    t_half = t_probe if Ns == 4096 else float(np.mean(cpu_eval_seconds(Ns // 2, D_FULL, hp0, 1, 0, cores)))
    n = Ns / 2.0
    a = (t - 4.0 * t_half) / (4.0 * n ** 3)
    b = (t_half - a * n ** 3) / n ** 2
    if a <= 0 or b < 0:                      # noisy fit: fall back to pure N^3 scaling
        a, b = t / Ns ** 3, 0.0
    t_full = a * N_FULL ** 3 + b * N_FULL ** 2
    scale = t_full / t
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle (numpy+scipy OpenBLAS, {cores} threads) log_loss_grad at N={Ns}, D=8, P=19: {t:.2f} s/eval measured "
                      f"({t_half:.2f} s at N={Ns // 2}); extrapolated to N=32768 with t = a N^3 + b N^2 fitted to the two sizes "
                      f"(x{scale:.0f}) -- extrapolated",
            "seconds_per_eval_sample": t, "sample_N": Ns, "scale": scale}, ts, Ns


def run_reference(args, rank, world):
    if rank != 0:
        return
    _, _, hp0 = make_problem(16, D_FULL)
    budget = 150.0
    base, ts, Ns = cpu_baseline(hp0, budget_s=budget, steps=args.steps, warmup=args.warmup)
    t = float(np.mean(ts))
    scale = base["scale"]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * scale * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "timing": f"host wall clock; each step is a bounded sample at N={Ns}, extrapolated with t = a N^3 + b N^2"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def measure_fp64_peak(torch, n=8192):
    """cuBLAS DGEMM through torch.matmul: the FP64 denominator (MEASURED_PEAKS.json has no FP64 figure)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / 40
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best / 1e9, 2 * n ** 3 / sustained / 1e9


def run_gpu(args, rank, world, local_rank):
    import torch
    import gpr_sm100a as g
    from gpr_sm100a import _ffi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    N, D = args.n, D_FULL
    x, y, hp0 = make_problem(N, D)
    P = len(hp0)
    ctx = g.Context(local_rank)
    cov = g.SquaredExp() + g.SquaredExp() + g.WhiteNoise()
    # host buffers of the e2e arm live in pinned memory
    xh = torch.empty((N, D), dtype=torch.float64).pin_memory()
    yh = torch.empty((N,), dtype=torch.float64).pin_memory()
    xh.numpy()[...] = x.T                     # (N, D) C-order == (D, N) column major
    yh.numpy()[...] = y
    x_f = xh.numpy().T                         # F-contiguous (D, N) view of the pinned buffer
    y_f = yh.numpy()
    mh = _ffi.ModelHandle(ctx, [1, 1, 2], D, x_f, y_f)

    fp64_burst = fp64_sust = None
    if rank == 0:
        fp64_burst, fp64_sust = measure_fp64_peak(torch)

    def step(s):
        return mh.nlml_grad(np.log(hp_at(hp0, s, rank)), log_scale=True)

    for s in range(args.warmup):
        step(s)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    stage = {}
    t0 = time.perf_counter()
    for s in range(args.steps):
        F, G = step(args.warmup + s)
        for k, v in mh.timings().items():     # CUDA events recorded by the library on ITS stream (torch.cuda.Event cannot see it)
            stage[k] = stage.get(k, 0.0) + v
    barrier()
    t_wall = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # device time of the K timed steps: event pair around every evaluation on the library stream, summed; max over ranks
    t_loc = stage["eval"] * 1e-3
    t_max = t_loc
    if dist is not None:
        tt = torch.tensor([t_loc, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max, t_wall = float(tt[0].item()), float(tt[1].item())
    value = world * args.steps / t_max

    # end-to-end through the C ABI with host buffers
    def e2e_step(s):
        mh.set_x(x_f)
        mh.set_y(y_f)
        return mh.nlml_grad(np.log(hp_at(hp0, s, rank)), log_scale=True)

    e2e_step(1000)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        F2, G2 = e2e_step(2000 + s)
    barrier()
    t_e = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([t_e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e = float(tt.item())
    e2e = {"value": world * args.steps / t_e, "unit": UNIT, "h2d_bytes_per_step": 8 * (D * N + N + P), "d2h_bytes_per_step": 8 * (P + 1)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    K = args.steps
    ms = {k: v / K for k, v in stage.items()}
    flops_factor_inverse = float(N) ** 3                                # N^3/3 potrf + 2N^3/3 trtri+lauum
    t_dense = (ms["potrf"] + ms["trtri"] + ms["lauum"]) * 1e-3
    achieved = flops_factor_inverse / t_dense / 1e12
    # the GEMM kernel alone, one launch (8192^3), timed with CUDA events inside the library
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((4096, 4096)))
    _, gemm_ms = _ffi.dbg_dgemm(ctx, "T", "N", 1.0, A, A, 0.0, np.zeros((4096, 4096), order="F"), reps=6)
    gemm_tf = 2 * 4096 ** 3 / gemm_ms / 1e9

    # secondary metrics of BASELINE.json: Cholesky TFLOP/s, predict points/s
    extra = {"cholesky_tflops": N ** 3 / 3 / ms["potrf"] / 1e9, "inverse_tflops": 2 * N ** 3 / 3 / (ms["trtri"] + ms["lauum"]) / 1e9,
             "stage_ms_per_step": {k: round(v, 3) for k, v in ms.items() if v > 0 and not k.startswith("pred")}}
    if not args.no_predict:
        M = 16384
        xp = np.asfortranarray(np.random.default_rng(4004).random((D, M)))
        mh.update_cache(hp0)
        mh.predict(xp, want_var=True)
        t0 = time.perf_counter()
        mu, var, _ = mh.predict(xp, want_var=True)
        dt = time.perf_counter() - t0
        tp = mh.timings()
        extra["predict_points_per_s"] = M / dt
        extra["predict_var_tflops"] = M * float(N) ** 2 / dt / 1e12
        extra["predict_stage_ms"] = {k: round(v, 3) for k, v in tp.items() if k.startswith("pred")}
        # the two streaming kernels (K* . wt and row norms of V) read 8*M*N bytes each
        hbm = 6478.9
        try:
            hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        # streams of the prediction: the K* tile is written once (8*M*N bytes) and V = K* U^-1 is read once for the row
        # norms; mu = K* wt is reduced inside the K* build (no separate stream since round 1c)
        extra["predict_stream_gbs"] = {"kstar_build_write": 8.0 * M * N / tp["pred_kstar"] / 1e6, "rownorm_v": 8.0 * M * N / tp["pred_rownorm"] / 1e6,
                                       "mean": "fused into the K* build", "hbm_peak_gbs": hbm}
        mh.predict(xp, want_var=False)
        t0 = time.perf_counter()
        mh.predict(xp, want_var=False)
        extra["predict_mean_only_points_per_s"] = M / (time.perf_counter() - t0)
        extra["predict_config"] = f"mean+diag variance, M={M} general test points, host in/out, N={N}"

    hp_peak = None
    try:
        hp_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM traffic of the dominant kernel per step: from the committed ncu capture of this command (profiles/), GB
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        if N == tj.get("N"):
            traffic = tj["dgemm128_dram_gb_per_eval"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": fp64_sust, "unit": "TFLOP/s", "frac": achieved / fp64_sust,
                "traffic": traffic, "traffic_unit": "GB of DRAM read+write by all dgemm128 launches of one step (ncu, profiles/dram_traffic.json)",
                "kernel": "dgemm128_kernel (DMMA) inside blocked potrf+trtri+lauum: N^3 flop per step / (potrf+trtri+lauum) CUDA-event ms",
                "peak_source": f"cuBLAS DGEMM 8192^3 via torch.matmul, sustained {fp64_sust:.1f} / burst {fp64_burst:.1f} TFLOP/s measured in this run "
                               "(MEASURED_PEAKS.json holds no FP64 figure" + (f"; its HBM copy figure is {hp_peak.get('hbm_gbs')} GB/s)" if hp_peak else ")"),
                "kernel_isolated_tflops": gemm_tf, "kernel_isolated_frac": gemm_tf / fp64_burst}

    base = None
    if world == 1 and not args.no_cpu:
        base, _, _ = cpu_baseline(hp0, budget_s=25.0)
        base = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if N == N_FULL else WORKLOAD.replace("N=32768", f"N={N}"),
                       "parallelism": "replicas only (one hyper-parameter set per GPU, no data-path collective)" if world > 1 else "1 GPU",
                       "l2": "inputs larger than L2 (K is 8.6 GB per evaluation); no explicit flush needed",
                       "timing": "CUDA events on the library stream around every evaluation (hp upload .. F, G read back), summed over the K "
                                 "steps, max over ranks; barrier + torch.cuda.synchronize on both sides; wall clock reported beside it"},
            "wall_ms_per_step": t_wall / args.steps * 1e3,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": base,
            "extra": extra, "F_last": F, "G_norm_last": float(np.linalg.norm(G))}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="training-set size (default: the metric's N=32768)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-predict", action="store_true", help="skip the secondary predict metric")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
