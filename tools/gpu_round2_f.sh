#!/bin/bash
# round 2, GPU call F (2 GPUs): bench.py under torchrun, exactly as the driver launches it
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -c 4000 gpurun_out/r2f_bench_n2.json; tail -15 gpurun_out/r2f_bench_n2.err
