#!/bin/bash
# round 2, GPU call M (8 GPUs): bench.py under torchrun, exactly as the driver launches it
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err
tail -c 2500 gpurun_out/r2m_bench_n8.json; tail -8 gpurun_out/r2m_bench_n8.err
