#!/bin/bash
# round 2, GPU call AC: strong-scaling base of config 5 -- N = 131072 in place on ONE B200 (137 GB matrix), block-cyclic driver with one rank
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
timeout 900 python tools/config5.py --n 131072 --gpus 1 --nb 2048 --evals 1 > gpurun_out/r2ac_base.log 2>&1; tail -2 gpurun_out/r2ac_base.log | cut -c1-900
