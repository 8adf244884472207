#!/bin/bash
# round 2, GPU call AF (4 GPUs): the 4- and 2-GPU points of the strong-scaling curve of config 5 (N = 131072, nb = 2048, one process driving all devices)
mkdir -p gpurun_out
timeout 900 python tools/config5.py --n 131072 --gpus 4,2 --nb 2048 --evals 1 > gpurun_out/r2af_scaling.log 2>&1; tail -2 gpurun_out/r2af_scaling.log | cut -c1-700
