#!/bin/bash
# round 2, GPU call A: parity suite + default bench (with the sharded extras) + host facts
set -x
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
