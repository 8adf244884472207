#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -k "mgpu or config5 or dist or ozaki" > gpurun_out/r2ab_pytest.log 2>&1; tail -3 gpurun_out/r2ab_pytest.log
timeout 300 python tools/config5.py --n 32768 --gpus 1 --nb 2048 --evals 1 2>&1 | tail -1 | cut -c1-400
