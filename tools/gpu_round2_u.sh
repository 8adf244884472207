#!/bin/bash
# round 2, GPU call U: A/B of the INT8 route's knobs (L2-resident digit extraction, two-window nine-digit product, smallest routed product)
mkdir -p gpurun_out
timeout 1200 python tools/ozaki_sweep.py > gpurun_out/r2u_sweep.log 2>&1; cat gpurun_out/r2u_sweep.log | tail -30
timeout 900 python -m pytest tests -q -m gpu -k "ozaki or config3 or mgpu_potrf" > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
