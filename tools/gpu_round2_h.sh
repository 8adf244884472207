#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/ozaki_eval.py > gpurun_out/r2h_ozaki_eval.log 2>&1; echo rc=$?
cat gpurun_out/r2h_ozaki_eval.log | tail -20
