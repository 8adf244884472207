#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ozaki_check.py quick > gpurun_out/r2p_check.log 2>&1; tail -4 gpurun_out/r2p_check.log
timeout 300 python tools/ozaki_probe.py > gpurun_out/r2p_probe.log 2>&1; cat gpurun_out/r2p_probe.log
timeout 300 python tools/prof_eval_n32768.py > gpurun_out/r2p_eval.log 2>&1; cat gpurun_out/r2p_eval.log
