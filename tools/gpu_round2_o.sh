#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ozaki_probe.py > gpurun_out/r2o_probe.log 2>&1; cat gpurun_out/r2o_probe.log
