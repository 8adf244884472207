#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ozaki_lauum_probe.py > gpurun_out/r2k_probe.log 2>&1; tail -8 gpurun_out/r2k_probe.log
