#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ozaki_check.py > gpurun_out/r2g_ozaki.log 2>&1; echo rc=$?
cat gpurun_out/r2g_ozaki.log | tail -40
