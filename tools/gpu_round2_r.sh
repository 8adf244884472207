#!/bin/bash
# round 2, GPU call R: nine-digit INT8 route for the W^T W product of the inverse (accuracy vs the N=32768 golden + stage times)
mkdir -p gpurun_out
timeout 900 python tools/ozaki_lauum9.py > gpurun_out/r2r_lauum9.log 2>&1; tail -20 gpurun_out/r2r_lauum9.log
