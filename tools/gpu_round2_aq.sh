#!/bin/bash
# round 2, GPU call AQ (the last 2 GPU-minutes): ncu --set full of the two-window 8-digit INT8 kernels as single CTAs and as multicast cluster
# pairs on the same 4096^3 product: the counters behind "25 % fewer L2 reads, same time"
mkdir -p gpurun_out
timeout 30 python tools/oz_mc_prof.py 4096 > gpurun_out/r2aq_prof.log 2>&1 && \
timeout 80 ncu --set full --clock-control none --import-source on -k regex:'oz_gemm_win' -c 4 -o gpurun_out/r2aq_mc -f python tools/oz_mc_prof.py 4096 > gpurun_out/r2aq_ncu.log 2>&1
cat gpurun_out/r2aq_prof.log; tail -3 gpurun_out/r2aq_ncu.log; ls -la gpurun_out/r2aq_mc.ncu-rep
