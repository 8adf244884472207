#!/bin/bash
# round 2, GPU call AK: source-level ncu profile of potrf_leaf_kernel (one warm launch)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:potrf_leaf -s 8 -c 1 -o gpurun_out/r2ak_leaf -f python tools/leaf_time.py > gpurun_out/r2ak_ncu.log 2>&1
tail -3 gpurun_out/r2ak_ncu.log
ncu -i gpurun_out/r2ak_leaf.ncu-rep --page source --csv > gpurun_out/r2ak_leaf.source.csv 2>/dev/null
ncu -i gpurun_out/r2ak_leaf.ncu-rep --page raw --csv > gpurun_out/r2ak_leaf.raw.csv 2>/dev/null
ls -la gpurun_out/r2ak_leaf.*
