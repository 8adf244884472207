#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/prof_eval_n32768.py > gpurun_out/r2l_eval.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2l_launches.csv python tools/prof_eval_n32768.py > gpurun_out/r2l_ncu1.log 2>&1
cat gpurun_out/r2l_eval.log
ncu --set full --clock-control none -k regex:'oz_gemm' -s 40 -c 2 -o gpurun_out/r2l_oz -f python tools/prof_eval_n32768.py > gpurun_out/r2l_ncu2.log 2>&1
ncu -i gpurun_out/r2l_oz.ncu-rep --page raw --csv > gpurun_out/r2l_oz.raw.csv 2>/dev/null
ls -la gpurun_out/r2l_oz.ncu-rep gpurun_out/r2l_launches.csv; if [ $(stat -c %s gpurun_out/r2l_oz.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/r2l_oz.ncu-rep; fi
gzip -f gpurun_out/r2l_launches.csv
