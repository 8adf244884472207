#!/bin/bash
# round 2, GPU call AO: ncu launch list of one evaluation on the last commit (time + DRAM bytes per launch)
mkdir -p gpurun_out
python tools/prof_eval_n32768.py > gpurun_out/r2ao_eval.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2ao_launches.csv python tools/prof_eval_n32768.py > gpurun_out/r2ao_ncu1.log 2>&1
cat gpurun_out/r2ao_eval.log; gzip -f gpurun_out/r2ao_launches.csv; ls -la gpurun_out/r2ao_launches.csv.gz
