#!/bin/bash
# round 2, GPU call AE: ncu --set full of the final INT8 kernels (eight-digit single pass, nine-digit two windows) + the new failure/recovery test
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "positive_definite or nine_digit or cabi" > gpurun_out/r2ae_pytest.log 2>&1; tail -3 gpurun_out/r2ae_pytest.log
python tools/ozaki_prof.py > gpurun_out/r2ae_ozprof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'oz_gemm' -c 3 -o gpurun_out/r2ae_oz -f python tools/ozaki_prof.py > gpurun_out/r2ae_ncu.log 2>&1
cat gpurun_out/r2ae_ozprof.log; tail -3 gpurun_out/r2ae_ncu.log
ncu -i gpurun_out/r2ae_oz.ncu-rep --page raw --csv > gpurun_out/r2ae_oz.raw.csv 2>/dev/null
ls -la gpurun_out/r2ae_oz.ncu-rep; if [ $(stat -c %s gpurun_out/r2ae_oz.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/r2ae_oz.ncu-rep; fi
