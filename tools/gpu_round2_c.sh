#!/bin/bash
# round 2, GPU call C: parity suite (all failures), covariance kernels timed + ncu --set full
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest_full.log 2>&1
tail -12 gpurun_out/r2c_pytest_full.log
grep -E "cond\(K\)|config|max rel err|split predict|predict \(4096" gpurun_out/r2c_pytest_full.log | head -80 > gpurun_out/r2c_pytest_errors.log
python tools/prof_cov.py 8192 16384 > gpurun_out/r2c_prof_cov.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kbuild|grad_reduce' -c 12 -o gpurun_out/r2c_cov -f python tools/prof_cov.py 8192 16384 > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_prof_cov.log; tail -3 gpurun_out/r2c_ncu.log
