#!/bin/bash
# round 2, GPU call D: parity suite, covariance kernels timed + ncu (small capture), TMA GEMM A/B
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest_full.log 2>&1
tail -12 gpurun_out/r2d_pytest_full.log
grep -E "cond\(K\)|config|max rel err|split predict|predict \(4096" gpurun_out/r2d_pytest_full.log | head -80 > gpurun_out/r2d_pytest_errors.log
python tools/prof_cov.py 8192 16384 > gpurun_out/r2d_prof_cov.log 2>&1; cat gpurun_out/r2d_prof_cov.log
python tools/gemm_ab.py > gpurun_out/r2d_gemm_ab.log 2>&1; cat gpurun_out/r2d_gemm_ab.log
python tools/prof_cov.py 8192 16384 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:'kbuild|grad_reduce' -c 3 -o gpurun_out/r2d_cov_a -f python tools/prof_cov.py 8192 16384 > gpurun_out/r2d_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'kbuild|grad_reduce' -s 4 -c 4 -o gpurun_out/r2d_cov_b -f python tools/prof_cov.py 8192 16384 >> gpurun_out/r2d_ncu.log 2>&1
for f in gpurun_out/r2d_cov_a gpurun_out/r2d_cov_b; do
  ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null
  ls -la $f.ncu-rep
  if [ $(stat -c %s $f.ncu-rep) -gt 20000000 ]; then rm -f $f.ncu-rep; fi
done
tail -3 gpurun_out/r2d_ncu.log
