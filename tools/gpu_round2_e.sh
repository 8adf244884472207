#!/bin/bash
# round 2, GPU call E: full parity suite, default bench, ncu of the two T,N GEMM rings, CPU reference arm (N = 16384 calibration)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2e_pytest_full.log 2>&1
tail -8 gpurun_out/r2e_pytest_full.log
grep -E "cond\(K\)|config|max rel err|split predict|predict \(4096" gpurun_out/r2e_pytest_full.log | head -80 > gpurun_out/r2e_pytest_errors.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 1500 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
python tools/gemm_prof.py 4096 > gpurun_out/r2e_gemm_prof.log 2>&1 && \
ncu --set full --clock-control none -k regex:'dgemm128' -c 4 -o gpurun_out/r2e_gemm -f python tools/gemm_prof.py 4096 > gpurun_out/r2e_ncu.log 2>&1
cat gpurun_out/r2e_gemm_prof.log
ncu -i gpurun_out/r2e_gemm.ncu-rep --page raw --csv > gpurun_out/r2e_gemm.raw.csv 2>/dev/null
ls -la gpurun_out/r2e_gemm.ncu-rep; if [ $(stat -c %s gpurun_out/r2e_gemm.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/r2e_gemm.ncu-rep; fi
GPR_REF_BUDGET_S=1000 timeout 1700 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_reference.json 2> gpurun_out/r2e_reference.err
cat gpurun_out/r2e_reference.err | tail -8; tail -c 2500 gpurun_out/r2e_reference.json
