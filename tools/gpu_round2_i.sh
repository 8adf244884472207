#!/bin/bash
# parity suite with the INT8-tensor-core products switched on for every context (potrf, trtri, prediction solves)
set -x
mkdir -p gpurun_out
GPR_OZAKI=8 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_pytest_ozaki8.log 2>&1
tail -6 gpurun_out/r2i_pytest_ozaki8.log; grep -E "cond\(K\)|config |config 2|predict \(4096|split predict" gpurun_out/r2i_pytest_ozaki8.log | head -40
GPR_OZAKI=8 GPR_OZAKI_MIN=128 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_pytest_ozaki8_min128.log 2>&1
tail -6 gpurun_out/r2i_pytest_ozaki8_min128.log; grep -E "cond\(K\)|config |config 2|predict \(4096|split predict|FAILED" gpurun_out/r2i_pytest_ozaki8_min128.log | head -60
