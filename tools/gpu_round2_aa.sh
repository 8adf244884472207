#!/bin/bash
# round 2, GPU call AA: nine-digit INT8 form of the block-cyclic lauum (mapped B rows): parity tests + one-rank timings at N = 32768
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -s -m gpu -k "mgpu or config5 or dist" > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log; grep "mgpu INT8" gpurun_out/r2aa_pytest.log | head -16
for nb in 1024 2048; do
  for la in 0 9; do
    echo "== nb=$nb GPR_OZAKI_LAUUM=$la"
    GPR_OZAKI_LAUUM=$la timeout 600 python tools/config5.py --n 32768 --gpus 1 --nb $nb --evals 2 2>&1 | tail -1 | cut -c1-600
  done
done > gpurun_out/r2aa_nb.log 2>&1
cat gpurun_out/r2aa_nb.log
