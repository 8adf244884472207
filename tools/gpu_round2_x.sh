#!/bin/bash
# round 2, GPU call X: INT8 route of the block-cyclic trtri (GEMM_MAP_KUPTO) -- parity tests, then the one-rank driver at N = 32768 with
# potrf only / potrf + trtri on INT8 / all DMMA, block widths 1024 and 2048
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -s -m gpu -k "mgpu or config5 or dist" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log; grep "mgpu potrf" gpurun_out/r2x_pytest.log | head -12
for nb in 1024 2048; do
  for ph in 0 1 11; do
    echo "== nb=$nb phases=$ph"
    if [ $ph = 0 ]; then export GPR_OZAKI=0; else export GPR_OZAKI=-1; fi
    GPR_OZAKI_PHASES=$ph timeout 600 python tools/config5.py --n 32768 --gpus 1 --nb $nb --evals 2 2>&1 | tail -1 | cut -c1-600
  done
done > gpurun_out/r2x_nb.log 2>&1
cat gpurun_out/r2x_nb.log
