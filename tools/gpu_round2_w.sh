#!/bin/bash
# round 2, GPU call W: block-cyclic driver on one GPU (one rank, N = 32768, D = 16): block width x INT8 route of the rank-nb updates of potrf
mkdir -p gpurun_out
for nb in 1024 2048 4096; do
  for oz in 0 -1; do
    echo "== nb=$nb GPR_OZAKI=$oz"
    GPR_OZAKI=$oz timeout 600 python tools/config5.py --n 32768 --gpus 1 --nb $nb --evals 2 2>&1 | tail -1 | cut -c1-700
  done
done > gpurun_out/r2w_nb.log 2>&1
cat gpurun_out/r2w_nb.log
