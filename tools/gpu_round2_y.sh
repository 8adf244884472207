#!/bin/bash
# round 2, GPU call Y (8 GPUs): config 5 (N = 131072) with the INT8 route in potrf + trtri of the block-cyclic drivers, block width 1024 vs 2048
# (one process driving all devices), then bench.py under torchrun exactly as the driver launches it, with the better width
mkdir -p gpurun_out
for nb in 1024 2048; do
  timeout 600 python tools/config5.py --n 131072 --gpus 8 --nb $nb --evals 2 2>&1 | tail -1 > gpurun_out/r2y_config5_nb$nb.json
  cut -c1-500 gpurun_out/r2y_config5_nb$nb.json
done
best=$(python - <<'PY'
import json
t={}
for nb in (1024,2048):
    try:
        d=json.loads(open(f'gpurun_out/r2y_config5_nb{nb}.json').read()); t[nb]=min(d['s_per_eval'])
    except Exception: pass
print(min(t,key=t.get) if t else 1024)
PY
)
echo "best nb: $best"
GPR_MGPU_NB=$best timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err
tail -c 3000 gpurun_out/r2y_bench_n8.json; tail -5 gpurun_out/r2y_bench_n8.err
