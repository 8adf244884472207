#!/bin/bash
# round 2, GPU call AL (2 GPUs): both bench arms under torchrun exactly as the driver launches them, on the last commit
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2al_bench_n2.json 2> gpurun_out/r2al_bench_n2.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2al_bench_n2.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['n_gpus'])
for k in ('config4','config5','config5_single_process'):
    v=d['extra'].get(k); print(k, json.dumps(v)[:500] if v else None)
PY
tail -3 gpurun_out/r2al_bench_n2.err
GPR_REF_BUDGET_S=60 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
