#!/bin/bash
# round 2, GPU call Z: final check of the committed state (gpu tests, smoke, bench with the sharded extras)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; tail -3 gpurun_out/r2z_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['scale'], d['clocks'])
print(d['extra']['stage_ms_per_step'])
for k in ('config4','config5'):
    print(k, json.dumps(d['extra'].get(k))[:700])
PY
