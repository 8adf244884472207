#!/bin/bash
# round 2, GPU call AD: deferred factorization status on the gradient path -- whole GPU suite, smoke, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ad_pytest.log 2>&1; tail -3 gpurun_out/r2ad_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2ad_smoke.log 2>&1; tail -2 gpurun_out/r2ad_smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-sharded > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ad_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
print(d['extra']['stage_ms_per_step'])
PY
