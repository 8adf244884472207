#!/bin/bash
# round 2, GPU call Q: exactly what the driver runs at round end (1 GPU): gpu tests, smoke, reference arm, bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
( time timeout 1790 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_reference.json 2> gpurun_out/r2q_reference.err ) 2>&1 | tail -3
tail -6 gpurun_out/r2q_reference.err; tail -c 1800 gpurun_out/r2q_reference.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline'].get('calibrated'))
PY
