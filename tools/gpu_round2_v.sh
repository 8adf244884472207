#!/bin/bash
# round 2, GPU call V: the round-end sequence on the final defaults (gpu tests, smoke, reference arm, bench with the sharded paths) + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v_smoke.log 2>&1; tail -2 gpurun_out/r2v_smoke.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2v_reference.json 2> gpurun_out/r2v_reference.err ) 2>&1 | tail -3
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['scale'])
print(d['extra']['stage_ms_per_step']); print(d['roofline']['int8_route'])
print({k:v for k,v in d['extra'].items() if not isinstance(v,dict)})
for k in ('config4','config5','config3_ext','train_free_running','dmma_only'):
    print(k, json.dumps(d['extra'].get(k))[:600])
r=json.loads([l for l in open('gpurun_out/r2v_reference.json') if l.startswith('{')][-1])
print('reference',r['value'],r['ms_per_step'],r['config']['step_N'],r['cpu_baseline']['scale'])
PY
python tools/prof_eval_n32768.py > gpurun_out/r2v_eval.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2v_launches.csv python tools/prof_eval_n32768.py > gpurun_out/r2v_ncu1.log 2>&1
cat gpurun_out/r2v_eval.log; gzip -f gpurun_out/r2v_launches.csv
