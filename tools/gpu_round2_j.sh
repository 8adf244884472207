#!/bin/bash
# parity suite with the automatic INT8 route + default bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2j_pytest.log 2>&1
tail -6 gpurun_out/r2j_pytest.log; grep -E "cond\(K\)|config |config 2|predict \(4096|split predict|ozaki |INT8|FAILED" gpurun_out/r2j_pytest.log | head -60 > gpurun_out/r2j_pytest_errors.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench.json') if l.startswith('{')][-1])
e=d['extra']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'peak',d['roofline']['peak'])
print(e['stage_ms_per_step'], 'chol',e['cholesky_tflops'],'inv',e['inverse_tflops'])
print('dmma_only',e['dmma_only']); print('int8',d['roofline']['int8_route'])
print({k:e[k] for k in e if k.startswith('predict') and not isinstance(e[k],dict)}, e['predict_stage_ms'])
c=e['config4']; print('config4 mean ms',c['mean_ms_all_points'],c['mean_stage_ms_rank0'],'var pts/s',c['var_points_per_s'],c['parity_split_vs_dense'])
print('config5',e['config5']['s_per_eval'],e['config5']['phase_ms']); print(e['config3_ext'], e['train_free_running']); print(d['cpu_baseline'])
PY
