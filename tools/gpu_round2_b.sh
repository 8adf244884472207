#!/bin/bash
# round 2, GPU call B: parity suite with the Gram/TMA covariance build + bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest_full.log 2>&1
tail -15 gpurun_out/r2b_pytest_full.log
grep -E "cond\(K\)|config|max rel err|split predict|predict \(4096" gpurun_out/r2b_pytest_full.log | head -60 > gpurun_out/r2b_pytest_errors.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
e=d['extra']
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
print(e['stage_ms_per_step'])
print({k:e[k] for k in e if k.startswith('predict') and not isinstance(e[k],dict)}, e['predict_stage_ms'])
c=e['config4']; print('config4 mean ms',c['mean_ms_all_points'],c['mean_stage_ms_rank0'],'var pts/s',c['var_points_per_s'],c['parity_split_vs_dense'])
print('config5',e['config5']['s_per_eval'],e['config5']['phase_ms'])
PY
tail -5 gpurun_out/r2b_bench.err
