#!/bin/bash
# round 2, GPU call S: nine-digit W^T W on the INT8 tensor cores as the default (128x256 tenth-diagonal tiles A/B), whole GPU suite with the
# default policy and with the automatic route lowered to 128-sized products, bench
mkdir -p gpurun_out
timeout 900 python tools/ozaki_lauum9.py > gpurun_out/r2s_lauum9.log 2>&1; tail -14 gpurun_out/r2s_lauum9.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; tail -4 gpurun_out/r2s_pytest.log
GPR_OZAKI_MIN=128 timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/r2s_pytest_min128.log 2>&1; tail -6 gpurun_out/r2s_pytest_min128.log
grep -E "cond\(K\)|FAILED" gpurun_out/r2s_pytest_min128.log | head -30
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-sharded > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2s_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:300])
print(d['extra']['stage_ms_per_step'])
PY
