#!/bin/bash
# round 2, GPU call AH: split-predict mean on the INT8 tensor cores (Hadamard epilogue) + pinned staging download: parity tests, config 4 timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -s -m gpu -k "ozaki_route or split or config3 or cabi or golden" > gpurun_out/r2ah_pytest.log 2>&1; tail -3 gpurun_out/r2ah_pytest.log; grep "split mean" gpurun_out/r2ah_pytest.log | head -4
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; tail -3 gpurun_out/r2ah_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ah_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print(json.dumps(d['extra']['config4'])[:900])
PY
