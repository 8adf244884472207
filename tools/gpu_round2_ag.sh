#!/bin/bash
# round 2, GPU call AG: bench line with the restructured roofline object (INT8 pipe roof from MEASURED_PEAKS.json, DGEMM roof beside it)
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; tail -3 gpurun_out/r2ag_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ag_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
print(json.dumps(d['roofline'])[:3000])
PY
