#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/leaf_time.py > gpurun_out/r2aj_leaf.log 2>&1; cat gpurun_out/r2aj_leaf.log
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-sharded --no-cpu --no-predict > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; tail -2 gpurun_out/r2aj_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2aj_bench.json') if l.startswith('{')][-1])
print('value',d['value'],d['config'].get('arithmetic','')[:200])
PY
