#!/bin/bash
# round 2, GPU call T: the round-end sequence on the final defaults (gpu tests, smoke, reference arm, bench with the sharded paths),
# then the ncu launch list of one N = 32768 evaluation and one --set full capture of the INT8 product kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; tail -2 gpurun_out/r2t_smoke.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t_reference.json 2> gpurun_out/r2t_reference.err ) 2>&1 | tail -3
tail -4 gpurun_out/r2t_reference.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['scale'])
print(d['extra']['stage_ms_per_step'])
r=json.loads([l for l in open('gpurun_out/r2t_reference.json') if l.startswith('{')][-1])
print('reference',r['value'],r['ms_per_step'],r['config']['step_N'],r['cpu_baseline']['scale'])
PY
python tools/prof_eval_n32768.py > gpurun_out/r2t_eval.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2t_launches.csv python tools/prof_eval_n32768.py > gpurun_out/r2t_ncu1.log 2>&1
cat gpurun_out/r2t_eval.log; gzip -f gpurun_out/r2t_launches.csv
python tools/ozaki_prof.py > gpurun_out/r2t_ozprof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'oz_gemm' -c 4 -o gpurun_out/r2t_oz -f python tools/ozaki_prof.py > gpurun_out/r2t_ncu2.log 2>&1
cat gpurun_out/r2t_ozprof.log; tail -3 gpurun_out/r2t_ncu2.log
ncu -i gpurun_out/r2t_oz.ncu-rep --page raw --csv > gpurun_out/r2t_oz.raw.csv 2>/dev/null
ls -la gpurun_out/r2t_oz.ncu-rep; if [ $(stat -c %s gpurun_out/r2t_oz.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/r2t_oz.ncu-rep; fi
