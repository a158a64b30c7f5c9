#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench3.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'])
r=d['roofline']; print('kernel_ms',r['kernel_ms'],'per_ray',r['per_ray'])
for s in d['sweep']: print(s['workload'][:50], s['value'], s['ms_per_step'], s['per_ray']['V_i'], s['per_ray']['V_s'])
PY
tail -3 gpurun_out/bench3.err
