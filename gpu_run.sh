#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
print(d['roofline']['tile_work'])
for s in d['sweep']: print(s['workload'][:60], s.get('value'), s.get('ms_per_step'), s.get('e2e',{}).get('value'))
PY
tail -3 gpurun_out/bench.err
