#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --config5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'])
for s in d['sweep']: print(s['workload'][:50], s.get('value'), s.get('ms_per_step'), s.get('e2e',{}).get('value'), s.get('rows',[{}])[-1])
PY
tail -3 gpurun_out/bench.err
