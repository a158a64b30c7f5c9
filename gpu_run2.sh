#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clk',d['clocks'])
r=d['roofline']; print('kernel_ms',r['kernel_ms'],'l2frac',r['frac'],'fp32',r['fp32'],'peaks',r['measured_peaks'])
for s in d['sweep']: print(s['workload'][:50], s['value'], s['ms_per_step'], s['per_ray']['V_i'], s.get('e2e'))
print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench2.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary -s 1 -c 1 -o gpurun_out/prof_primary_r1 -f \
   python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log
