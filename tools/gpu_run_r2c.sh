#!/bin/bash
# round 2, re-entry: all GPU tests, the N=1 bench line, ncu launch list and full captures of both hot kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -8 gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'launches', d['gpu_launches'])
print('roofline', {k:v for k,v in d['roofline'].items() if not isinstance(v,(dict,list))})
for v in d.get('e2e_variants',[]): print('  ', v['output'][:60], v['value'], v['ms_per_step'])
for r in d.get('sweep',[]):
    if 'rows' in r: continue
    print('  sweep', r['workload'][:70], r['value'], r['unit'], r.get('e2e',{}).get('value'))
print(d['clocks']); print(d['cpu_baseline'])
PY
timeout 300 python tools/primary_bench.py > gpurun_out/plain_primary.log 2>&1; cat gpurun_out/plain_primary.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_packet -s 3 -c 1 -f \
    -o gpurun_out/prof_packet_r2a python tools/primary_bench.py > gpurun_out/ncu_packet.log 2>&1
echo "ncu packet exit $?"
timeout 300 python tools/bounce_bench.py --modes 2 > gpurun_out/plain_bounce.log 2>&1; cat gpurun_out/plain_bounce.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_wide -s 1 -c 1 -f \
    -o gpurun_out/prof_wide_r2a python tools/bounce_bench.py --modes 2 > gpurun_out/ncu_wide.log 2>&1
echo "ncu wide exit $?"
ls -la gpurun_out/*.ncu-rep
