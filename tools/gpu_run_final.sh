#!/bin/bash
# round-end check on one GPU: all GPU tests, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -5 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches', d['gpu_launches'])
print('roofline', {k:v for k,v in d['roofline'].items() if not isinstance(v,(dict,list,str))})
for v in d.get('e2e_variants',[]): print('  ', v['output'][:60], v['value'], v['ms_per_step'])
for r in d.get('sweep',[]):
    if 'rows' in r: continue
    print('  sweep', r['workload'][:70], r['value'], r['unit'], r.get('e2e',{}).get('value'))
print(d['clocks']); print(d['cpu_baseline'])
r=json.loads(open('gpurun_out/bench_ref_n1.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('unit'), r.get('cpu_baseline'))
PY
