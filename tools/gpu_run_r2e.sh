#!/bin/bash
# final captures of the round: ncu of both packet kernels on the headline frame, then the default bench line
mkdir -p gpurun_out
timeout 300 python tools/primary_bench.py > gpurun_out/plain_primary.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_macro -s 3 -c 1 -f \
    -o gpurun_out/prof_macro_r2 python tools/primary_bench.py > gpurun_out/ncu_macro.log 2>&1
echo "ncu macro exit $?"; cat gpurun_out/plain_primary.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_packet -s 3 -c 1 -f \
    -o gpurun_out/prof_packet_r2b python tools/primary_bench.py > gpurun_out/ncu_packet.log 2>&1
echo "ncu packet exit $?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches', d['gpu_launches'])
print('roofline', {k:v for k,v in d['roofline'].items() if not isinstance(v,(dict,list))})
print('tile_work', d['roofline']['tile_work'])
for v in d.get('e2e_variants',[]): print('  ', v['output'][:60], v['value'], v['ms_per_step'])
for r in d.get('sweep',[]):
    if 'rows' in r: continue
    print('  sweep', r['workload'][:60], r['value'], r['unit'], r.get('e2e',{}).get('value'), r.get('kernel','')[:40])
PY
