#!/bin/bash
# round 2, second GPU call: all GPU tests, packet-kernel variants, the N=1 bench line, ncu captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest r2 exit $?" >> gpurun_out/pytest_r2.log
tail -15 gpurun_out/pytest_r2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
bash tools/gpu_run_variants.sh "python tools/primary_bench.py" > /dev/null 2>&1
cat gpurun_out/variants.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms'], d['roofline']['kernel_ms_ids_only'],'bound',d['roofline']['bound'],'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for v in d['e2e_variants']: print('  ', v['output'][:50], v['value'], v['ms_per_step'])
for r in d['sweep']:
    if 'rows' in r: continue
    print('  sweep', r['workload'][:60], r['value'], r['unit'], r.get('e2e',{}).get('value'))
print(d['clocks']); print(d['cpu_baseline'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_packet -s 3 -c 1 -f \
    -o gpurun_out/prof_packet_r2a python tools/primary_bench.py > gpurun_out/ncu_packet.log 2>&1
echo "ncu packet exit $?"
