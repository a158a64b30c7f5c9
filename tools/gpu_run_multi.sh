#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}   # usage: gpurun --gpus N -- bash tools/gpu_run_multi.sh N [extra bench args]
shift
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -i "numa\|^CPU(s)\|model name" >> gpurun_out/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'].get('samples'))
print('roofline', {k:v for k,v in d['roofline'].items() if not isinstance(v,(dict,list,str))})
print('config', json.dumps(d['config'])[:1500])
for k in ('frame_consumer','render_path'):
    if k in d: print(k, json.dumps(d[k])[:600])
for r in d.get('rows',[]) + d.get('sweep',[]):
    print('  row', json.dumps(r)[:700])
PY
