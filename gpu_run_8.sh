#!/bin/bash
mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
python - <<PY
import json
for f in ('gpurun_out/bench_n8.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d.get('render_path',{}).get('value'), d['config'].get('assembled_frame_check','')[-4:], d['clocks'])
PY
