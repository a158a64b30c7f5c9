#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}   # usage: gpurun --gpus N -- bash tools/gpu_run_multi.sh N
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d.get('render_path',{}).get('value'), d['config'].get('assembled_frame_check','')[-4:], d['clocks'].get('samples'))
PY
