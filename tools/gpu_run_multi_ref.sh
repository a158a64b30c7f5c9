#!/bin/bash
# the driver's two arms at N GPUs, launched the way the driver launches them
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?"
tail -c 600 gpurun_out/bench_ref_n$N.json
bash tools/gpu_run_multi.sh $N
