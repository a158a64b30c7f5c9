#!/bin/bash
# first GPU pass: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
