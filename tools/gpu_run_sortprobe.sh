#!/bin/bash
# how much would ray reordering buy the wide kernel?  (host-side sort, timing of the kernel only)
mkdir -p gpurun_out
: > gpurun_out/sortprobe.log
for s in "" "--sort morton" "--sort morton_dir" "--sort shuffle"; do
  echo "== sort '$s'" >> gpurun_out/sortprobe.log
  timeout 400 python tools/bounce_bench.py --modes 2 --rays 8388608 $s 2>&1 | grep "^mode" >> gpurun_out/sortprobe.log
done
cat gpurun_out/sortprobe.log
