#!/bin/bash
# one `ncu --set full` capture of the packet kernel (headline frame) and of the wide ray-buffer kernel (config 4);
# each harness is first run once without ncu and must exit 0 (B200_PROFILING.md)
mkdir -p gpurun_out
tag=${1:-r1e}
timeout 300 python tools/primary_bench.py > gpurun_out/plain_primary.log 2>&1 || { echo "plain primary run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_packet -s 3 -c 1 -f \
    -o gpurun_out/prof_packet_$tag python tools/primary_bench.py > gpurun_out/ncu_packet.log 2>&1
echo "ncu packet exit $?"
if [ "$2" = "wide" ]; then
  timeout 300 python tools/bounce_bench.py --modes 2 > gpurun_out/plain_bounce.log 2>&1 || { echo "plain bounce run failed"; exit 1; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_wide -s 1 -c 1 -f \
      -o gpurun_out/prof_wide_$tag python tools/bounce_bench.py --modes 2 > gpurun_out/ncu_wide.log 2>&1
  echo "ncu wide exit $?"
fi
ls -la gpurun_out/*.ncu-rep | tail -3
