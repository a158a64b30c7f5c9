#!/bin/bash
# round 2: NS-1 experiments (tools/gpu_run_ns1.sh) + packet-kernel output/block-shape variants on the headline frame
mkdir -p gpurun_out
V=cs201_sah-bvh_ray_tracer_b200/build/variants
: > gpurun_out/pkt_variants.log
for lib in cs201_sah-bvh_ray_tracer_b200/libsahbvh_b200.so $V/libsahbvh_warpout.so $V/libsahbvh_w2.so $V/libsahbvh_w2warpout.so; do
  echo "== $lib" >> gpurun_out/pkt_variants.log
  SAHBVH_LIB=$PWD/$lib timeout 300 python tools/primary_bench.py 2>&1 | grep -v "^build+upload" >> gpurun_out/pkt_variants.log
done
cat gpurun_out/pkt_variants.log
bash tools/gpu_run_ns1.sh
