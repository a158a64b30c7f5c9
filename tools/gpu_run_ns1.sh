#!/bin/bash
# NS-1 experiments: top-of-tree residency for the ray-buffer kernel -- shared-memory top levels (build variants) and
# a persisting-L2 window (run-time option), at 1 M spheres (tree fits L2) and 10 M spheres (tree 320 MB > 126 MB L2)
mkdir -p gpurun_out
: > gpurun_out/ns1.log
for lib in cs201_sah-bvh_ray_tracer_b200/libsahbvh_b200.so cs201_sah-bvh_ray_tracer_b200/build/variants/libsahbvh_top21.so cs201_sah-bvh_ray_tracer_b200/build/variants/libsahbvh_top85.so; do
  echo "== $lib, 1 M spheres, 2^24 bounce rays" >> gpurun_out/ns1.log
  SAHBVH_LIB=$PWD/$lib timeout 300 python tools/bounce_bench.py --modes 2 2>&1 | grep -v "^build+upload" >> gpurun_out/ns1.log
done
echo "== default library, 1 M spheres, L2 window" >> gpurun_out/ns1.log
timeout 300 python tools/bounce_bench.py --modes 2 --l2-mb 1,8,30 2>&1 | grep -v "^build+upload" >> gpurun_out/ns1.log
echo "== default library, 10 M spheres, 2^24 bounce rays, L2 window" >> gpurun_out/ns1.log
timeout 600 python tools/bounce_bench.py --spheres 10000000 --modes 2 --l2-mb 8,32,64,96 2>&1 | grep -v "^build+upload" >> gpurun_out/ns1.log
echo "== top85 variant, 10 M spheres" >> gpurun_out/ns1.log
SAHBVH_LIB=$PWD/cs201_sah-bvh_ray_tracer_b200/build/variants/libsahbvh_top85.so timeout 600 python tools/bounce_bench.py --spheres 10000000 --modes 2 2>&1 | grep -v "^build+upload" >> gpurun_out/ns1.log
cat gpurun_out/ns1.log
