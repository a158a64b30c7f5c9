#!/bin/bash
echo base; python tools/bounce_bench.py --modes 2 --steps 5 2>&1 | grep "mode 2"
for v in A B C; do echo "variant $v"; SAHBVH_LIB=$PWD/cs201_sah-bvh_ray_tracer_b200/build/var_$v.so python tools/bounce_bench.py --modes 2 --steps 5 2>&1 | grep "mode 2"; done
