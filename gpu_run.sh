#!/bin/bash
echo strided16; python tools/bounce_bench.py --modes 2,0 --rays 1048576 --stride 16 --steps 10 2>&1 | grep "mode"
echo full; python tools/bounce_bench.py --modes 2,1 --steps 5 2>&1 | grep "mode"
python tools/primary_bench.py 2>&1 | grep "mode"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
