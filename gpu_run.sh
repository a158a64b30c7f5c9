#!/bin/bash
python tools/bounce_bench.py --modes 2 --steps 5 2>&1 | grep "mode 2"
python tools/primary_bench.py 2>&1 | grep "mode 1"
