#!/bin/bash
# round 2, first GPU call: the new GPU tests, then the old ones, then the e2e probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest r2 exit $?" >> gpurun_out/pytest_r2.log
tail -25 gpurun_out/pytest_r2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python tools/e2e_primary_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "probe exit $?"
cat gpurun_out/e2e_probe.log
