#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -30 gpurun_out/r2_pytest_gpu.log
bash tools/runs/r2_run5.sh
