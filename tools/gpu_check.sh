#!/bin/bash
# First-contact GPU check: staged-load path first, TMA path second, each under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "staged" > gpurun_out/pytest_ldg.log 2>&1; echo "ldg exit $?" >> gpurun_out/status.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "not staged" > gpurun_out/pytest_tma.log 2>&1; echo "tma exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
tail -5 gpurun_out/pytest_ldg.log gpurun_out/pytest_tma.log gpurun_out/smoke.log; cat gpurun_out/status.txt
