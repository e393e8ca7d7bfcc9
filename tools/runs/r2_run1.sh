#!/bin/bash
# round 2, GPU call 1: parity of the column kernels on the zoo + A/B of fused fast-waves column candidates
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "column" > gpurun_out/r2_col_parity.log 2>&1
echo "parity rc $?" >> gpurun_out/r2_col_parity.log
. tools/runs/r2_cands1.sh
timeout 600 python tools/ab_test.py --check fastwaves "$BASE" "$A" "$B" "$C" "$D" "$E" "$F" "$G" "$H" > gpurun_out/r2_col_ab1.log 2>&1
echo "ab rc $?" >> gpurun_out/r2_col_ab1.log
tail -3 gpurun_out/r2_col_parity.log; cat gpurun_out/r2_col_ab1.log
