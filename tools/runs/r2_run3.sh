#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands1.sh
timeout 600 python -X faulthandler tools/ab_test.py --check fastwaves "$BASE" "$A" "$C" "$D" "$F" "$G" "$H" "$I" "$J" > gpurun_out/r2_col_ab1.log 2>&1
echo "ab rc $?" >> gpurun_out/r2_col_ab1.log
tail -40 gpurun_out/r2_col_ab1.log
