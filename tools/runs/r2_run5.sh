#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands3.sh
timeout 600 python -X faulthandler tools/ab_test.py --check fastwaves "$BASE" "$P0" "$P1" "$P2" "$P3" "$P4" "$P5" "$P6" "$P7" 2>&1 | grep -v "^loading" > gpurun_out/r2_col_ab3.log
echo "ab rc $?" >> gpurun_out/r2_col_ab3.log
cat gpurun_out/r2_col_ab3.log | grep "median\|differs\|rc" | sed 's/"_clean.*"COLUMN"//'
