#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands2.sh
timeout 600 python -X faulthandler tools/ab_test.py --check fastwaves "$BASE" "$H0" "$H1" "$H2" "$H4" "$M" "$K" "$L" "$N" 2>&1 | grep -v "^loading" > gpurun_out/r2_col_ab2.log
echo "ab rc $?" >> gpurun_out/r2_col_ab2.log
cut -c1-40,160- gpurun_out/r2_col_ab2.log | tail -24
