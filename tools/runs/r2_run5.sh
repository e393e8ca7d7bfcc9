#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands3.sh
for promo in 3 2 0; do
echo "== L2 promotion level $promo" >> gpurun_out/r2_col_ab4.log
ABSINTHE_TMA_L2_PROMOTION=$promo timeout 600 python -X faulthandler tools/ab_test.py fastwaves "$BASE" "$P0" "$P1" "$P2" "$P3" "$P4" "$P5" "$P6" "$P7" 2>&1 | grep -v "^loading" >> gpurun_out/r2_col_ab4.log
done
cat gpurun_out/r2_col_ab4.log | grep "median\|differs\|rc\|==" | sed 's/"_clean.*"WAIT_SLEEP": 32, //'
