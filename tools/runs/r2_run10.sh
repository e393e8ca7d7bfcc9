#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands5.sh
rm -f gpurun_out/r2_ab_fastpath.log
for k in diffusion fastwaves advection; do
  echo "== $k" >> gpurun_out/r2_ab_fastpath.log
  timeout 300 python tools/ab_test.py --check $k "$F0" "$F1" "$F2" "$F3" 2>&1 | grep -v "^loading" >> gpurun_out/r2_ab_fastpath.log
done
cat gpurun_out/r2_ab_fastpath.log
