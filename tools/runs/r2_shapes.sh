#!/bin/bash
mkdir -p gpurun_out
source tools/runs/r2_shapes_cands.sh
timeout 900 python tools/ab_test.py --check fastwaves "${CANDS[@]}" > gpurun_out/r2_ab_shapes.txt 2>&1
echo "ab rc $?"; grep -v "^loading" gpurun_out/r2_ab_shapes.txt | tail -40
