#!/bin/bash
mkdir -p gpurun_out
. tools/runs/r2_cands6.sh
timeout 600 python tools/ab_test.py --check diffusion "$D0" "$D1" "$D2" "$D3" "$D4" "$D5" "$D6" "$D8" "$D9" "$D10" "$D11" 2>&1 | grep -v "^loading" > gpurun_out/r2_ab_diffusion.log
cat gpurun_out/r2_ab_diffusion.log
