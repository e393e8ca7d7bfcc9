#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "store-ptr" > gpurun_out/r2_storeptr_parity.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/r2_storeptr_parity.log
for K in diffusion fastwaves advection; do
  timeout 600 python tools/ab_test.py --check $K '{}' '{"STORE_PTR":true}' '{}' '{"STORE_PTR":true}' > gpurun_out/r2_ab_storeptr_$K.txt 2>&1
  echo "$K rc $?"; grep -v "^loading" gpurun_out/r2_ab_storeptr_$K.txt
done
