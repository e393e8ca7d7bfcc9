#!/bin/bash
mkdir -p gpurun_out
for A in 000000111 000000000; do
timeout 2000 python tools/autotune.py --kernel fastwaves --assignment $A --tiles 256x8x8,128x8x16,512x4x4,256x4x10,128x16x8,256x8x80 --budgets 110,220 --threads 256 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[1,2],[1,4],[2,2],[2,1]],\"DIRECT\":[\"auto\",\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[true,false]}" \
  --out gpurun_out/tune_fastwaves_$A.json > gpurun_out/tune_fastwaves_$A.log 2>&1; echo "tune $A exit $?" >> gpurun_out/status.txt
done
grep BEST gpurun_out/tune_fastwaves_*.log
bash tools/gpu_ncu.sh diffusion absinthe_group1
