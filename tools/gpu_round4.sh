#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/status.txt
timeout 2000 python tools/autotune.py --kernel diffusion --assignment 0000000000000000 --tiles 256x8x1,256x16x1,256x8x2,512x4x1,512x8x1,128x16x1 --budgets 110,220 --threads 256,512 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[2,1],[4,1]],\"DIRECT\":[\"auto\",\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[true,false]}" \
  --out gpurun_out/tune_diffusion_v3.json > gpurun_out/tune_diffusion_v3.log 2>&1; echo "tune exit $?" >> gpurun_out/status.txt
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/status.txt; grep BEST gpurun_out/tune_diffusion_v3.log
