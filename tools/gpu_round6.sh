#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/status.txt
timeout 2400 python tools/autotune.py --kernel diffusion --assignment 0000000000000000 --tiles 512x8x1,256x16x1,512x16x1,256x8x1,1024x8x1 --budgets 110,220 --threads 256,384,512 --stages 2,3,4 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[2,1],[4,1]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[true,false]}" \
  --out gpurun_out/tune_diffusion_v4.json > gpurun_out/tune_diffusion_v4.log 2>&1; echo "tune exit $?" >> gpurun_out/status.txt
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/status.txt; grep BEST gpurun_out/tune_diffusion_v4.log
