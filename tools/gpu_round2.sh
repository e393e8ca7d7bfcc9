#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/status.txt
FL='["ufli","uflj","vfli","vflj","wfli","wflj","ppfli","ppflj"]'
timeout 1500 python tools/autotune.py --kernel diffusion --assignment 0000000000000000 --tiles 128x16x1,256x8x1,64x32x1 --budgets 110,220 --threads 256,512 \
  --sweep "{\"INLINE\":[\"all\",$FL],\"BLOCK\":[[1,1],[2,1],[4,1]],\"DIRECT\":[\"auto\"],\"FUSE_HALO\":[true]}" \
  --out gpurun_out/tune_diffusion1.json > gpurun_out/tune_diffusion1.log 2>&1; echo "tune1 exit $?" >> gpurun_out/status.txt
timeout 1500 python tools/autotune.py --kernel diffusion --assignment 0000111122223333 --tiles 128x16x1,64x32x1 --budgets 110 --threads 256,512 \
  --sweep "{\"INLINE\":[\"all\",$FL,\"none\"],\"BLOCK\":[[1,1],[2,1],[4,1]],\"DIRECT\":[\"auto\",\"none\"],\"FUSE_HALO\":[true]}" \
  --out gpurun_out/tune_diffusion4b.json > gpurun_out/tune_diffusion4b.log 2>&1; echo "tune4 exit $?" >> gpurun_out/status.txt
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/status.txt; grep BEST gpurun_out/tune_diffusion*.log
