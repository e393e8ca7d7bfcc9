#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/autotune.py --kernel diffusion --assignment 0000000000000000 --tiles 512x16x1,512x32x1,1024x16x1,512x24x1 --budgets 220 --threads 256,384,512 --stages 2,3 --runs 20 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[8,1],[4,1],[6,1]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[true]}" \
  --out gpurun_out/tune_diffusion_v5.json > gpurun_out/tune_diffusion_v5.log 2>&1; grep BEST gpurun_out/tune_diffusion_v5.log
