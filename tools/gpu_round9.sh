#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/autotune.py --kernel fastwaves --assignment 000000111 --tiles 128x8x16,128x8x20,256x8x16,128x16x16,128x8x40,128x8x80 --budgets 110,220 --threads 256,384 --stages 2,3 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[1,4],[1,2],[2,2],[2,4]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[false]}" \
  --out gpurun_out/tune_fastwaves_v5.json > gpurun_out/tune_fastwaves_v5.log 2>&1; grep BEST gpurun_out/tune_fastwaves_v5.log
