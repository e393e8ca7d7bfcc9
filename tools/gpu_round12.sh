#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python tools/autotune.py --kernel fastwaves --assignment 000000111 --tiles 128x8x80,256x8x80,256x4x80,128x16x80,512x4x80,128x8x40 --budgets 110,220 --threads 256,384 --runs 3 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[1,1],[2,1],[4,1]],\"FUSE_HALO\":[true],\"SHUFFLE\":[false],\"STREAM_K\":[true],\"PREFETCH\":[1,2,3]}" \
  --out gpurun_out/tune_fastwaves_stream2.json > gpurun_out/tune_fastwaves_stream2.log 2>&1; grep BEST gpurun_out/tune_fastwaves_stream2.log
