#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q -k "cache" 2>&1 | tail -n 2
timeout 900 python tools/autotune.py --kernel diffusion --assignment 0000000000000000 --tiles 512x16x1 --budgets 220 --threads 384 --stages 3 --runs 20 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[4,1]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[true],\"STORE_CS\":[true],\"L2_ONCE\":[true,false]}" | grep -v BEST | cut -c1-30,150-420
timeout 900 python tools/autotune.py --kernel fastwaves --assignment 000000111 --tiles 128x8x40 --budgets 110 --threads 384 --stages 2 --runs 20 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[1,2]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[false],\"STORE_CS\":[true],\"L2_ONCE\":[true,false]}" | grep -v BEST | cut -c1-30,150-420
timeout 900 python tools/autotune.py --kernel advection --assignment 00000000 --tiles 512x8x1 --budgets 110 --threads 256 --stages 3 --runs 20 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[4,1]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[false],\"STORE_CS\":[true,false]}" | grep -v BEST | cut -c1-30,150-420
