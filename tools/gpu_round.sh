#!/bin/bash
# tests + bench + a first tile sweep, each under its own timeout; logs into gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/status.txt
timeout 900 python tools/autotune.py --kernel diffusion --assignment 0000111122223333 --tiles 128x16x1,256x8x1,128x32x1,64x32x1,128x16x4,1024x4x1 --budgets 110,220 --threads 256,512 --out gpurun_out/tune_diffusion4.json > gpurun_out/tune_diffusion4.log 2>&1; echo "tune exit $?" >> gpurun_out/status.txt
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/status.txt; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; grep BEST gpurun_out/tune_diffusion4.log
