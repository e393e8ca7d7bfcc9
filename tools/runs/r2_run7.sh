#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -12 gpurun_out/r2_pytest_gpu.log
. tools/runs/r2_cands4.sh
for k in fastwaves diffusion advection; do
  echo "== $k" >> gpurun_out/r2_ab_poll_fmad.log
  timeout 300 python tools/ab_test.py $k "$B0" "$B1" "$B2" 2>&1 | grep -v "^loading" >> gpurun_out/r2_ab_poll_fmad.log
done
cat gpurun_out/r2_ab_poll_fmad.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1gpu_a.json 2> gpurun_out/r2_bench_1gpu_a.err
echo "bench rc $?"; cat gpurun_out/r2_bench_1gpu_a.json; tail -3 gpurun_out/r2_bench_1gpu_a.err
