#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -12 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1gpu_c.json 2> gpurun_out/r2_bench_1gpu_c.err
echo "bench rc $?"; cut -c1-600 gpurun_out/r2_bench_1gpu_c.json; tail -3 gpurun_out/r2_bench_1gpu_c.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "reference rc $?"; cat gpurun_out/r2_bench_reference.json | cut -c1-900
