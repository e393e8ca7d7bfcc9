#!/bin/bash
# round-end style validation: smoke, the whole GPU suite, the bench (both arms), and the ncu launch list
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" > gpurun_out/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/status.txt
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/status.txt
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -n 2 gpurun_out/smoke.log gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-3000
