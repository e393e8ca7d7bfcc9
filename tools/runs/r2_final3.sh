#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fullsize_gpu.py tests/test_distributed.py tests/test_protocol_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_tuned.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/r2_pytest_tuned.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1gpu_i.json 2> gpurun_out/r2_bench_1gpu_i.err
echo "bench rc $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu_i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['roofline']['kernel'], d['roofline']['frac'], {k[:12]: (v['kernel_ms'], round(v['GBps']/6545.6,3)) for k, v in d['breakdown'].items()}, 'sustained', d['sustained']['ms_per_step'], d['sustained']['frac'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
