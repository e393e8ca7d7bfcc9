#!/bin/bash
# full GPU suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_f.json 2> gpurun_out/r2_bench_1gpu_f.err
echo "bench rc $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu_f.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], {k[:12]: (v['kernel_ms'], round(v['GBps']/6545.6,3)) for k, v in d['breakdown'].items()}, 'sustained', d['sustained']['ms_per_step'], d['sustained']['frac'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
