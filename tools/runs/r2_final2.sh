#!/bin/bash
# the driver's two arms on a fresh box: reference first, then ours (default flags), then the cluster launcher test
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_g.json 2> gpurun_out/r2_bench_reference_g.err
echo "reference rc $?"; tail -1 gpurun_out/r2_bench_reference_g.json | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_g.json 2> gpurun_out/r2_bench_1gpu_g.err
echo "bench rc $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu_g.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['roofline']['kernel'], d['roofline']['frac'], {k[:12]: (v['kernel_ms'], round(v['GBps']/6545.6,3)) for k, v in d['breakdown'].items()}, 'sustained', d['sustained']['ms_per_step'], d['sustained']['frac'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
timeout 600 python -m pytest tests/test_launcher_gpu.py -m gpu -q -x 2>&1 | tail -2
