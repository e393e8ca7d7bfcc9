#!/bin/bash
# two GPUs: the multi-GPU tests, bench --check across two slabs, scaling with both exchanges
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_2gpu_list.txt
timeout 900 python -m pytest tests/test_distributed.py tests/test_launcher_gpu.py -m gpu -q > gpurun_out/r2_pytest_2gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_2gpu.log; tail -5 gpurun_out/r2_pytest_2gpu.log
for ex in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --check --exchange $ex > gpurun_out/r2_check_2gpu_$ex.json 2> gpurun_out/r2_check_2gpu_$ex.err
  echo "check $ex rc $?"; tail -1 gpurun_out/r2_check_2gpu_$ex.json | cut -c1-400
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 50 --warmup 5 --exchange $ex > gpurun_out/r2_bench_2gpu_$ex.json 2> gpurun_out/r2_bench_2gpu_$ex.err
  echo "bench $ex rc $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_2gpu_$ex.json").read().strip().splitlines()[-1])
    print("$ex", d["value"], d["ms_per_step"], d["sustained"]["ms_per_step"], {k: (v["kernel_ms"], v["halo_ms"]) for k, v in d["breakdown"].items()}, d["e2e"]["ms_per_step"])
except Exception as e:
    print("parse failed", e)
PY
  tail -3 gpurun_out/r2_bench_2gpu_$ex.err
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-extras > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu_b.json').read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'], d['sustained']['ms_per_step'])"
