#!/bin/bash
# weak scaling at N GPUs (N = number of visible GPUs): seam check, bench with the P2P exchange and with NCCL
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --check > gpurun_out/r2_check_${N}gpu.json 2> gpurun_out/r2_check_${N}gpu.err
echo "check rc $?"; tail -1 gpurun_out/r2_check_${N}gpu.json | cut -c1-120
for ex in p2p nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 50 --warmup 5 --exchange $ex > gpurun_out/r2_scale_${N}gpu_$ex.json 2> gpurun_out/r2_scale_${N}gpu_$ex.err
  echo "bench $ex rc $?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_scale_${N}gpu_$ex.json") if l.startswith("{")][-1])
    print("$ex N=$N", d["value"], d["ms_per_step"], "sustained", d["sustained"]["ms_per_step"], {k[:12]: (v["kernel_ms"], v["gap_before_ms"]) for k, v in d["breakdown"].items()}, "e2e", d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("parse failed", e)
PY
done
