#!/bin/bash
# final scaling lines: N = visible GPUs, P2P exchange (and one single-GPU run on the same box for the efficiency)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-extras > gpurun_out/r2_final_1gpu_on_${N}gpu_box.json 2> gpurun_out/r2_final_1gpu_on_${N}gpu_box.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --check > gpurun_out/r2_final_check_${N}gpu.json 2> gpurun_out/r2_final_check_${N}gpu.err
echo "check rc $?"; tail -1 gpurun_out/r2_final_check_${N}gpu.json | cut -c1-60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_final_${N}gpu.json 2> gpurun_out/r2_final_${N}gpu.err
echo "bench rc $?"
python - <<PY
import json
for f in ("r2_final_1gpu_on_${N}gpu_box", "r2_final_${N}gpu"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], "sustained", d["sustained"]["ms_per_step"], {k[:12]: (v["kernel_ms"], v["gap_before_ms"]) for k, v in d["breakdown"].items()}, "e2e", d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "parse failed", e)
PY
