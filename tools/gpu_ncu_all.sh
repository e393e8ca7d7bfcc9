#!/bin/bash
# ncu --set full (with source counters) of the tuned bench kernels: diffusion g1, fastwaves g1 + g2
mkdir -p gpurun_out
for K in diffusion fastwaves; do
  CMD="python tools/autotune.py --kernel $K --from-tuned --runs 2"
  $CMD > gpurun_out/ncu_plain_$K.log 2>&1 || { echo "plain run of $K failed"; tail -n 5 gpurun_out/ncu_plain_$K.log; exit 1; }
done
N=1
for K in diffusion fastwaves; do
  CMD="python tools/autotune.py --kernel $K --from-tuned --runs 2"
  [ $K = fastwaves ] && N=2
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:absinthe_ -s $((2 * N)) -c $N -f -o gpurun_out/prof_full_$K $CMD > gpurun_out/ncu_full_$K.log 2>&1
  echo "$K ncu exit $?"; tail -n 2 gpurun_out/ncu_full_$K.log
done
ls -la gpurun_out/*.ncu-rep
