#!/bin/bash
# ncu full-set capture of the tuned diffusion and fast-waves kernels (one ncu tool, two kernels)
mkdir -p gpurun_out
for K in ${KERNELS:-diffusion fastwaves}; do
CMD="python tools/autotune.py --kernel $K --from-tuned --runs 2"
$CMD > gpurun_out/ncu_plain_$K.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:absinthe_group -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_full_$K.log 2>&1
echo "$K exit $?"
done
