#!/bin/bash
# ncu capture of one kernel of the tuned variant of $1: plain run first, then launch list + full set
mkdir -p gpurun_out
CMD="python tools/autotune.py --kernel $1 --from-tuned --runs 2"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${2:-absinthe_} -s 2 -c 1 -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/ncu_full.log; cat gpurun_out/ncu_plain.log | tail -n 2
