#!/bin/bash
# training sweeps on the B200 at stride 1 (every tile runs: throughput, not launch latency), then counters of a sample
mkdir -p gpurun_out
ROOT=$PWD
export PYTHONPATH=$ROOT:$PYTHONPATH
for FAM in fitcache fitddr; do
  cd $ROOT/build/training/$FAM
  ( time timeout 1500 python -m absinthe_b200.runner --runs 10 $(ls *.desc | sed 's/\.desc$//' | sed 's/^/.\//') > output.txt ) > $ROOT/gpurun_out/r2_${FAM}_run.log 2>&1
  echo "$FAM run rc $?"; tail -3 $ROOT/gpurun_out/r2_${FAM}_run.log
  python - <<PY
import sys
sys.path.insert(0, "$ROOT")
from absinthe_b200.generator import parse_results, write_results
write_results(parse_results("output.txt", 2), "$ROOT/gpurun_out/r2_${FAM}_results_s1.csv")
PY
done
cd $ROOT
wc -l gpurun_out/r2_fit*_results_s1.csv
for FAM in fitcache fitddr; do
  cd $ROOT/build/training/$FAM
  SAMPLE=$(ls *.desc | sed 's/\.desc$//' | awk 'NR % 8 == 1' | sed 's/^/.\//')
  timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:absinthe --csv --log-file $ROOT/gpurun_out/r2_${FAM}_counters.csv python -m absinthe_b200.runner --runs 1 --no-flush $SAMPLE > $ROOT/gpurun_out/r2_${FAM}_counters.log 2>&1
  echo "$FAM ncu rc $?"; wc -l $ROOT/gpurun_out/r2_${FAM}_counters.csv
done
