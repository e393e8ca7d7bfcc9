#!/bin/bash
# re-capture after tuned.json switched fast waves and advection to UNIFORM_WARP; then the bench line
mkdir -p gpurun_out
for K in diffusion fastwaves advection; do
  CMD="python tools/profile_candidate.py $K {} 3"
  $CMD > gpurun_out/r2_ncu_plain_$K.log 2>&1 || { echo "plain run of $K failed"; continue; }
  N=1; [ $K = fastwaves ] && N=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:absinthe_ -s $((2 * N)) -c $N -f -o gpurun_out/r2_prof_$K $CMD > gpurun_out/r2_ncu_full_$K.log 2>&1
  echo "$K ncu exit $?"
done
true
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --sustain 0 > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --sustain 0 > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list exit $?"
