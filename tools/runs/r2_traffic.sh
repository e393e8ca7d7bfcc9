#!/bin/bash
# DRAM traffic of the final tuned kernels (a few ncu metrics, one launch each)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for K in diffusion fastwaves; do
  N=1; [ $K = fastwaves ] && N=2
  timeout 200 ncu --metrics $M --clock-control none -k regex:absinthe_ -s $N -c $N --csv --log-file gpurun_out/r2_traffic_$K.csv python tools/profile_candidate.py $K '{}' 3 > gpurun_out/r2_traffic_$K.log 2>&1
  echo "$K ncu exit $?"
done
