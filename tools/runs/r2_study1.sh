#!/bin/bash
# model study: MILP estimate (round-1 B200 parameters) vs measured time of 32 groupings per sequence
mkdir -p gpurun_out
TUNED='{"STREAM_K":"auto","INLINE":"all","BLOCK":[2,1],"FUSE_HALO":true,"SHUFFLE":true,"THREADS":256,"WAIT_SLEEP":32,"COLUMN":{"UNROLL":1}}'
for k in fastwaves advection diffusion; do
  timeout 900 python tools/model_study.py measure --variants profiles/r02_study/variants_$k.json --runs 4 --out gpurun_out/r2_model_${k}_generic.json > gpurun_out/r2_model_${k}_generic.log 2>&1
  tail -1 gpurun_out/r2_model_${k}_generic.log
  timeout 900 python tools/model_study.py measure --variants profiles/r02_study/variants_$k.json --runs 4 --cuda "$TUNED" --out gpurun_out/r2_model_${k}_tuned.json > gpurun_out/r2_model_${k}_tuned.log 2>&1
  tail -1 gpurun_out/r2_model_${k}_tuned.log
done
