#!/bin/bash
# model study with the round-2 parameters: MILP estimate vs measured time of 32 groupings per sequence
mkdir -p gpurun_out
TUNED='{"STREAM_K":"auto","INLINE":"all","BLOCK":[2,1],"FUSE_HALO":true,"SHUFFLE":true,"THREADS":256,"WAIT_SLEEP":32,"COLUMN":{"UNROLL":1}}'
MERGED='{"STREAM_K":"auto","INLINE":"all","BLOCK":[1,1],"FUSE_HALO":true,"SHUFFLE":true,"THREADS":256,"WAIT_SLEEP":32,"SMEM_BUDGET":232448,"COLUMN":{"UNROLL":1,"MERGE":true,"WARPS":[2,9]}}'
for k in fastwaves advection diffusion; do
  timeout 900 python tools/model_study.py measure --variants profiles/r02_study/variants2_$k.json --runs 4 --out gpurun_out/r2b_model_${k}_generic.json > gpurun_out/r2b_model_${k}_generic.log 2>&1
  echo "$k generic: $(tail -1 gpurun_out/r2b_model_${k}_generic.log)"
  timeout 900 python tools/model_study.py measure --variants profiles/r02_study/variants2_$k.json --runs 4 --cuda "$TUNED" --out gpurun_out/r2b_model_${k}_tuned.json > gpurun_out/r2b_model_${k}_tuned.log 2>&1
  echo "$k tuned: $(tail -1 gpurun_out/r2b_model_${k}_tuned.log)"
done
timeout 900 python tools/model_study.py measure --variants profiles/r02_study/variants2_fastwaves.json --runs 4 --cuda "$MERGED" --out gpurun_out/r2b_model_fastwaves_merged.json > gpurun_out/r2b_model_fastwaves_merged.log 2>&1
echo "fastwaves merged: $(tail -1 gpurun_out/r2b_model_fastwaves_merged.log)"
