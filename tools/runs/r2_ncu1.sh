#!/bin/bash
# ncu full-set capture of one column candidate: plain run first, then the profile (third launch of the kernel)
mkdir -p gpurun_out
. tools/runs/r2_cands1.sh
for name in H A; do
  eval "cand=\$$name"
  timeout 300 python tools/profile_candidate.py fastwaves "$cand" > gpurun_out/r2_prof_$name.plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_prof_$name.plain.log; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:absinthe_fastwaves -s 2 -c 1 -f -o gpurun_out/r2_prof_col_$name python tools/profile_candidate.py fastwaves "$cand" > gpurun_out/r2_prof_$name.ncu.log 2>&1
  echo "$name ncu rc $?"
done
ls -la gpurun_out/*.ncu-rep
