#!/bin/bash
# DRAM traffic of the cluster column kernels (ncu, a few metrics only)
mkdir -p gpurun_out
source tools/runs/r2_cluster_cands.sh
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,launch__grid_size,launch__cluster_dim_x,launch__cluster_max_active,launch__cluster_max_potential_size,lts__t_bytes.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
for I in ${NCU_LIST:-2 3 4 10}; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:absinthe_ -s 1 -c 1 --csv --log-file gpurun_out/r2_cluster_ncu_$I.csv python tools/profile_candidate.py fastwaves "${CANDS[$I]}" 3 > gpurun_out/r2_cluster_ncu_$I.log 2>&1
  echo "cand $I ncu exit $?"
  grep -v "^==" gpurun_out/r2_cluster_ncu_$I.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print(d.get("Metric Name"), d.get("Metric Value"), d.get("Metric Unit"))
"
done
