#!/bin/bash
# round 2 evidence: L2-promotion A/B of the tuned kernels, full-set ncu of the bench kernels, the fused
# column kernel, advection; launch list of bench.py
mkdir -p gpurun_out
for promo in 3 2 1 0; do
  echo "== L2 promotion level $promo" >> gpurun_out/r2_ab_promotion.log
  for k in diffusion fastwaves advection; do
    ABSINTHE_TMA_L2_PROMOTION=$promo timeout 300 python tools/ab_test.py $k '{}' 2>&1 | grep median | sed "s/^/$k /" >> gpurun_out/r2_ab_promotion.log
  done
done
cat gpurun_out/r2_ab_promotion.log
N=1
for K in diffusion fastwaves advection; do
  CMD="python tools/profile_candidate.py $K {} 3"
  $CMD > gpurun_out/r2_ncu_plain_$K.log 2>&1 || { echo "plain run of $K failed"; tail -n 5 gpurun_out/r2_ncu_plain_$K.log; continue; }
  N=1; [ $K = fastwaves ] && N=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:absinthe_ -s $((2 * N)) -c $N -f -o gpurun_out/r2_prof_$K $CMD > gpurun_out/r2_ncu_full_$K.log 2>&1
  echo "$K ncu exit $?"
done
COL='{"_clean":true,"_assignment":[0,0,0,0,0,0,0,0,0],"STREAM_K":"regs","SMEM_BUDGET":232448,"FUSE_HALO":true,"WAIT_SLEEP":32,"_shapes":[[57,9,80]],"BLOCK":[1,1],"COLUMN":{"WARPS":[2,9],"UNROLL":1,"SLOTS":3}}'
python tools/profile_candidate.py fastwaves "$COL" 3 > gpurun_out/r2_ncu_plain_column.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:absinthe_ -s 2 -c 1 -f -o gpurun_out/r2_prof_column python tools/profile_candidate.py fastwaves "$COL" 3 > gpurun_out/r2_ncu_full_column.log 2>&1
echo "column ncu exit $?"
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --sustain 0 > gpurun_out/r2_bench_for_list.json 2> gpurun_out/r2_bench_for_list.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --sustain 0 > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list exit $?"; ls -la gpurun_out/r2_prof_*.ncu-rep
