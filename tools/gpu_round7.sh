#!/bin/bash
mkdir -p gpurun_out
K=diffusion
CMD="python tools/autotune.py --kernel $K --from-tuned --runs 2"
$CMD > gpurun_out/ncu_plain_$K.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:absinthe_group -s 2 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_full_$K.log 2>&1
echo "ncu $K exit $?"
# the drivers' auto-tuning recipe on the reference's default domain (64 x 64 x 60)
for KERNEL in diffusion advection fastwaves; do
  rm -rf /tmp/$KERNEL; mkdir -p /tmp/$KERNEL
  python $KERNEL.py -a -g -f /tmp/$KERNEL > gpurun_out/${KERNEL}_gen.log 2>&1
  ( cd /tmp/$KERNEL && make -j16 > /dev/null 2>&1; echo "$KERNEL make exit $?"; time bash run.sh > output.txt ) > gpurun_out/${KERNEL}_run.log 2>&1
  python $KERNEL.py -a -p output.txt -f /tmp/$KERNEL > /dev/null
  cp /tmp/$KERNEL/auto.csv gpurun_out/${KERNEL}_auto.csv
  wc -l gpurun_out/${KERNEL}_auto.csv; tail -n 4 gpurun_out/${KERNEL}_run.log
done
# advection on the large domain
timeout 900 python tools/autotune.py --kernel advection --assignment 00000000 --tiles 512x8x1,256x16x1,512x16x1,256x8x2,1024x8x1 --budgets 110,220 --threads 256,384 --stages 2,3 \
  --sweep "{\"INLINE\":[\"all\"],\"BLOCK\":[[1,1],[2,1],[4,1]],\"DIRECT\":[\"none\"],\"FUSE_HALO\":[true],\"SHUFFLE\":[false]}" \
  --out gpurun_out/tune_advection.json > gpurun_out/tune_advection.log 2>&1; grep BEST gpurun_out/tune_advection.log
