#!/bin/bash
# training sweep on the GPU: generate, build, run, parse, fit (ref README "Training" recipe)
mkdir -p gpurun_out
echo skip > gpurun_out/status.txt
for FAM in fitcache fitddr; do
  rm -rf /tmp/$FAM; mkdir -p /tmp/$FAM
  ( time python $FAM.py -g -f /tmp/$FAM ) > gpurun_out/${FAM}_gen.log 2>&1
  ( cd /tmp/$FAM && time make -j16 ) > gpurun_out/${FAM}_make.log 2>&1; echo "$FAM make exit $?" >> gpurun_out/status.txt
  ( cd /tmp/$FAM && time bash run.sh > output.txt ) > gpurun_out/${FAM}_run.log 2>&1; echo "$FAM run exit $?" >> gpurun_out/status.txt
  python $FAM.py -p output.txt -f /tmp/$FAM > /dev/null
  cp /tmp/$FAM/results.csv gpurun_out/${FAM}_results.csv
done
python -m absinthe_b200.fit --cache gpurun_out/fitcache_results.csv --ddr gpurun_out/fitddr_results.csv --cores 148 > gpurun_out/fit_unconstrained.json 2>&1
python -m absinthe_b200.fit --cache gpurun_out/fitcache_results.csv --ddr gpurun_out/fitddr_results.csv --cores 148 --nonnegative > gpurun_out/fit_nonnegative.json 2>&1
cat gpurun_out/status.txt; tail -n 3 gpurun_out/pytest_training.log; tail -n 4 gpurun_out/fitcache_make.log gpurun_out/fitcache_run.log gpurun_out/fitddr_run.log; cat gpurun_out/fit_unconstrained.json gpurun_out/fit_nonnegative.json; wc -l gpurun_out/*_results.csv
