#!/bin/bash
# the reference's "Optimization" recipe on the B200 (ref README.md:86-90): fastwaves.py -e -g was run in the
# development container (MILPs), here: run.sh -> output.txt -> -p -> report
mkdir -p gpurun_out
ROOT=$PWD
export PYTHONPATH=$ROOT:$PYTHONPATH
K=${K:-fastwaves}; E=${EXPLORE_DIR:-build/explore}; cd $ROOT/$E/$K
( time bash run.sh > output.txt ) > $ROOT/gpurun_out/r2_explore_run.log 2>&1; echo "run.sh rc $?"
cd $ROOT/$E
python $ROOT/$K.py -p output.txt -f ./$K > $ROOT/gpurun_out/r2_explore_parse.log 2>&1; echo "parse rc $?"
python -m absinthe_b200.report ./$K > $ROOT/gpurun_out/r2_explore_report_$K.json 2>&1; echo "report rc $?"
mkdir -p $ROOT/gpurun_out/explore_$K; cp $K/results.csv $K/estimates.csv $K/summary.csv $ROOT/gpurun_out/explore_$K/
tail -6 $ROOT/gpurun_out/r2_explore_report_$K.json
