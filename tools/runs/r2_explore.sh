#!/bin/bash
# the reference's "Optimization" recipe on the B200 (ref README.md:86-90): fastwaves.py -e -g was run in the
# development container (MILPs), here: run.sh -> output.txt -> -p -> report
mkdir -p gpurun_out
ROOT=$PWD
export PYTHONPATH=$ROOT:$PYTHONPATH
cd $ROOT/build/explore/fastwaves
( time bash run.sh > output.txt ) > $ROOT/gpurun_out/r2_explore_run.log 2>&1; echo "run.sh rc $?"
cd $ROOT/build/explore
python $ROOT/fastwaves.py -p output.txt -f ./fastwaves > $ROOT/gpurun_out/r2_explore_parse.log 2>&1; echo "parse rc $?"
python -m absinthe_b200.report ./fastwaves > $ROOT/gpurun_out/r2_explore_report.json 2>&1; echo "report rc $?"
cp fastwaves/results.csv $ROOT/gpurun_out/r2_explore_results.csv; cp fastwaves/estimates.csv $ROOT/gpurun_out/r2_explore_estimates.csv; cp fastwaves/summary.csv $ROOT/gpurun_out/r2_explore_summary.csv
tail -30 $ROOT/gpurun_out/r2_explore_report.json
