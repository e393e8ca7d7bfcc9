#!/bin/bash
mkdir -p gpurun_out
NEW='{"_shapes":[[256,4,40],[64,4,80]],"GROUPS":{"1":{"SUBTILE":[64,4,2]}}}'
NEWB='{"_shapes":[[1024,2,20],[64,4,80]],"GROUPS":{"1":{"SUBTILE":[64,2,4]}}}'
timeout 600 python tools/ab_test.py fastwaves '{}' "$NEW" "$NEWB" '{}' "$NEW" "$NEWB" > gpurun_out/r2_ab_shapes_fw.txt 2>&1
echo "fw rc $?"; grep -v "^loading" gpurun_out/r2_ab_shapes_fw.txt | tail -8
timeout 600 python tools/ab_test.py --check diffusion '{}' '{"_shapes":[[1024,16,1]]}' '{"_shapes":[[256,16,1]]}' '{"_shapes":[[512,32,1]]}' '{"_shapes":[[1024,32,1]]}' '{"_shapes":[[512,16,2]]}' '{"_shapes":[[1024,8,1]]}' '{"_shapes":[[512,8,1]]}' '{"_shapes":[[1024,16,2]]}' '{}' > gpurun_out/r2_ab_shapes_diff.txt 2>&1
echo "diff rc $?"; grep -v "^loading" gpurun_out/r2_ab_shapes_diff.txt | tail -20
