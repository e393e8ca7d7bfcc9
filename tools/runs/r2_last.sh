#!/bin/bash
mkdir -p gpurun_out
G1='"1":{"SUBTILE":[64,4,2]}'
timeout 150 python tools/ab_test.py --check fastwaves '{}' "{\"GROUPS\":{$G1,\"2\":{\"SUBTILE\":[64,2,6],\"STORE_PTR\":true}}}" '{}' "{\"GROUPS\":{$G1,\"2\":{\"SUBTILE\":[64,2,6],\"STORE_PTR\":true}}}" > gpurun_out/r2_ab_last_fw.txt 2>&1
grep -v "^loading" gpurun_out/r2_ab_last_fw.txt
timeout 150 python tools/ab_test.py --check diffusion '{}' '{"WAIT_SLEEP":64}' '{"FAST_PATH":true}' '{}' '{"WAIT_SLEEP":64}' '{"FAST_PATH":true}' > gpurun_out/r2_ab_last_diff.txt 2>&1
grep -v "^loading" gpurun_out/r2_ab_last_diff.txt
