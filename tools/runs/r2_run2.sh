#!/bin/bash
# round 2, GPU call 2: column parity after the lane-halo fix, diff locations of the fused-halo failures, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "column" > gpurun_out/r2_col_parity2.log 2>&1
echo "parity rc $?" >> gpurun_out/r2_col_parity2.log
timeout 300 python tools/debug_diff.py fastwaves-hand '{"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2), "PLANES": 2}, "FUSE_HALO": True}' > gpurun_out/r2_diff1.log 2>&1
timeout 300 python tools/debug_diff.py fastwaves-hand '{"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2), "PLANES": 2}, "FUSE_HALO": True, "GROUPS": {"2": {"STREAM_K": False}}}' > gpurun_out/r2_diff2.log 2>&1
timeout 300 python tools/debug_diff.py fastwaves-hand '{"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2), "PLANES": 2}, "FUSE_HALO": True, "GROUPS": {"1": {"STREAM_K": False}}}' > gpurun_out/r2_diff3.log 2>&1
. tools/runs/r2_cands1.sh
timeout 600 python -X faulthandler tools/ab_test.py --check fastwaves "$BASE" "$A" "$B" "$C" "$D" "$E" "$F" "$G" "$H" > gpurun_out/r2_col_ab1.log 2>&1
echo "ab rc $?" >> gpurun_out/r2_col_ab1.log
tail -3 gpurun_out/r2_col_parity2.log; cat gpurun_out/r2_diff*.log; tail -30 gpurun_out/r2_col_ab1.log
