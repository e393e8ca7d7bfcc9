#!/bin/bash
# Cluster column kernels: parity on the zoo, then interleaved A/B at 1024x1024x80 (bit-identity against the tuned two-group variant)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "column-cluster" > gpurun_out/r2_cluster_parity.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_cluster_parity.log; tail -3 gpurun_out/r2_cluster_parity.log
source tools/runs/r2_cluster_cands.sh
timeout 900 python tools/ab_test.py --check fastwaves "${CANDS[@]}" > gpurun_out/r2_ab_cluster.txt 2>&1
echo "ab rc $?"; grep -v "^loading" gpurun_out/r2_ab_cluster.txt | sed 's/"_clean.*"COLUMN"//' | tail -30
bash tools/runs/r2_cluster2.sh
