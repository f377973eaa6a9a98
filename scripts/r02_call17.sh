#!/bin/bash
# Round-2 call 17 (1 GPU): ring-based tensor-core forward (d = 64 / 128) vs v1 and the FP32-pipe kernels: accuracy + time.
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/dense_bench.py fwd > $O/r02_dense_fwd_ring.txt 2>&1; echo "ring rc=$?"
YR_FWD_TC_V1=1 timeout 300 python scripts/dense_bench.py fwd > $O/r02_dense_fwd_v1.txt 2>&1; echo "v1 rc=$?"
cat $O/r02_dense_fwd_ring.txt; echo ---; grep "d=64" $O/r02_dense_fwd_v1.txt
