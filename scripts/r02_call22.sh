#!/bin/bash
# Round-2 call 22 (1 GPU): ncu --set full of the ring backward and forward at d = 128, n = 1.5 M (dense_bench launches 74 / 73).
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:ngcf_dense_bwd_tc_kernel -s 74 -c 1 -o $O/r02_bwd_ring_d128 -f python scripts/dense_bench.py bwd > $O/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:ngcf_dense_fwd_tc_kernel -s 72 -c 1 -o $O/r02_fwd_ring_d128b -f python scripts/dense_bench.py fwd > $O/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
grep -c "==PROF==" $O/ncu_bwd.log $O/ncu_fwd.log
