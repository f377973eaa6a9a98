#!/bin/bash
# Round-2 call 20 (1 GPU): ring forward at d = 128: TMA weight stream vs shared-memory staging (parts disabled).
set -u
O=gpurun_out; mkdir -p $O
for dbg in 9 11 13 25 27 29; do echo "dbg=$dbg"; YR_FWD_DBG=$dbg timeout 120 python scripts/dense_bench.py fwd 2>&1 | grep "mode=1" | grep "n=1500000"; done > $O/r02_dense_fwd_dbg3.txt 2>&1
cat $O/r02_dense_fwd_dbg3.txt
