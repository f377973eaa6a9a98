#!/bin/bash
# Round-2 call 19 (1 GPU): ring forward with split accumulators + transposed epilogue: accuracy, time, parts disabled; NGCF tests.
set -u
O=gpurun_out; mkdir -p $O
for dbg in 0 0 1 8 9 15; do echo "dbg=$dbg"; YR_FWD_DBG=$dbg timeout 120 python scripts/dense_bench.py fwd 2>&1 | grep "mode=1" | grep -v "n=4099"; done > $O/r02_dense_fwd_dbg2.txt 2>&1
cat $O/r02_dense_fwd_dbg2.txt
timeout 900 python -m pytest tests/test_gpu_ngcf.py tests/test_gpu_shard.py -x -q -m gpu > $O/r02_tests12.log 2>&1; echo "tests rc=$?"; tail -5 $O/r02_tests12.log
