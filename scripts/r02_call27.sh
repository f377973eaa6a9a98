#!/bin/bash
# Round-2 call 27 (1 GPU): sharded NGCF with the ordered (atomics-free) tail accumulate: Adam drift per dense mode twice (run-to-run), shard tests.
set -u
O=gpurun_out; mkdir -p $O
for i in 1 2; do timeout -s KILL 300 python scripts/adam_dense_modes.py 2>&1 | grep "mode="; echo; done > $O/r02_adam_modes_sorted.txt 2>&1; cat $O/r02_adam_modes_sorted.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_mf.py -x -q -m gpu > $O/r02_tests16.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests16.log
