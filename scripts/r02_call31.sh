#!/bin/bash
# Round-2 call 31 (1 GPU): item-sliced evaluation: parity test, time per slice count for the shard of a 1 / 2 / 4 / 8-GPU run; eval tests.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_gpu_eval.py tests/test_cabi.py -x -q -m gpu > $O/r02_tests19.log 2>&1; echo "tests rc=$?"; tail -5 $O/r02_tests19.log
timeout -s KILL 600 python scripts/eval_slices_bench.py > $O/r02_eval_slices.txt 2>&1; echo "bench rc=$?"; grep "slices=" $O/r02_eval_slices.txt
