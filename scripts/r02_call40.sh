#!/bin/bash
# Round-2 call 40 (1 GPU): sliced-evaluation parity on the final tree.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_eval.py -x -q -m gpu -k "sliced" > $O/r02_tests25.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests25.log
