#!/bin/bash
# Round-2 call 21 (1 GPU): ring tensor-core backward (d = 64 / 128) first run + forward with the bulk L2 prefetch; NGCF tests.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 180 python scripts/dense_bench.py both > $O/r02_dense_both.txt 2>&1; echo "dense rc=$?"
cat $O/r02_dense_both.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_ngcf.py -x -q -m gpu > $O/r02_tests13.log 2>&1; echo "tests rc=$?"; tail -15 $O/r02_tests13.log
