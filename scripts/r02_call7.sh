#!/bin/bash
# Round-2 call 7 (1 GPU): evaluation kernel with the mask lists staged in shared memory: parity tests, then timing with and
# without the staging (YR_EVAL_MCAP=0).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_ngcf.py tests/test_cdae.py -m gpu -q > $O/r02_tests7.log 2>&1; echo "tests rc=$?"; tail -5 $O/r02_tests7.log
for m in 64 0 32 128; do
  echo "MCAP=$m"; YR_EVAL_MCAP=$m timeout 300 python bench.py --only eval --steps 10 --warmup 3 2>/dev/null | tail -2
done
timeout 300 python scripts/stress_eval.py > $O/r02_stress_eval.log 2>&1; echo "stress rc=$?"; tail -3 $O/r02_stress_eval.log
