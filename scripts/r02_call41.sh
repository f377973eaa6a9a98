#!/bin/bash
# Round-2 call 41 (1 GPU): last-layer backward on the FP32-pipe kernel below 16 k listed rows: trainer tests + the NGCF step time.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_ngcf.py -x -q -m gpu -k "train_steps or row_sparse or full_size or yelp" > $O/r02_tests26.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests26.log
timeout -s KILL 200 python bench.py --only ngcf --steps 200 --warmup 20 > $O/r02_ngcf_only.txt 2>&1; echo "bench rc=$?"; tail -3 $O/r02_ngcf_only.txt | cut -c1-400
