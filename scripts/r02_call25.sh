#!/bin/bash
# Round-2 call 25 (1 GPU): Adam drift per dense mode (sharded NGCF world 1 vs port); BPR-MF tests + bench with the monotonic grid barrier.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 python scripts/adam_dense_modes.py > $O/r02_adam_modes.txt 2>&1; echo "modes rc=$?"; grep "mode=" $O/r02_adam_modes.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_mf.py tests/test_trainer_loop.py -x -q -m gpu > $O/r02_tests15.log 2>&1; echo "mf tests rc=$?"; tail -3 $O/r02_tests15.log
timeout -s KILL 300 python bench.py --only mf --steps 2000 --warmup 200 > $O/r02_mf_barrier.txt 2>&1; echo "mf bench rc=$?"; tail -5 $O/r02_mf_barrier.txt
