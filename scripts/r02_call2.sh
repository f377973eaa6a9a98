#!/bin/bash
# Round-2 call 2: ordered (atomics-free) BPR-MF trainer parity + SpMM variant sweep.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_mf.py tests/test_cabi.py -m gpu -q > $O/r02_tests_mf.log 2>&1; echo "mf tests rc=$?"; tail -15 $O/r02_tests_mf.log
timeout 600 python scripts/spmm_bench.py 30 -1,1,2,3,4,5,6,7,8,9 0,1 > $O/r02_spmm_sweep.txt 2>&1; echo "sweep rc=$?"; grep "d=64 acc=0\|d=128 acc=0" $O/r02_spmm_sweep.txt
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_mf.py > $O/r02_tests2.log 2>&1; echo "tests rc=$?"; tail -6 $O/r02_tests2.log
timeout 300 python bench.py --only mf --steps 50 --warmup 5 > $O/r02_bench_mf.json 2> $O/r02_bench_mf.err; echo "bench mf rc=$?"; tail -c 1500 $O/r02_bench_mf.json
