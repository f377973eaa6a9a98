#!/bin/bash
# Round-2 call 39 (1 GPU): last sanity of the final tree: sliced-evaluation parity, ABI, smoke.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 400 python -m pytest tests/test_gpu_eval.py tests/test_cabi.py -x -q -m gpu -k "sliced or cabi or symbols or dataframe" > $O/r02_tests24.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests24.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r02_smoke2.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02_smoke2.log
