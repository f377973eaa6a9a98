#!/bin/bash
# Round-2 call 24 (1 GPU): full GPU test suite with the tensor-core backward as the default, dense micro-benchmark, bench line without config 5,
# then config 5 on one GPU.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > $O/r02_tests14.log 2>&1; echo "tests rc=$?"; tail -6 $O/r02_tests14.log
timeout -s KILL 180 python scripts/dense_bench.py both > $O/r02_dense_both2.txt 2>&1; echo "dense rc=$?"; grep "mode=[12]" $O/r02_dense_both2.txt
timeout -s KILL 600 python bench.py --no-c5 --no-cpu-baseline > $O/r02_bench_n1c.json 2> $O/r02_bench_n1c.err; echo "bench rc=$?"
python - <<'P2'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1c.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')})
P2
YR_C5_SKIP_MF=1 timeout -s KILL 600 python bench.py --only-c5 > $O/r02_c5_n1_tc.json 2> $O/r02_c5_n1_tc.err; echo "c5 rc=$?"; tail -c 1500 $O/r02_c5_n1_tc.json
