#!/bin/bash
# Round-2 call 3: FP64 probe, MF bench (register / ordered / Adam), SpMM sweep around the new default, GPU tests, NGCF step,
# config 5 at 1/10 scale on one GPU.
set -u
O=gpurun_out
mkdir -p $O
timeout 60 scripts/bin/fp64_probe > $O/r02_fp64_probe.txt 2>&1; cat $O/r02_fp64_probe.txt
for opt in sgd adam; do
  YR_BENCH_MF_OPT=$opt timeout 200 python bench.py --only mf --steps 200 --warmup 5 2>/dev/null | tail -1
done
YR_BENCH_MF_DET=1 timeout 200 python bench.py --only mf --steps 200 --warmup 5 2>/dev/null | tail -1
timeout 600 python scripts/spmm_bench.py 30 -1,0,1,2,3,4,5,7,8,9 1 > $O/r02_spmm_sweep2.txt 2>&1; echo "sweep rc=$?"; grep "acc=0" $O/r02_spmm_sweep2.txt
timeout 900 python -m pytest tests -m gpu -q > $O/r02_tests3.log 2>&1; echo "tests rc=$?"; tail -12 $O/r02_tests3.log
timeout 300 python bench.py --only ngcf --steps 50 --warmup 5 2>/dev/null | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --c5-scale 0.1 > $O/r02_bench_c5s.json 2> $O/r02_bench_c5s.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c5s.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','long_run')})
    for k,v in d['extra'].items():
        print(k, json.dumps(v)[:600])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_c5s.err').read()[-3000:])
P
