#!/bin/bash
# Round-2 call 10 (1 GPU): whole GPU suite on the current build, then the full bench line (config 5 at scale), saved for profiles/.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_tests10.log 2>&1; echo "tests rc=$?"; tail -6 $O/r02_tests10.log
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1b.json 2> $O/r02_bench_n1b.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n1b.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','long_run')})
    for k,v in d['extra'].items():
        if k.startswith('c5') or k.startswith('cdae') or k.startswith('eval'): print(k, json.dumps(v)[:1000])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n1b.err').read()[-3000:])
P
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "ref rc=$?"; tail -c 600 $O/r02_bench_ref.json
