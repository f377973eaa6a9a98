#!/bin/bash
# Round-2 call 37 (1 GPU): final build: whole GPU suite, smoke(), the full bench line (config 5 at scale) and the reference arm, the launch
# list of one NGCF step (ncu --metrics gpu__time_duration.sum) — all for profiles/.
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x > $O/r02_tests23.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests23.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02_smoke.log
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1d.json 2> $O/r02_bench_n1d.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_n1d.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','long_run','clocks','gpu_launches')})
    print(json.dumps(d['roofline'])[:900])
    for k,v in d['extra'].items():
        if isinstance(v, dict): print(k, {kk:v.get(kk) for kk in ('ms_per_step','ms','value','e2e_value','efficiency_vs_n1','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','item_slices','error') if kk in v})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n1d.err').read()[-3000:])
P
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref2.json 2> $O/r02_bench_ref2.err; echo "ref rc=$?"; tail -c 400 $O/r02_bench_ref2.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_ngcf_step2.csv python bench.py --only ngcf --steps 2 --warmup 1 > $O/ncu_launch.log 2>&1; echo "ncu rc=$?"
