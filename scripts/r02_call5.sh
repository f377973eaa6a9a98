#!/bin/bash
# Round-2 call 5 (1 GPU): new tests, the full bench line with config 5 at its stated scale, ncu launch list + full captures.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_builders.py tests/test_gpu_eval.py tests/test_gpu_mf.py -m gpu -q > $O/r02_tests5.log 2>&1; echo "tests rc=$?"; tail -6 $O/r02_tests5.log
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','long_run','cpu_baseline')})
    print(json.dumps(d['roofline'])[:1500])
    for k,v in d['extra'].items():
        print(k, json.dumps(v)[:1200])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n1.err').read()[-3000:])
P
NG="python bench.py --only ngcf --steps 2 --warmup 1"
$NG > $O/r02_plain_ngcf.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_ngcf.csv $NG > $O/r02_ncu_ngcf.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm_chunk|dense_fwd|dense_bwd|ngcf_tail|dense_opt" -s 22 -c 14 -o $O/r02_prof_ngcf $NG > $O/r02_ncu_ngcf_full.log 2>&1; echo "ncu full rc=$?"
MF="python bench.py --only mf --steps 50 --warmup 3"
YR_BENCH_MF_OPT=adam $MF > $O/r02_plain_mf.log 2>&1 && \
YR_BENCH_MF_OPT=adam timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bpr_mf_train|mf_sort" -s 2 -c 2 -o $O/r02_prof_mf_adam $MF > $O/r02_ncu_mf_full.log 2>&1; echo "ncu mf rc=$?"
ls -la $O | grep r02_prof
