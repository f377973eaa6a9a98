#!/bin/bash
# Round-2 call 13 (8 GPUs): config 5 with and without the column-panel schedule.
set -u
O=gpurun_out
mkdir -p $O
for CP in 1 0; do export YR_C5_SKIP_MF=1
YR_SHARD_COLPANELS=$CP timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2964$CP bench.py --gpus 8 --only-c5 > $O/r02_c5_n8_cp$CP.json 2> $O/r02_c5_n8_cp$CP.err; echo "bench rc=$?"
python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5_n8_cp$CP.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k,v in d['extra'].items():
        if k.startswith('c5'): print($CP, k, {kk:v.get(kk) for kk in ('ms_per_step','value','efficiency_vs_n1','spmm_ms_per_layer','exchange_ms_per_layer_alone')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5_n8_cp$CP.err').read()[-3000:])
P2
done
