#!/bin/bash
# Round-2 call 36 (8 GPUs): the whole bench line on 8 GPUs as the driver launches it (item-sliced evaluation shards, config 5 with the
# symmetric-memory exchange and the tensor-core dense kernels).
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8b.json 2> $O/r02_bench_n8b.err; echo "n8 rc=$?"
python - <<'P'
import json
f='gpurun_out/r02_bench_n8b.json'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','e2e')})
    for k,v in d['extra'].items():
        if isinstance(v, dict): print(k, {kk:v.get(kk) for kk in ('ms_per_step','ms','value','efficiency_vs_n1','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','exchange','error','slices') if kk in v})
except Exception as e:
    print("parse failed", f, e); print(open(f.replace('.json','.err')).read()[-3000:])
P
