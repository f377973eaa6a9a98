#!/bin/bash
# Round-2 call 8 (8 GPUs): the whole bench at N = 8 with config 5 at its stated scale, launched the way the driver does.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo.txt 2>&1
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
    for k,v in d['extra'].items():
        if k.startswith('c5') or k.startswith('eval') or k.startswith('mf_train'): print(k, json.dumps(v)[:1100])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n8.err').read()[-4000:])
P
