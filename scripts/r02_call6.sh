#!/bin/bash
# Round-2 call 6 (2 GPUs): sharded-trainer tests (all-to-all exchange, ordered accumulate, catch-up Adam; NCCL world 2), then
# the whole bench at N = 2 with config 5 at its stated scale.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_shard.py tests/test_gpu_mf.py -m gpu -q > $O/r02_tests6.log 2>&1; echo "tests rc=$?"; tail -8 $O/r02_tests6.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
    for k,v in d['extra'].items():
        if k.startswith('c5') or k.startswith('eval'): print(k, json.dumps(v)[:1100])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n2.err').read()[-3000:])
P
