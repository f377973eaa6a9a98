#!/bin/bash
# Round-2 call 4 (2 GPUs): all GPU tests (incl. the NCCL world-2 workers and the N-GPU vs 1-GPU parity of the sharded NGCF),
# then the whole bench at N = 2 with config 5 at 1/10 scale.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_tests4.log 2>&1; echo "tests rc=$?"; tail -8 $O/r02_tests4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --c5-scale 0.1 > $O/r02_bench_n2_c5s.json 2> $O/r02_bench_n2_c5s.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2_c5s.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')})
    for k,v in d['extra'].items():
        if k.startswith('c5') or k.startswith('eval') or k.startswith('mf_train'): print(k, json.dumps(v)[:900])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_bench_n2_c5s.err').read()[-3000:])
P
