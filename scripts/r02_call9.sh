#!/bin/bash
# Round-2 call 9 (8 GPUs): hub-row finalize kernel + 8 row panels; config 5 only (bench with --only-c5), twice: panels 8 and 4.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ngcf.py -m gpu -q -k "spmm" > $O/r02_tests9.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests9.log
for P in 8 4; do
YR_SHARD_PANELS=$P timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$P bench.py --gpus 8 --steps 5 --warmup 3 --only-c5 > $O/r02_c5_n8_p$P.json 2> $O/r02_c5_n8_p$P.err; echo "bench rc=$?"
python - <<P2
import json
try:
    d=json.loads(open('gpurun_out/r02_c5_n8_p$P.json').read().strip().splitlines()[-1])
    for k,v in d['extra'].items():
        if k.startswith('c5'): print(k, json.dumps(v)[:900])
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5_n8_p$P.err').read()[-3000:])
P2
done
