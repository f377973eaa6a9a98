#!/bin/bash
# Round-2 call 38 (4 GPUs): config-5 NGCF on 4 GPUs, symmetric-memory push exchange vs NCCL send/recv (the default switches between them
# by world size: 2 and 8 GPUs are measured, this is the missing point).
set -u
O=gpurun_out; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 4 --only-c5 > $O/r02_c5n4_$name.json 2> $O/r02_c5n4_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n4_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','value','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','exchange','loss_mean')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n4_$name.err').read()[-2500:])
P2
}
run symm YR_SHARD_EXCHANGE=symm
run p2p YR_SHARD_EXCHANGE=p2p
