#!/bin/bash
# Round-2 call 35 (2 GPUs): panel exchange as copy-engine pushes through symmetric memory (YR_SHARD_EXCHANGE=symm): NCCL world-2 parity
# test in that mode, then config-5 NGCF step time against the NCCL send/recv exchange (285.5 ms) and compute only (256.4 ms).
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 600 env YR_SHARD_EXCHANGE=symm python -m pytest tests/test_gpu_shard.py -x -q -m gpu -k "nccl" > $O/r02_tests22.log 2>&1; echo "tests rc=$?"; tail -25 $O/r02_tests22.log | cut -c1-300
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --only-c5 > $O/r02_c5n2_$name.json 2> $O/r02_c5n2_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n2_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','spmm_ms_per_layer','exchange_ms_per_layer_alone','loss_mean')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n2_$name.err').read()[-3000:])
P2
}
run symm YR_SHARD_EXCHANGE=symm
