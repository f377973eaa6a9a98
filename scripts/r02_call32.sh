#!/bin/bash
# Round-2 call 32 (2 GPUs): config-5 NGCF with SMs reserved for the exchange's NCCL kernels (YR_SHARD_RESERVE_SMS), with and
# without column panels, against compute only (YR_SHARD_EXCHANGE=none) — does the exchange now run underneath the SpMM?
set -u
O=gpurun_out; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29680 bench.py --gpus 2 --only-c5 > $O/r02_c5n2_$name.json 2> $O/r02_c5n2_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n2_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','loss_mean')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n2_$name.err').read()[-2500:])
P2
}
run res16 YR_SHARD_RESERVE_SMS=16
run res16_col YR_SHARD_RESERVE_SMS=16 YR_SHARD_INTERLEAVE=1
run res16_none YR_SHARD_RESERVE_SMS=16 YR_SHARD_EXCHANGE=none
run res8_col YR_SHARD_RESERVE_SMS=8 YR_SHARD_INTERLEAVE=1
timeout -s KILL 600 env YR_SHARD_RESERVE_SMS=16 python -m pytest tests/test_gpu_shard.py -x -q -m gpu -k "nccl" > $O/r02_tests20.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests20.log
