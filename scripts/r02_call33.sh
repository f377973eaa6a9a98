#!/bin/bash
# Round-2 call 33 (2 GPUs): does the panel exchange overlap once NCCL's kernels are launched without CGA clusters (a cluster needs
# several free SMs of ONE GPC) — with and without reserved SMs. Reference points: r02_c5n2_res16_none (compute only, 281.8 ms),
# r02_c5n2_hiprio (default, 285.7 ms).
set -u
O=gpurun_out; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29690 bench.py --gpus 2 --only-c5 > $O/r02_c5n2_$name.json 2> $O/r02_c5n2_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n2_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','spmm_ms_per_layer','exchange_ms_per_layer_alone')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n2_$name.err').read()[-2500:])
P2
}
run cga0 NCCL_CGA_CLUSTER_SIZE=0
run cga0_none NCCL_CGA_CLUSTER_SIZE=0 YR_SHARD_EXCHANGE=none
run cga0_res16 NCCL_CGA_CLUSTER_SIZE=0 YR_SHARD_RESERVE_SMS=16
run cga0_res16_cta8 NCCL_CGA_CLUSTER_SIZE=0 YR_SHARD_RESERVE_SMS=16 NCCL_MAX_CTAS=8
run cga0_res32 NCCL_CGA_CLUSTER_SIZE=0 YR_SHARD_RESERVE_SMS=32
