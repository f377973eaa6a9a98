#!/bin/bash
# Round-2 call 14 (8 GPUs): config-5 NGCF exchange experiments: compute only / one NCCL all-gather per layer / more NCCL P2P channels.
set -u
O=gpurun_out
mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --only-c5 > $O/r02_c5x_$name.json 2> $O/r02_c5x_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5x_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','efficiency_vs_n1','spmm_ms_per_layer','dense_fwd_ms_per_layer','exchange_ms_per_layer_alone')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5x_$name.err').read()[-2000:])
P2
}
run none YR_SHARD_EXCHANGE=none
run allgather YR_SHARD_EXCHANGE=allgather
run p2p_ch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run p2p_p16 YR_SHARD_PANELS=16
