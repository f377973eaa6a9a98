#!/bin/bash
# Round-2 call 29 (8 GPUs): config-5 NGCF with the tensor-core dense kernels: row panels (default) vs interleaved panels + column-panel SpMM.
set -u
O=gpurun_out; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus 8 --only-c5 > $O/r02_c5n8_$name.json 2> $O/r02_c5n8_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n8_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','value','efficiency_vs_n1','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','row_panels')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n8_$name.err').read()[-2000:])
P2
}
run rowpanels YR_SHARD_INTERLEAVE=0
run colpanels YR_SHARD_INTERLEAVE=1
