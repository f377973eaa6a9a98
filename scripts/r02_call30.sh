#!/bin/bash
# Round-2 call 30 (2 GPUs): config-5 NGCF, panel exchange on high-priority NCCL streams vs default priority; NCCL world-2 parity test.
set -u
O=gpurun_out; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env YR_C5_SKIP_MF=1 "$@" timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29670 bench.py --gpus 2 --only-c5 > $O/r02_c5n2_$name.json 2> $O/r02_c5n2_$name.err; echo "$name rc=$?"
  python - <<P2
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5n2_$name.json').read().strip().splitlines() if l.startswith('{')][-1])
    v=d['extra']['c5_ngcf']; print("$name", {kk:v.get(kk) for kk in ('ms_per_step','value','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','row_panels')})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r02_c5n2_$name.err').read()[-2000:])
P2
}
run hiprio YR_SHARD_HIPRIO=1
run loprio YR_SHARD_HIPRIO=0
run hiprio_col YR_SHARD_HIPRIO=1 YR_SHARD_INTERLEAVE=1
timeout -s KILL 600 python -m pytest tests/test_gpu_shard.py -x -q -m gpu -k "nccl or reproducible or world1" > $O/r02_tests18.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests18.log
