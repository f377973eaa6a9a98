#!/bin/bash
# Round-2 call 11 (2 GPUs): leaner all-to-all plan of the sharded MF trainer: tests (world 1 + NCCL world 2), config 5 at N = 1 and 2.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -q > $O/r02_tests11.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests11.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --only-c5 > $O/r02_c5_n1_lean.json 2> $O/r02_c5_n1_lean.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --only-c5 > $O/r02_c5_n2_lean.json 2> $O/r02_c5_n2_lean.err; echo "n2 rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r02_c5_n1_lean.json','gpurun_out/r02_c5_n2_lean.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in d['extra'].items(): print(d['n_gpus'], k, {kk:v.get(kk) for kk in ('ms_per_step','value','efficiency_vs_n1')})
    except Exception as e:
        print("parse failed", f, e); print(open(f.replace('.json','.err')).read()[-2000:])
P
