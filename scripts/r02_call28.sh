#!/bin/bash
# Round-2 call 28 (2 GPUs): shard tests (world 1 + NCCL world 2) with the ordered NGCF tail and the tensor-core dense kernels; the whole bench
# line on 2 GPUs as the driver launches it; config 5 on 1 GPU (the new efficiency base).
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_ngcf.py -x -m gpu -q > $O/r02_tests17.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests17.log
CUDA_VISIBLE_DEVICES=0 timeout -s KILL 600 python bench.py --only-c5 > $O/r02_c5_n1_tc2.json 2> $O/r02_c5_n1_tc2.err; echo "n1 rc=$?"
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2b.json 2> $O/r02_bench_n2b.err; echo "n2 rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r02_c5_n1_tc2.json','gpurun_out/r02_bench_n2b.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')})
        for k,v in d['extra'].items():
            if k.startswith('c5'): print(d['n_gpus'], k, {kk:v.get(kk) for kk in ('ms_per_step','value','efficiency_vs_n1','spmm_ms_per_layer','dense_fwd_ms_per_layer','dense_bwd_ms_per_layer','exchange_ms_per_layer_alone','error')})
    except Exception as e:
        print("parse failed", f, e); print(open(f.replace('.json','.err')).read()[-2000:])
P
