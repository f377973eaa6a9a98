"""Profiling aid: full-catalog MF evaluation time for row shards of different sizes (the 1/2/4/8-GPU shard of the
Yelp-shape eval set). SEGS/YR_EVAL_SEGMENTS drove the item-segment experiment recorded in DESIGN.md section 6 (dropped);
the library ignores it now. Usage: python scripts/eval_seg_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from yelprecommendation_b200 import ops
from yelprecommendation_b200.data import synthetic as syn

w = bench.build_workload()
dev = torch.device("cuda", 0)
full = ops.DeviceEvalCSR(w.ecsr, dev, 10)
Up, Vp = syn.planted_embeddings(w.inter)
Ud, Vd = torch.from_numpy(Up).to(dev), torch.from_numpy(Vp).to(dev)
Vt, _ = ops.transpose_items(Vd)
ref = ops.eval_topk_metrics(Ud, Vd, full, Vt, mode="exact")
for shards in (1, 2, 4, 8):
    n = w.ecsr.n_eval // shards
    part = full.slice(0, n) if shards > 1 else full
    for S in (os.environ.get("SEGS", "0")).split(","):
        if S == "0":
            os.environ.pop("YR_EVAL_SEGMENTS", None)
        else:
            os.environ["YR_EVAL_SEGMENTS"] = S
        for _ in range(3):
            out = ops.eval_topk_metrics(Ud, Vd, part, Vt, mode="tc")
        ms = [bench.timed(lambda i: ops.eval_topk_metrics(Ud, Vd, part, Vt, mode="tc"), 1) for _ in range(7)]
        ok = torch.equal(out[0], ref[0][:n]) and torch.equal(out[2], ref[2][:n])
        fb = int(ops.eval_topk_metrics.last_fallback_rows.item())
        print(f"rows {n:6d} (1/{shards})  S={S:>2s}  {np.median(ms):7.3f} ms  {n / (np.median(ms) * 1e-3) / 1e6:6.2f} M users/s  "
              f"bit-identical={ok} fallback={fb}", flush=True)
