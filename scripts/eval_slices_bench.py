"""Full-catalog evaluation of a row shard (Yelp shape, planted MF tables and the d_eff = 256 NGCF form) per number of item slices:
what one rank of an N-GPU run evaluates (n_eval = 31,668 / N rows), timed with CUDA events."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yelprecommendation_b200 import ops
from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.data.graph import build_eval_csr
inter = syn.make_interactions()
split = syn.split_per_user(inter, seed=42)
uid, pos, mask = syn.eval_lists(split, "valid")
U, V = syn.planted_embeddings(inter)
rng = np.random.default_rng(2)
U4 = np.concatenate([U] + [(U * s_ + 0.05 * rng.standard_normal(U.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
V4 = np.concatenate([V] + [(V * s_ + 0.05 * rng.standard_normal(V.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
dev = torch.device("cuda")
for name, Ue, Ve in (("mf d=64", U, V), ("ngcf d_eff=256", U4, V4)):
    Ud, Vd = torch.from_numpy(Ue).to(dev), torch.from_numpy(Ve).to(dev)
    Vt, _ = ops.transpose_items(Vd)
    for world in (1, 2, 4, 8):
        n = len(uid) // world
        csr = build_eval_csr(uid[:n], pos[:n], mask[:n], inter.num_items)
        ecsr = ops.DeviceEvalCSR(csr, dev, 10)
        for S in sorted({1, ops.eval_item_slices(n, inter.num_items, 10, dev), 2, 4, 6} if world > 1 else {1, 2}):
            fn = lambda: ops.eval_topk_metrics(Ud, Vd, ecsr, Vt=Vt, mode="tc", slices=S)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"{name} world={world} rows={n} slices={S}: {ms:.3f} ms  -> {n * world / ms / 1e3:.1f} M users/s aggregate", flush=True)
