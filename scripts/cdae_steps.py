"""Profiling aid: a few CDAE train steps (B = 32, Yelp-shape catalog) with device-resident masks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from yelprecommendation_b200.trainers import CDAETrainer

w = bench.build_workload()
dev = torch.device("cuda", 0)
nI, Bc, n_b = w.inter.num_items, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(4)
users = rng.choice(w.inter.num_users, Bc * n_b, replace=False)
xin = np.zeros((Bc * n_b, nI), np.float32); neg = np.zeros_like(xin)
for r, u in enumerate(users):
    items = w.split.train_items[w.split.train_ptr[u]:w.split.train_ptr[u + 1]]
    xin[r, items] = 1.0
    cand = rng.integers(0, nI, size=5 * len(items) + 8)
    neg[r, cand[xin[r, cand] == 0][: 5 * len(items)]] = 1.0
cfg = bench.cfg(hidden_size=64, corruption_level=0.6, hidden_activation="sigmoid", output_activation="sigmoid",
                negative_sampling=True, loss_name="bce", lr=1e-4, optimizer="adam")
torch.manual_seed(42)
tr = CDAETrainer(cfg, nI, w.inter.num_users)
db = [{"user_id": torch.from_numpy(users[s:s + Bc].copy()).to(dev), "input_mask": torch.from_numpy(xin[s:s + Bc]).to(dev),
       "negative_mask": torch.from_numpy(neg[s:s + Bc]).to(dev)} for s in range(0, Bc * n_b, Bc)]
tr.train(db[:2])
ms = bench.timed(lambda i: tr.train(db[2:]), 1)
print(f"cdae: {ms / (n_b - 2):.3f} ms/step", flush=True)
