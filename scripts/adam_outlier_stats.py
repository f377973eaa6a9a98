"""Distribution of the Adam outlier count of the MF width test (elements beyond 1e-5*max against the C oracle) over
repeated runs. Usage: python scripts/adam_outlier_stats.py [d] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import cfg, batches_from
from oracle import cport
from yelprecommendation_b200.trainers import MFTrainer

d = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(d)
nU, nI, B, steps = 3000, 2500, 1024, 6
u, p, n = rng.integers(0, nU, B * steps), rng.integers(0, nI, B * steps), rng.integers(0, nI, B * steps)
u[:32] = 5
b = batches_from(u, p, n, B)
out = []
for r in range(reps):
    tr = MFTrainer(cfg(optimizer="adam", lr=1e-2, weight_decay=0.0, batch_size=B, embed_size=d), nI, nU)
    U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy()
    V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
    tr.train(b)
    orc = cport.MFTrainerOracle(U0, V0, "adam", 1e-2, 0.0)
    orc.train([{k: v.numpy() for k, v in x.items()} for x in b])
    row = []
    for got, want in ((tr.model.user_embedding.weight, orc.U), (tr.model.item_embedding.weight, orc.V)):
        g = got.detach().cpu().numpy().astype(np.float64)
        e = np.abs(g - want)
        s = np.abs(want).max()
        row += [int((e > 1e-5 * s).sum()), float(e.max() / s), float(np.linalg.norm(g - want) / np.linalg.norm(want))]
    out.append(row)
a = np.array(out)
print(f"d={d} reps={reps}: outliers U mean {a[:,0].mean():.1f} max {a[:,0].max():.0f}; V mean {a[:,3].mean():.1f} max {a[:,3].max():.0f}; "
      f"worst |err|/max U {a[:,1].max():.2e} V {a[:,4].max():.2e}; worst rel_fro U {a[:,2].max():.2e} V {a[:,5].max():.2e}")
print("V counts:", a[:, 3].astype(int).tolist())
print("U counts:", a[:, 0].astype(int).tolist())
