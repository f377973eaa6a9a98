import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import cport
from util import cfg, batches_from
from yelprecommendation_b200.trainers import MFTrainer
rng = np.random.default_rng(1)
nU, nI, B, steps = 31_668, 38_048, 2048, 12
tr = MFTrainer(cfg(optimizer="adam", lr=1e-2, weight_decay=0.0, batch_size=B), nI, nU)
U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy()
V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
u = rng.integers(0, nU, B * steps); u[:64] = 7
p, n = rng.integers(0, nI, B * steps), rng.integers(0, nI, B * steps)
b = batches_from(u, p, n, B)
orc = cport.MFTrainerOracle(U0, V0, "adam", 1e-2, 0.0)
for s, x in enumerate(b):
    tr.train([x]); orc.train([{k: v.numpy() for k, v in x.items()}])
    Ug = tr.model.user_embedding.weight.detach().cpu().numpy(); Vg = tr.model.item_embedding.weight.detach().cpu().numpy()
    mg = tr.optimizer.state["U"][0].cpu().numpy(); vg = tr.optimizer.state["U"][1].cpu().numpy()
    dU = np.abs(Ug - orc.U); dV = np.abs(Vg - orc.V)
    r, k = np.unravel_index(dU.argmax(), dU.shape)
    cnt = int((x["user_id"].numpy() == r).sum())
    print(f"step {s+1}: maxdU {dU.max():.3e} (|U|max {np.abs(orc.U).max():.3f}) at row {r} k {k} touched_now {cnt} "
          f"Ug {Ug[r,k]:.8f} Uo {orc.U[r,k]:.8f} m_g {mg[r,k]:.4e} m_o {orc.mU[r,k]:.4e} v_g {vg[r,k]:.4e} v_o {orc.vU[r,k]:.4e} | maxdV {dV.max():.3e}")
    nbad = int((dU > 1e-6).sum())
    print("   elements with dU>1e-6:", nbad, " rel m err max", float(np.abs(mg-orc.mU).max()/np.abs(orc.mU).max()))
