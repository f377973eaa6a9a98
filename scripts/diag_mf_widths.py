import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from util import cfg, batches_from
from oracle import cport
from yelprecommendation_b200.trainers import MFTrainer
for d in (32, 64, 128):
  for steps in (1, 2, 6):
    rng = np.random.default_rng(d)
    nU, nI, B = 3000, 2500, 1024
    torch.manual_seed(0)
    tr = MFTrainer(cfg(optimizer="sgd", lr=1e-2, weight_decay=1e-3, batch_size=B, embed_size=d), nI, nU)
    U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy(); V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
    u, p, n = rng.integers(0, nU, B * steps), rng.integers(0, nI, B * steps), rng.integers(0, nI, B * steps)
    u[:32] = 5
    b = batches_from(u, p, n, B)
    tr.train(b)
    orc = cport.MFTrainerOracle(U0, V0, "sgd", 1e-2, 1e-3)
    orc.train([{k: v.numpy() for k, v in x.items()} for x in b])
    Ug = tr.model.user_embedding.weight.detach().cpu().numpy()
    err = np.abs(Ug - orc.U); r, c = np.unravel_index(err.argmax(), err.shape)
    touched = np.zeros(nU, bool); touched[u] = True
    print(f"d={d} steps={steps} max err {err.max():.3e} rel {err.max()/np.abs(orc.U).max():.2e} at row {r} col {c} touched={touched[r]} "
          f"n_bad_rows={(err.max(1) > 1e-5*np.abs(orc.U).max()).sum()} bad_untouched={((err.max(1) > 1e-5*np.abs(orc.U).max()) & ~touched).sum()} "
          f"ours {Ug[r,c]:.8f} orc {orc.U[r,c]:.8f} U0 {U0[r,c]:.8f}")

# leftover gradient scratch / flags after a dense-semantics step (must be all zero)
for d in (32, 64):
    rng = np.random.default_rng(d)
    nU, nI, B = 3000, 2500, 1024
    torch.manual_seed(0)
    tr = MFTrainer(cfg(optimizer="sgd", lr=1e-2, weight_decay=1e-3, batch_size=B, embed_size=d), nI, nU)
    U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy(); V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
    u, p, n = rng.integers(0, nU, B), rng.integers(0, nI, B), rng.integers(0, nI, B)
    tr.train(batches_from(u, p, n, B))
    orc = cport.MFTrainerOracle(U0, V0, "sgd", 1e-2, 1e-3)
    orc.train([{k: v.numpy() for k, v in x.items()} for x in batches_from(u, p, n, B)])
    Ug = tr.model.user_embedding.weight.detach().cpu().numpy()
    gfull = (U0 * (1 - 1e-2 * 1e-3) - orc.U) / 1e-2          # the oracle's gradient rows
    gl = tr._scratch["gU"].cpu().numpy()
    bad = np.nonzero(np.abs(gl).max(1) > 0)[0]
    for r in bad[:4]:
        print(f"d={d} row {r}: leftover/full-gradient ratio (first 4 cols) {gl[r,:4] / gfull[r,:4]}, param err / (lr*leftover) {((Ug[r,:4] - orc.U[r,:4]) / (1e-2 * gl[r,:4]))}, cols nonzero {int((gl[r] != 0).sum())}/{d}")
    sc = tr._scratch
    gU = sc["gU"].cpu().numpy(); fU = sc["flagU"].cpu().numpy()
    rows = np.nonzero(np.abs(gU).max(1) > 0)[0]
    cnt = np.bincount(u, minlength=nU)
    print(f"d={d}: leftover gU rows {len(rows)} {rows[:10]} flags set {int(fU.sum())}; multiplicity of those users in the batch {cnt[rows[:10]]}")
    pos_of = {int(r): np.nonzero(u == r)[0].tolist() for r in rows[:12]}
    print(f"d={d}: batch positions of the affected users: {pos_of}")
    touched_rows = np.unique(u)
    print(f"d={d}: touched rows with r%8==0: {int((touched_rows % 8 == 0).sum())}, of those >= 768: {int(((touched_rows % 8 == 0) & (touched_rows >= 768)).sum())}; "
          f"affected rows %8: {np.unique(rows % 8)} min {rows.min() if len(rows) else -1}")
    gV = sc["gV"].cpu().numpy(); rowsV = np.nonzero(np.abs(gV).max(1) > 0)[0]
    print(f"d={d}: leftover gV rows {len(rowsV)} {rowsV[:10]} (nU={nU}: sweep index of item row r is nU + r; (nU + r) % 8 = {np.unique((nU + rowsV) % 8)})")
