"""Diagnostic: full-size NGCF SGD steps, fused step (row-sparse / dense last layer, TC / FP32 forward) vs the CPU port."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from util import cfg, rel_fro, rel_err
from oracle import torch_port as tp
from yelprecommendation_b200 import _cabi
from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.data.graph import build_laplacian
from yelprecommendation_b200.trainers import NGCFTrainer
lib = _cabi.load()
inter = syn.make_interactions()
L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
split = syn.split_per_user(inter, seed=42)
tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
batches = syn.to_batches(tu, tpos, tneg, 2048)[:2]
ref = None
for rows, dense in ((1, 1), (0, 1), (1, 0), (0, 0)):
    torch.manual_seed(5)
    tr = NGCFTrainer(cfg(optimizer="sgd", lr=0.05, num_orders=3, batch_size=2048, ngcf_top_rows_mode=rows, ngcf_dense_mode=dense),
                     inter.num_items, inter.num_users, L)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    if ref is None:
        port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(3)],
                           [sd[f"W2.{l}.weight"] for l in range(3)], inter.num_users, L, "sgd", 0.05, 0.0)
        ptotal, psteps = port.train(batches)
        ref = (port, psteps)
    port, psteps = ref
    total = tr.train(batches)
    E = tr.model.embedding.weight.detach().cpu().numpy()
    d_ours, d_port = E - sd["embedding.weight"].numpy(), port.emb.detach().numpy() - sd["embedding.weight"].numpy()
    print(f"rows={rows} tc={dense}: losses {rel_err(tr.last_step_losses.cpu().numpy(), psteps):.2e}  E {rel_fro(E, port.emb.detach().numpy()):.2e}  "
          f"dE {rel_fro(d_ours, d_port):.2e}  |dE| {np.linalg.norm(d_port):.3e}  " +
          " ".join(f"W1.{l} {rel_fro(tr.model.W1[l].weight.detach().cpu().numpy(), port.W1[l].detach().numpy()):.1e} "
                   f"W2.{l} {rel_fro(tr.model.W2[l].weight.detach().cpu().numpy(), port.W2[l].detach().numpy()):.1e}" for l in range(3)), flush=True)
