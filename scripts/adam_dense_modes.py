"""Sharded NGCF (world 1) vs the torch-CPU port under Adam at every yr_dense_mode: how far the parameters drift (rel. Frobenius)."""
import os, sys
from types import SimpleNamespace
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_shard import _ngcf_problem
from util import rel_fro
from oracle.torch_port import NGCFPort
from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer
for optname, lr, wd, layers, d in [("adam", 1e-2, 0.0, 3, 128), ("adam", 1e-2, 1e-4, 3, 64), ("sgd", 0.05, 0.0, 3, 128), ("adam", 1e-3, 0.0, 3, 128)]:
    inter, L, batches, init = _ngcf_problem(layers=layers, d=d)
    port = NGCFPort(init["embedding.weight"], [init[f"W1.{l}.weight"] for l in range(layers)],
                    [init[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, optname, lr, wd)
    port.train(batches)
    for mode in (0, 1, 2):
        cfg = SimpleNamespace(embed_size=d, num_orders=layers, optimizer=optname, lr=lr, weight_decay=wd, seed=1, ngcf_dense_mode=mode)
        tr = ShardedNGCFTrainer(cfg, inter.num_items, inter.num_users, L, init=init, n_panels=3)
        tr.train(batches)
        e = rel_fro(tr.gather_embedding().cpu(), port.emb.detach())
        w = max(max(rel_fro(tr.W1[l].cpu(), port.W1[l].detach()), rel_fro(tr.W2[l].cpu(), port.W2[l].detach())) for l in range(layers))
        print(f"{optname} lr={lr} d={d} mode={mode}: emb {e:.2e}  W {w:.2e}", flush=True)
