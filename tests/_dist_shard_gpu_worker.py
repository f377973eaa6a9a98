"""N-GPU NCCL worker for test_gpu_shard: ShardedMFTrainer (C-ABI kernels + NCCL) against the oracle MFPort."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.torch_port import MFPort, NGCFPort                                              # noqa: E402
from yelprecommendation_b200.data import synthetic as syn                          # noqa: E402
from yelprecommendation_b200.data.graph import build_laplacian                     # noqa: E402
from yelprecommendation_b200.trainers.sharded_mf_trainer import ShardedMFTrainer   # noqa: E402
from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer  # noqa: E402


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    rank = dist.get_rank()
    inter = syn.make_interactions(num_users=3001, num_items=1571, nnz=60000, seed=11, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    for d, optname, lr, wd in ((64, "sgd", 0.05, 0.0), (128, "adam", 1e-2, 1e-4), (32, "adamw", 1e-2, 1e-2)):
        U0, V0 = (torch.from_numpy(np.ascontiguousarray(a)) for a in syn.planted_embeddings(inter, d=d, seed=5))
        batches = syn.to_batches(u, p, n, 2047)[:5]
        cfg = SimpleNamespace(embed_size=d, optimizer=optname, lr=lr, weight_decay=wd, seed=1)
        port = MFPort(U0, V0, optimizer=optname, lr=lr, weight_decay=wd)
        ref_loss, _ = port.train(batches)
        got_tables = {}
        for exchange, adam_mode in (("all_to_all", "sparse"), ("all_to_all", "dense"), ("all_reduce", "dense")):
            tr = ShardedMFTrainer(cfg, inter.num_items, inter.num_users, init=(U0, V0), exchange=exchange, adam_mode=adam_mode)
            loss = tr.train(batches)
            U, V = tr.gather_tables()
            got_tables[(exchange, adam_mode)] = (U.clone(), V.clone())
            for got, ref in ((U.cpu(), port.user.weight.detach()), (V.cpu(), port.item.weight.detach())):
                rel = float((got - ref).norm() / ref.norm())
                assert rel < 1e-5, (optname, exchange, adam_mode, rel)
            assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss), (loss, ref_loss)
        a, b = got_tables[("all_to_all", "sparse")], got_tables[("all_to_all", "dense")]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "catch-up Adam must equal the dense sweep bit for bit"
    # ---- row-sharded NGCF (rectangular SpMM blocks, all-gather per layer, all-reduce of dW)
    inter = syn.make_interactions(num_users=1203, num_items=958, nnz=30000, seed=21, n_clusters=4, star_ratings=True)
    split = syn.split_per_user(inter, seed=42)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    batches = syn.to_batches(u, p, n, 1023)[:4]
    N, d, layers = inter.num_users + inter.num_items, 64, 3
    g = torch.Generator().manual_seed(3)
    init = {"embedding.weight": torch.randn(N, d, generator=g) * 0.3}
    for l in range(layers):
        init[f"W1.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
        init[f"W2.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
    for optname, lr, wd in (("sgd", 0.05, 0.0), ("adam", 1e-2, 1e-4)):
        cfg = SimpleNamespace(embed_size=d, num_orders=layers, optimizer=optname, lr=lr, weight_decay=wd, seed=1)
        tr = ShardedNGCFTrainer(cfg, inter.num_items, inter.num_users, L, init=init)
        loss = tr.train(batches)
        E0 = tr.gather_embedding().cpu()
        port = NGCFPort(init["embedding.weight"], [init[f"W1.{l}.weight"] for l in range(layers)],
                        [init[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, optname, lr, wd)
        ref_loss, _ = port.train(batches)
        rel = float((E0 - port.emb.detach()).norm() / port.emb.detach().norm())
        assert rel < 1e-5, (optname, "E", rel)
        for l in range(layers):
            for got, ref in ((tr.W1[l].cpu(), port.W1[l].detach()), (tr.W2[l].cpu(), port.W2[l].detach())):
                rel = float((got - ref).norm() / ref.norm())
                assert rel < 1e-5, (optname, "W", l, rel)
        assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss), (loss, ref_loss)
    # ---- BASELINE config 5 in miniature: device-built ScaledGraph, d = 128 x 3 layers, N ranks vs ONE rank of own code
    from yelprecommendation_b200.data.scaled import make_scaled_graph
    sg = make_scaled_graph(6000, 1500, 150_000, seed=5, device="cuda")
    rng = np.random.default_rng(0)
    B = 8192
    batches = [{"user_id": torch.from_numpy(rng.integers(0, 6000, B)), "pos_item": torch.from_numpy(rng.integers(0, 1500, B)),
                "neg_item": torch.from_numpy(rng.integers(0, 1500, B))} for _ in range(2)]
    g = torch.Generator().manual_seed(4)
    init = {"embedding.weight": torch.randn(7500, 128, generator=g) * 0.3}
    for l in range(3):
        init[f"W1.{l}.weight"] = (torch.rand(128, 128, generator=g) * 2 - 1) / 11
        init[f"W2.{l}.weight"] = (torch.rand(128, 128, generator=g) * 2 - 1) / 11
    # lr = 1e-3: at 1e-2 without weight decay this model is ill-conditioned under Adam (profiles/r02_adam_dense_modes.txt) and the
    # dW partials are summed in a different order on N ranks than on one
    cfg = SimpleNamespace(embed_size=128, num_orders=3, optimizer="adam", lr=1e-3, weight_decay=0.0, seed=1)
    trN = ShardedNGCFTrainer(cfg, 1500, 6000, sg, init=init)                       # all ranks, 4 row panels each
    lossN = trN.train(batches)
    EN = trN.gather_embedding().cpu()
    if rank == 0:
        # the same model on ONE rank (solo=True: world 1, no collectives) for the N-GPU vs 1-GPU parity of own code
        tr1 = ShardedNGCFTrainer(cfg, 1500, 6000, sg, init=init, n_panels=1, solo=True)
        loss1 = tr1.train(batches)
        E1 = tr1.gather_embedding().cpu()
        rel = float((EN - E1).norm() / E1.norm())
        assert rel < 1e-5, ("N-GPU vs 1-GPU", rel)
        assert abs(lossN - loss1) < 1e-5 * abs(loss1), (lossN, loss1)
    dist.barrier()
    if rank == 0:
        print("DIST_SHARD_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
