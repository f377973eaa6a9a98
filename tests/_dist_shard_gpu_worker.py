"""N-GPU NCCL worker for test_gpu_shard: ShardedMFTrainer (C-ABI kernels + NCCL) against the oracle MFPort."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.torch_port import MFPort                                              # noqa: E402
from yelprecommendation_b200.data import synthetic as syn                          # noqa: E402
from yelprecommendation_b200.trainers.sharded_mf_trainer import ShardedMFTrainer   # noqa: E402


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
        tr = ShardedMFTrainer(cfg, inter.num_items, inter.num_users, init=(U0, V0))
        loss = tr.train(batches)
        U, V = tr.gather_tables()
        port = MFPort(U0, V0, optimizer=optname, lr=lr, weight_decay=wd)
        ref_loss, _ = port.train(batches)
        for got, ref in ((U.cpu(), port.user.weight.detach()), (V.cpu(), port.item.weight.detach())):
            rel = float((got - ref).norm() / ref.norm())
            assert rel < 1e-5, (optname, rel)
        assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss), (loss, ref_loss)
    dist.barrier()
    if rank == 0:
        print("DIST_SHARD_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
