"""world_size-2 gloo worker for test_host_logic: shards eval rows like bench.py/parallel.py do on N GPUs and checks
that the all-reduced metric sums equal the single-process result. The per-shard sums come from the oracle here
(CPU box); on GPUs the same code path feeds it the kernel's sums."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cport                                            # noqa: E402
from yelprecommendation_b200 import parallel                         # noqa: E402
from yelprecommendation_b200.data import synthetic as syn            # noqa: E402
from yelprecommendation_b200.data.graph import build_eval_csr        # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    inter = syn.make_interactions(num_users=300, num_items=400, nnz=6000, seed=3, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    uid, pos, mask = syn.eval_lists(split, "valid")
    U, V = syn.planted_embeddings(inter)
    lo, hi = parallel.shard_range(len(uid), rank, world)
    csr = build_eval_csr(uid[lo:hi], pos[lo:hi], mask[lo:hi], inter.num_items)
    _, _, _, sums = cport.eval_topk_metrics(U, V, csr.eval_uid, csr.mask_ptr, csr.mask_idx, csr.act_ptr, csr.act_idx, 10)
    total = parallel.all_reduce_sums(torch.from_numpy(sums))
    full = build_eval_csr(uid, pos, mask, inter.num_items)
    _, _, _, ref = cport.eval_topk_metrics(U, V, full.eval_uid, full.mask_ptr, full.mask_idx, full.act_ptr, full.act_idx, 10)
    assert np.allclose(total.numpy(), ref, rtol=1e-12), (total, ref)
    assert sum(parallel.shard_range(len(uid), r, world)[1] - parallel.shard_range(len(uid), r, world)[0] for r in range(world)) == len(uid)
    dist.barrier()
    if rank == 0:
        print("DIST_EVAL_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
