"""Negative sampler (reference data/datasets/mf_dataset.py:18-22).
CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) pin the oracle's stream; distribution properties.
GPU: yr_sample_negatives is bit-exact against the oracle; full-size properties; trainers consume the device loader."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import cport
from yelprecommendation_b200.data import synthetic as syn

KAT = [  # counter, key, expected — Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _problem(nu=400, ni=300, nnz=9000, seed=5):
    inter = syn.make_interactions(num_users=nu, num_items=ni, nnz=nnz, seed=seed, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    ptr = split.train_ptr.astype(np.int64)
    rows = np.repeat(np.arange(nu), np.diff(ptr))
    order = np.lexsort((split.train_items, rows))
    return inter, split, ptr, split.train_items[order].astype(np.int32), rows.astype(np.int64)


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert tuple(int(x) for x in cport.philox4x32_10(ctr, key)) == want


def test_oracle_sampler_properties():
    inter, split, ptr, idx, users = _problem()
    neg, failed = cport.sample_negatives(users, ptr, idx, inter.num_items, seed=11)
    assert failed == 0 and neg.min() >= 0 and neg.max() < inter.num_items
    pos_keys = set((users * inter.num_items + idx).tolist())
    assert not any(int(k) in pos_keys for k in users * inter.num_items + neg)
    # a function of (seed, global index): a shard [a, b) with offset a reproduces the slice
    part, _ = cport.sample_negatives(users[1000:3000], ptr, idx, inter.num_items, seed=11, offset=1000)
    assert np.array_equal(part, neg[1000:3000])
    assert not np.array_equal(cport.sample_negatives(users, ptr, idx, inter.num_items, seed=12)[0], neg)
    # uniform over the complement: one user with 3 of 6 items positive, 60k draws -> each allowed item 1/3 +- 4 sigma
    n = 60000
    neg, _ = cport.sample_negatives(np.zeros(n, np.int64), np.array([0, 3]), np.array([1, 2, 4]), 6, seed=3)
    cnt = np.bincount(neg, minlength=6)
    assert cnt[[1, 2, 4]].sum() == 0
    assert np.all(np.abs(cnt[[0, 3, 5]] - n / 3) < 4 * np.sqrt(n * (1 / 3) * (2 / 3)))
    # a user holding every item: reported, not an endless loop
    neg, failed = cport.sample_negatives(np.zeros(4, np.int64), np.array([0, 3]), np.array([0, 1, 2]), 3, seed=3, max_blocks=4)
    assert failed == 4 and np.all(neg == -1)


@pytest.mark.gpu
def test_gpu_sampler_bit_exact_vs_oracle():
    from yelprecommendation_b200 import ops
    inter, split, ptr, idx, users = _problem(nu=3000, ni=1700, nnz=120000)
    dev = torch.device("cuda")
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(dev)
    for seed, off in ((11, 0), (2 ** 40 + 5, 123456789), (2 ** 63 + 1, 2 ** 41)):
        got = ops.sample_negatives(t(users, np.int64), t(ptr, np.int32), t(idx, np.int32), inter.num_users, inter.num_items,
                                   seed, off).cpu().numpy()
        want, failed = cport.sample_negatives(users, ptr, idx, inter.num_items, seed, off)
        assert failed == 0 and np.array_equal(got, want)
    with pytest.raises(IndexError):
        ops.sample_negatives(torch.tensor([inter.num_users], device=dev), t(ptr, np.int32), t(idx, np.int32), inter.num_users,
                             inter.num_items, 1)
    with pytest.raises(RuntimeError):
        ops.sample_negatives(torch.zeros(4, dtype=torch.int64, device=dev), t(np.array([0, 3]), np.int32),
                             t(np.array([0, 1, 2]), np.int32), 1, 3, 1, max_blocks=4)
    assert ops.sample_negatives(torch.zeros(0, dtype=torch.int64, device=dev), t(ptr, np.int32), t(idx, np.int32),
                                inter.num_users, inter.num_items, 1).numel() == 0


@pytest.mark.gpu
def test_gpu_sampler_full_size_properties():
    """Yelp2018 shape (937k train triples): no negative is a train positive, item marginal ~ uniform, epochs differ."""
    from yelprecommendation_b200.data.sampler import DeviceTripleLoader
    inter = syn.make_interactions()
    split = syn.split_per_user(inter, seed=42)
    ld = DeviceTripleLoader.from_split(split, inter.num_items, batch_size=2048, seed=42)
    u0, p0, n0 = (x.cpu().numpy() for x in ld.epoch_triples())
    u1, p1, n1 = (x.cpu().numpy() for x in ld.epoch_triples())
    nI = inter.num_items
    rows = np.repeat(np.arange(inter.num_users), np.diff(split.train_ptr))
    pos_keys = np.unique(rows.astype(np.int64) * nI + split.train_items)
    for u, p, n in ((u0, p0, n0), (u1, p1, n1)):
        assert n.min() >= 0 and n.max() < nI
        assert not np.isin(u * nI + n, pos_keys).any()                    # rejection property
        assert np.array_equal(np.sort(u * nI + p), np.sort(pos_keys))     # every interaction exactly once
    assert not np.array_equal(n0, n1) and not np.array_equal(u0, u1)      # fresh negatives and order every epoch
    cnt = np.bincount(n0, minlength=nI).astype(np.float64)
    assert abs(cnt.mean() - len(n0) / nI) < 1e-9 and cnt.std() < 2.0 * np.sqrt(len(n0) / nI)
    assert len(ld) == (len(u0) + 2047) // 2048 and sum(b["user_id"].numel() for b in ld) == len(u0)


@pytest.mark.gpu
def test_trainers_consume_device_loader():
    """MFTrainer.train(DeviceTripleLoader) == MFTrainer.train(the same triples as host batches) to 1e-6 (SGD)."""
    from yelprecommendation_b200.data.sampler import DeviceTripleLoader
    from yelprecommendation_b200.trainers import MFTrainer
    import tempfile
    inter, split, ptr, idx, users = _problem(nu=1500, ni=900, nnz=50000)
    mk = lambda: SimpleNamespace(device="cuda", model_dir=tempfile.mkdtemp(), embed_size=64, optimizer="sgd", lr=0.05,
                                 weight_decay=0.0, top_n=10, wandb=False, epochs=1, patience=1, best_metric="loss", batch_size=512)
    ld = DeviceTripleLoader.from_split(split, inter.num_items, batch_size=512, seed=7)
    torch.manual_seed(0)
    a = MFTrainer(mk(), inter.num_items, inter.num_users)
    torch.manual_seed(0)
    b = MFTrainer(mk(), inter.num_items, inter.num_users)
    la = a.train(ld)
    ld.set_epoch(0)
    u, p, n = (x.cpu() for x in ld.epoch_triples())
    lb = b.train(syn.to_batches(u.numpy(), p.numpy(), n.numpy(), 512))
    # same triples, same kernel; only the order in which duplicate rows of a batch are summed may differ (atomics)
    assert abs(la - lb) <= 1e-6 * abs(lb)
    for x, y in ((a.model.user_embedding.weight.data, b.model.user_embedding.weight.data),
                 (a.model.item_embedding.weight.data, b.model.item_embedding.weight.data)):
        assert float((x - y).norm() / y.norm()) < 1e-6
