"""CPU: host-side logic — split/triple/eval-CSR builders against the reference goldens, batch staging rules,
and the user-sharded evaluation reduction under a world_size-2 gloo group."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import lists_from, load_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_matches_reference_pipeline():
    """data/datasets/mf_data_pipeline.py:18-52 (sklearn train_test_split per user) — exact lists and order."""
    from yelprecommendation_b200.data import synthetic as syn
    g = load_npz("mf_small.npz")
    inter = syn.Interactions(int(g["num_users"]), int(g["num_items"]), g["user"], g["item"], g["rating"], None, None)
    split = syn.split_per_user(inter, seed=42)
    for mode, prefix in (("valid", "valid_eval"), ("test", "test_eval")):
        uid, pos, mask = syn.eval_lists(split, mode)
        assert np.array_equal(uid, g[f"{prefix}_uid"])
        assert pos == lists_from(g, prefix, "pos_items")
        assert mask == lists_from(g, prefix, "mask_items")
    tr_user = np.repeat(np.arange(inter.num_users), np.diff(split.train_ptr))
    assert np.array_equal(tr_user, g["train_user"]) and np.array_equal(split.train_items, g["train_item"])
    assert np.array_equal(split.valid_items, g["valid_item"])


def test_presampled_triples_respect_rejection_rule():
    from yelprecommendation_b200.data import synthetic as syn
    inter = syn.make_interactions(num_users=200, num_items=150, nnz=4000, seed=2, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=42)
    assert len(u) == split.train_items.size
    train = {(int(a), int(b)) for a, b in zip(np.repeat(np.arange(200), np.diff(split.train_ptr)), split.train_items)}
    assert all((int(a), int(b)) in train for a, b in zip(u, p))
    assert not any((int(a), int(b)) in train for a, b in zip(u, n))     # mf_dataset.py:18-22
    b = syn.to_batches(u, p, n, 256)
    assert sum(x["user_id"].numel() for x in b) == len(u) and b[0]["user_id"].dtype == torch.int64


def test_synthetic_graph_has_yelp2018_shape_small_and_floors():
    from yelprecommendation_b200.data import synthetic as syn
    inter = syn.make_interactions(num_users=1000, num_items=1500, nnz=30_000, seed=1)
    assert inter.user.size == 30_000
    assert np.unique(inter.user * 1500 + inter.item).size == 30_000
    assert np.bincount(inter.user, minlength=1000).min() >= 10 and np.bincount(inter.item, minlength=1500).min() >= 1


def test_eval_csr_builder_edge_cases():
    from yelprecommendation_b200.data.graph import build_eval_csr
    csr = build_eval_csr([3, 9, 4], [[5, 1, 5], [], [2]], [[7, 2, 2, -1], [], [0]], 10)
    assert csr.mask_idx.tolist() == [2, 7, 9, 0] and csr.mask_ptr.tolist() == [0, 3, 3, 4]   # sorted, unique, wrapped
    assert csr.act_idx.tolist() == [5, 1, 5, 2] and csr.act_nuniq.tolist() == [2, 0, 1]       # order kept
    with pytest.raises(IndexError):
        build_eval_csr([0], [[1]], [[10]], 10)


def test_laplacian_csr_roundtrip_and_transpose():
    from yelprecommendation_b200.data.graph import build_laplacian, laplacian_to_csr
    from yelprecommendation_b200.data import synthetic as syn
    inter = syn.make_interactions(num_users=120, num_items=90, nnz=2000, seed=3, n_clusters=3, star_ratings=True)
    L = build_laplacian(inter.user, inter.item, inter.rating, 120, 90)
    csr = laplacian_to_csr(L, "cpu")
    dense = L.to_dense().numpy()
    rp, ci, va = csr.fwd.rowptr.numpy(), csr.fwd.col.numpy(), csr.fwd.val.numpy()
    rec = np.zeros_like(dense)
    for r in range(210):
        rec[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    assert np.array_equal(rec, dense)
    rpt, cit, vat = csr.bwd.rowptr.numpy(), csr.bwd.col.numpy(), csr.bwd.val.numpy()
    rect = np.zeros_like(dense)
    for r in range(210):
        rect[r, cit[rpt[r]:rpt[r + 1]]] = vat[rpt[r]:rpt[r + 1]]
    assert np.array_equal(rect, dense.T)
    assert np.all(np.isfinite(dense))


def test_user_sharded_metric_reduction_gloo_world2():
    """N>1 evaluation path: rows are sharded contiguously, each rank reduces its 6 metric sums, one all_reduce."""
    script = os.path.join(ROOT, "tests", "_dist_eval_worker.py")
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_EVAL_OK" in r.stdout


def test_row_sharded_mf_choreography_gloo_world2():
    """BASELINE config 5 path: ShardedMFTrainer's gather / all-reduce / sliced grads / all-gather / owner update
    over 2 ranks equals the oracle's single-process MFPort (SGD, Adam + L2, AdamW), ragged batch, OOB id -> IndexError."""
    script = os.path.join(ROOT, "tests", "_dist_shard_worker.py")
    port = 31500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_SHARD_OK" in r.stdout


def test_unsupported_widths_fail_loudly():
    """No silent fallback: a width the kernels do not take raises at construction (INTEGRATION.md lists the widths)."""
    from types import SimpleNamespace
    from yelprecommendation_b200._cabi import YelprecError
    from yelprecommendation_b200.models.cdae import CDAE
    from yelprecommendation_b200.models.ngcf import NGCF
    with pytest.raises(YelprecError):
        NGCF(SimpleNamespace(embed_size=48, num_orders=2), 10, 12)
    NGCF(SimpleNamespace(embed_size=128, num_orders=1), 10, 12)
    base = dict(device="cuda", corruption_level=0.5, hidden_activation="sigmoid", output_activation="sigmoid")
    with pytest.raises(YelprecError):
        CDAE(SimpleNamespace(hidden_size=48, **base), 12, 10)
    with pytest.raises(YelprecError):
        CDAE(SimpleNamespace(hidden_size=64, **{**base, "hidden_activation": "relu"}), 12, 10)
    with pytest.raises(YelprecError):
        CDAE(SimpleNamespace(hidden_size=64, **{**base, "output_activation": "identity"}), 12, 10)
    CDAE(SimpleNamespace(hidden_size=1024, **{**base, "hidden_activation": "identity"}), 12, 10)


def test_shard_layout_positions_are_a_monotone_bijection():
    """data/scaled.py::ShardLayout — rank k owns a block of users AND a block of items; positions in the gathered operand are
    rank-major, every node has exactly one position, users / items map monotonically (so CSR rows keep their column order at
    any world size), and local_rows / to_node_order are inverse to each other."""
    import torch
    from yelprecommendation_b200.data.scaled import ShardLayout
    for nU, nI, world, panels in ((10, 4, 1, 1), (103, 57, 3, 1), (1000, 333, 8, 1), (7, 5, 8, 1), (10, 4, 1, 3), (1000, 333, 8, 4),
                                  (5003, 1201, 2, 8)):
        lay = ShardLayout(nU, nI, world, panels)
        assert lay.per == lay.panels * lay.pp and (panels == 1 or lay.pp % 128 == 0)
        pos = lay.node_pos(torch.arange(nU + nI)).numpy()
        assert len(set(pos.tolist())) == nU + nI and pos.min() >= 0 and pos.max() < world * lay.per
        assert np.all(np.diff(pos[:nU]) > 0) and np.all(np.diff(pos[nU:]) > 0)
        owners = pos // lay.per
        for k in range(world):
            (u0, u1), (i0, i1) = lay.user_block(k), lay.item_block(k)
            assert np.all(owners[u0:u1] == k) and np.all(owners[nU + i0: nU + i1] == k)
        table = torch.arange((nU + nI) * 3, dtype=torch.float32).view(nU + nI, 3)
        gathered = torch.cat([lay.local_rows(k, table) for k in range(world)])
        assert torch.equal(lay.to_node_order(gathered), table)
        if world == 1 and panels == 1:
            assert np.array_equal(pos, np.arange(nU + nI))
        if panels > 1:                              # every panel holds a slice of the users AND a slice of the items
            local = pos % lay.per
            for p_ in range(lay.panels):
                a, b = lay.panel_rows(p_)
                in_p = (local >= a) & (local < b)
                if nU >= 4 * world * panels and nI >= 4 * world * panels:
                    assert in_p[:nU].any() and in_p[nU:].any()


def test_sparse_cdae_batch_equals_the_dense_masks():
    """data/cdae_sparse.py: index lists carry exactly the information of the reference's dense CDAE batch
    (data/datasets/cdae_dataset.py:38-62): active inputs, and the loss positions target + negative_mask != 0 with their targets."""
    import torch
    from yelprecommendation_b200.data.cdae_sparse import sparse_cdae_batch
    rng = np.random.default_rng(0)
    B, nI = 9, 101
    x = (rng.random((B, nI)) < 0.1).astype(np.float32)
    valid = ((rng.random((B, nI)) < 0.05) & (x == 0)).astype(np.float32)
    neg = ((rng.random((B, nI)) < 0.2) & (x == 0) & (valid == 0)).astype(np.float32)
    x[3] = 0                                                               # a user without inputs
    data = {"user_id": torch.arange(B), "input_mask": torch.from_numpy(x), "valid_mask": torch.from_numpy(valid),
            "negative_mask": torch.from_numpy(neg)}
    for extra in (None, "valid_mask"):
        s = sparse_cdae_batch(data, target_extra=extra)
        tgt = x if extra is None else x + valid
        ip, ii, lp, li, lv = (s[k].numpy() for k in ("input_ptr", "input_idx", "loss_ptr", "loss_idx", "loss_val"))
        for b in range(B):
            assert np.array_equal(ii[ip[b]:ip[b + 1]], np.nonzero(x[b])[0])
            want = np.nonzero((tgt[b] + neg[b]) != 0)[0]
            assert np.array_equal(li[lp[b]:lp[b + 1]], want)
            assert np.array_equal(lv[lp[b]:lp[b + 1]], tgt[b][want])


def test_eval_slice_rule_and_sliced_masks():
    """Item-sliced evaluation, host side: the slice-count rule (ops.slices_for) and the per-slice mask CSRs (ops._sliced_masks:
    entries of a slice, re-based, ascending per row) against a numpy restatement; rows that would have fewer than K unmasked
    items inside a slice make the call fall back to the unsliced path."""
    import torch
    from yelprecommendation_b200 import ops
    # Yelp shape, K = 10, 148 SMs: what one rank of a 1 / 2 / 4 / 8-GPU run evaluates
    assert [ops.slices_for(31668 // w, 38048, 10, 148) for w in (1, 2, 4, 8)] == [1, 1, 2, 4]
    assert ops.slices_for(3958, 38048, 16, 148) == 4 and ops.slices_for(3958, 38048, 10, 148, forced=6) == 6
    assert ops.slices_for(100, 3000, 10, 148) == 1 and ops.slices_for(100, 5000, 10, 148) == 2      # slices stay >= 2,048 items
    assert ops.slices_for(0, 38048, 10, 148) == 6                                                   # 64 // K caps S * K
    rng = np.random.default_rng(0)
    n, nI, K, S = 37, 1000, 10, 3
    per = (-(-nI // S) + 127) // 128 * 128                 # 384
    rows = [np.sort(rng.choice(nI, size=rng.integers(0, 60), replace=False)).astype(np.int32) for _ in range(n)]
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    idx = np.concatenate(rows) if ptr[-1] else np.zeros(0, np.int32)
    t = torch.from_numpy
    ecsr = ops.DeviceEvalCSR.from_device(t(np.arange(n, dtype=np.int64)), t(ptr), t(idx), t(np.zeros(n + 1, np.int32)),
                                         t(np.zeros(1, np.int32)), t(np.zeros(n, np.int32)), K)
    masks = ops._sliced_masks(ecsr, S, per, nI)
    assert masks is not None and len(masks) == S
    for s_, (p_s, i_s) in enumerate(masks):
        p_s, i_s = p_s.numpy(), i_s.numpy()
        for r in range(n):
            want = rows[r][(rows[r] >= s_ * per) & (rows[r] < (s_ + 1) * per)] - s_ * per
            assert np.array_equal(i_s[p_s[r]:p_s[r + 1]], want)
    assert ops._sliced_masks(ecsr, S, per, nI) is masks                                   # cached with the evaluation set
    # a row that masks all but K - 1 items of the last slice (232 items: 768 .. 999): not sliceable
    full = np.arange(nI - 232 + (K - 1), nI, dtype=np.int32)
    ptr2 = np.array([0, full.size], np.int32)
    e2 = ops.DeviceEvalCSR.from_device(t(np.zeros(1, np.int64)), t(ptr2), t(full), t(np.zeros(2, np.int32)), t(np.zeros(1, np.int32)),
                                       t(np.zeros(1, np.int32)), K)
    assert ops._sliced_masks(e2, S, per, nI) is None


def test_default_exchange_by_world_size():
    from yelprecommendation_b200.trainers.sharded_ngcf_trainer import _default_exchange
    assert [_default_exchange(w) for w in (1, 2, 3, 4, 5, 8)] == ["p2p", "symm", "symm", "symm", "p2p", "p2p"]
