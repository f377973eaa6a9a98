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
