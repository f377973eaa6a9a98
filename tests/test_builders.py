"""GPU: device-side Laplacian / CSR builder against the host restatement of
data/datasets/ngcf_data_pipeline.py:19-44 (data/graph.py::build_laplacian, itself pinned to the reference's dense
construction in tests/test_host_logic.py). Integer / index outputs and fp32 values are compared bit for bit."""
import numpy as np
import pytest
import torch

from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.data.graph import build_laplacian, build_laplacian_csr_device, laplacian_to_csr

pytestmark = pytest.mark.gpu


def _check(user, item, rating, nU, nI):
    dev = torch.device("cuda")
    want = laplacian_to_csr(build_laplacian(user, item, rating, nU, nI), "cpu")
    got = build_laplacian_csr_device(torch.from_numpy(np.asarray(user, dtype=np.int64)).to(dev),
                                     torch.from_numpy(np.asarray(item, dtype=np.int64)).to(dev),
                                     torch.from_numpy(np.asarray(rating, dtype=np.float32)).to(dev), nU, nI)
    assert got.n == want.n and got.symmetric
    assert np.array_equal(got.fwd.rowptr.cpu().numpy(), want.fwd.rowptr.numpy())
    assert np.array_equal(got.fwd.col.cpu().numpy(), want.fwd.col.numpy())
    a, b = got.fwd.val.cpu().numpy(), want.fwd.val.numpy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), float(np.abs(a - b).max())
    return got, want


def test_small_with_duplicates_zero_means_and_random_order():
    rng = np.random.default_rng(3)
    nU, nI, n = 57, 43, 900
    user, item = rng.integers(0, nU, n), rng.integers(0, nI, n)           # plenty of duplicate (u, i) pairs
    rating = rng.integers(0, 6, n).astype(np.float32)                      # zeros included: zero MEANS are dropped
    # make sure every node keeps a positive degree (the reference yields inf/NaN otherwise)
    user = np.concatenate([user, np.arange(nU), rng.integers(0, nU, nI)])
    item = np.concatenate([item, rng.integers(0, nI, nU), np.arange(nI)])
    rating = np.concatenate([rating, np.full(nU + nI, 3.0, np.float32)])
    _check(user, item, rating, nU, nI)


@pytest.mark.parametrize("stars", [False, True])
def test_yelp_shape_bit_exact(stars):
    inter = syn.make_interactions(star_ratings=stars)
    perm = np.random.default_rng(1).permutation(len(inter.user))
    got, want = _check(inter.user[perm], inter.item[perm], inter.rating[perm], inter.num_users, inter.num_items)
    assert got.fwd.nnz == 2 * 1_561_406
    # the plan built from the device arrays drives the same SpMM
    from yelprecommendation_b200 import ops
    X = torch.randn(got.n, 64, device="cuda")
    host = laplacian_to_csr(build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items), "cuda")
    assert torch.equal(ops.spmm_csr(got.fwd, X), ops.spmm_csr(host.fwd, X))


def test_hub_row_longer_than_shared_memory_sort():
    """One item rated by 40,000 users (> the 16,384-entry shared-memory sort): global-scratch bitonic path."""
    rng = np.random.default_rng(5)
    nU, nI = 40_000, 64
    user = np.concatenate([np.arange(nU), rng.integers(0, nU, 30_000)])
    item = np.concatenate([np.zeros(nU, np.int64), rng.integers(1, nI, 30_000)])
    rating = rng.integers(1, 6, len(user)).astype(np.float32)
    perm = rng.permutation(len(user))
    _check(user[perm], item[perm], rating[perm], nU, nI)


def test_out_of_range_id_raises():
    dev = torch.device("cuda")
    u = torch.tensor([0, 1, 5], device=dev)
    i = torch.tensor([0, 1, 1], device=dev)
    with pytest.raises(IndexError):
        build_laplacian_csr_device(u, i, torch.ones(3, device=dev), 3, 2)


def test_device_split_matches_host_restatement_and_eval_csr():
    """MFDataPipeline.split on the device (yr_split_per_user: MT19937 + numpy's legacy shuffle) == data.synthetic.split_per_user
    (pinned to the reference's sklearn split, tests/test_host_logic.py) list for list, order included; the evaluation CSRs
    built from it on the device == build_eval_csr(eval_lists(...)); and MFTrainer.evaluate accepts them."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_eval_csr
    from yelprecommendation_b200.data.split import eval_csr_device, split_per_user_device
    for seed, (nu, ni, nnz) in ((42, (900, 700, 30_000)), (7, (31_668, 38_048, 1_561_406))):
        inter = syn.make_interactions(num_users=nu, num_items=ni, nnz=nnz, seed=3 if nu < 1000 else 2018,
                                      n_clusters=4 if nu < 1000 else 16)
        host = syn.split_per_user(inter, seed=seed)
        deg = np.bincount(inter.user, minlength=nu)
        ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)).cuda()
        dsp = split_per_user_device(ptr, torch.from_numpy(inter.item.astype(np.int64)).cuda(), seed=seed)
        for name in ("train", "valid", "test"):
            assert np.array_equal(getattr(dsp, f"{name}_ptr").cpu().numpy(), getattr(host, f"{name}_ptr")), name
            assert np.array_equal(getattr(dsp, f"{name}_items").cpu().numpy(), getattr(host, f"{name}_items")), name
        for mode in ("valid", "test"):
            uid, pos, mask = syn.eval_lists(host, mode)
            ref = build_eval_csr(uid, pos, mask, ni)
            got = eval_csr_device(dsp, mode, ni, 10)
            for k in ("eval_uid", "mask_ptr", "mask_idx", "act_ptr", "act_idx", "act_nuniq"):
                assert np.array_equal(getattr(got, k).cpu().numpy()[: getattr(ref, k).size], getattr(ref, k)), (mode, k)
