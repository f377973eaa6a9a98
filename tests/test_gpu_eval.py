"""GPU parity: fused full-catalog evaluation (score + mask + top-K + metrics) vs the oracle and the reference."""
from math import isclose

import numpy as np
import pytest
import torch

from oracle import cport
from util import RTOL, cfg, lists_from, load_npz, metric_cases, rel_err

pytestmark = pytest.mark.gpu


MODES = ["exact", "tc"]      # FP32-pipe kernel / tensor-core filter + exact re-score: identical outputs required


def tc_ok(d, K):
    from yelprecommendation_b200 import _cabi
    return _cabi.load().yr_eval_tc_supported(d, K) != 0


def run_gpu(U, V, csr, K=10, mode="exact"):
    from yelprecommendation_b200 import ops
    dev = torch.device("cuda")
    ecsr = ops.DeviceEvalCSR(csr, dev, K)
    topk, tsc, um, sums, err = ops.eval_topk_metrics(torch.from_numpy(U).to(dev), torch.from_numpy(V).to(dev), ecsr,
                                                     mode=mode)
    assert int(err.item()) == 0
    return topk.cpu().numpy(), tsc.cpu().numpy(), um.cpu().numpy(), sums.cpu().numpy()


def fallback_rows():
    from yelprecommendation_b200 import ops
    return int(ops.eval_topk_metrics.last_fallback_rows.item())


def check_vs_oracle(U, V, csr, K=10, mode="exact"):
    if mode == "tc" and not tc_ok(U.shape[1], K):
        pytest.skip("tensor-core filter needs d % 32 == 0, d <= 256, K <= 16")
    topk, tsc, um, sums = run_gpu(U, V, csr, K, mode)
    otopk, otsc, oum, osums = cport.eval_topk_metrics(U, V, csr.eval_uid, csr.mask_ptr, csr.mask_idx, csr.act_ptr,
                                                      csr.act_idx, K)
    assert np.array_equal(topk, otopk), "top-K ids must be bit-exact under (score desc, id asc)"
    assert np.array_equal(tsc, otsc), "scores are one fma chain on both sides"
    assert np.array_equal(um, oum), "per-row metric terms use the same double arithmetic"
    assert np.allclose(sums, osums, rtol=1e-12, atol=0)
    return topk, sums


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("prefix", ["valid_eval", "test_eval"])
def test_eval_golden_small(prefix, mode):
    from yelprecommendation_b200.data.graph import build_eval_csr
    from yelprecommendation_b200 import ops
    g = load_npz("mf_small.npz")
    U, V = g["eval_U"].astype(np.float32), g["eval_V"].astype(np.float32)
    uid = g[f"{prefix}_uid"]
    csr = build_eval_csr(uid, lists_from(g, prefix, "pos_items"), lists_from(g, prefix, "mask_items"), int(g["num_items"]))
    topk, sums = check_vs_oracle(U, V, csr, 10, mode)
    ref = g[f"{prefix}_topk"]
    diff = [r for r in range(len(uid)) if not np.array_equal(topk[r], ref[r])]
    for r in diff:   # only fp32-noise ties against the reference's NumPy order (Q5)
        sa, sb = (cport.mf_score(U, V, np.full(10, uid[r]), x[r]) for x in (topk, ref))
        assert np.abs(sa - sb).max() <= 4e-6 * np.abs(sb).max()
    assert len(diff) <= max(1, len(uid) // 50)
    if not diff:
        assert np.allclose(ops.metrics_from_sums(sums, csr.n_eval), g[f"{prefix}_metrics"], rtol=1e-12, atol=0)


def _random_problem(rng, nU, nI, d, n_eval, heavy=False, ties=False):
    from yelprecommendation_b200.data.graph import build_eval_csr
    U = rng.standard_normal((nU, d)).astype(np.float32)
    V = rng.standard_normal((nI, d)).astype(np.float32)
    if ties:
        V[nI // 2:] = V[: nI - nI // 2]                      # exact score ties between item i and i + nI/2
        V[5] = 0
    uid = rng.integers(0, nU, n_eval)
    pos, mask = [], []
    for e in range(n_eval):
        la = int(rng.integers(0, 25)) if e % 7 else 0       # some rows have no positives (Q6)
        pos.append(rng.permutation(nI)[:la].tolist())
        lm = int(rng.integers(0, 60))
        if heavy and e % 5 == 0:
            lm = int(nI * 0.9)                               # > kMCap masked pairs per tile -> CSR fallback
        m = rng.permutation(nI)[:lm].tolist()
        if e % 11 == 0 and lm:
            m = m + m[:3]                                    # duplicates in mask_items
        mask.append(m)
    return U, V, build_eval_csr(uid, pos, mask, nI)


@pytest.mark.parametrize("nU,nI,d,n_eval,K", [(64, 100, 64, 1, 10), (300, 1000, 64, 129, 10), (500, 3001, 64, 400, 20),
                                               (200, 777, 256, 130, 10), (128, 640, 32, 128, 1), (90, 512, 128, 77, 32),
                                               (50, 33, 20, 40, 10), (400, 5000, 64, 300, 16),
                                               (300, 2000, 64, 200, 5), (300, 2000, 96, 150, 13), (150, 900, 64, 100, 2)])
@pytest.mark.parametrize("mode", MODES)
def test_eval_random_vs_oracle(nU, nI, d, n_eval, K, mode):
    rng = np.random.default_rng(nI + d)
    U, V, csr = _random_problem(rng, nU, nI, d, n_eval)
    check_vs_oracle(U, V, csr, K, mode)
    if mode == "tc":
        assert fallback_rows() <= max(1, n_eval // 50)       # the filter decides (almost) every row itself


@pytest.mark.parametrize("nU,nI,d,n_eval,K", [(300, 1500, 512, 260, 10), (200, 900, 1024, 130, 10), (150, 700, 320, 77, 32),
                                               (120, 600, 300, 129, 10), (100, 400, 1024, 1, 5)])
def test_eval_wide_tables_vs_oracle(nU, nI, d, n_eval, K):
    """Widths beyond the shared-memory user tile (the reference's mf_sweep_config.yaml goes to embed_size 1,024; NGCF
    concatenates d * (num_orders + 1) columns): the tail of the transposed user tile is read from the HBM workspace.
    Same fma chain -> ids, scores and metric terms stay bit-exact."""
    rng = np.random.default_rng(nI + d)
    U, V, csr = _random_problem(rng, nU, nI, d, n_eval)
    check_vs_oracle(U, V, csr, K, "exact")
    check_vs_oracle(U, V, csr, K, None)            # default dispatch must pick a kernel that takes the width


@pytest.mark.parametrize("mode", MODES)
def test_eval_heavy_masks_and_ties(mode):
    rng = np.random.default_rng(5)
    U, V, csr = _random_problem(rng, 100, 2000, 64, 150, heavy=True, ties=True)
    topk, _ = check_vs_oracle(U, V, csr, 10, mode)
    # masked items never appear unless fewer than K unmasked items exist
    for e in range(csr.n_eval):
        m = set(csr.mask_idx[csr.mask_ptr[e]:csr.mask_ptr[e + 1]].tolist())
        if 2000 - len(m) >= 10:
            assert not (set(topk[e].tolist()) & m)


def test_eval_more_masked_than_catalog_minus_k():
    """Fewer than K unmasked items: masked ones (-3.40282e+38) fill the tail, id ascending."""
    from yelprecommendation_b200.data.graph import build_eval_csr
    rng = np.random.default_rng(9)
    U, V = rng.standard_normal((4, 64)).astype(np.float32), rng.standard_normal((40, 64)).astype(np.float32)
    csr = build_eval_csr([0, 1], [[1, 2], [3]], [list(range(35)), list(range(40))], 40)
    check_vs_oracle(U, V, csr, 10, "exact")
    check_vs_oracle(U, V, csr, 10, "tc")
    assert fallback_rows() == 2          # fewer than K unmasked items -> both rows go to the exact kernel


def test_tc_filter_adversarial_near_ties():
    """Scores packed inside the TF32 error window: hundreds of items within 2*eps of the K-th best. The filter must
    hand those rows to the exact kernel (candidate overflow) and the result must still be bit-exact."""
    from yelprecommendation_b200.data.graph import build_eval_csr
    rng = np.random.default_rng(3)
    d, nI = 64, 3000
    base = rng.standard_normal(d).astype(np.float32)
    V = (base[None, :] + 1e-4 * rng.standard_normal((nI, d))).astype(np.float32)     # near-identical items
    V[::7] = V[3]                                                                     # plus exact duplicates
    U = rng.standard_normal((64, d)).astype(np.float32)
    csr = build_eval_csr(np.arange(64), [[1, 5, 9]] * 64, [[2, 3]] * 64, nI)
    check_vs_oracle(U, V, csr, 10, "tc")
    assert fallback_rows() > 0
    # moderately close scores: the window holds a few extra candidates, no fallback needed
    V2 = (base[None, :] * 0.1 + rng.standard_normal((nI, d))).astype(np.float32)
    check_vs_oracle(U, V2, csr, 10, "tc")
    assert fallback_rows() == 0


def test_trainer_evaluate_dataframe_interface():
    """MFTrainer.evaluate(eval_data DataFrame) + _generate_top_k_recommendation, as train.py calls them."""
    import pandas as pd
    from yelprecommendation_b200.trainers import MFTrainer
    g = load_npz("mf_small.npz")
    tr = MFTrainer(cfg(), int(g["num_items"]), int(g["num_users"]))
    with torch.no_grad():
        tr.model.user_embedding.weight.copy_(torch.from_numpy(g["eval_U"].astype(np.float32)))
        tr.model.item_embedding.weight.copy_(torch.from_numpy(g["eval_V"]))
    uid = g["valid_eval_uid"]
    ev = pd.DataFrame({"pos_items": lists_from(g, "valid_eval", "pos_items"),
                       "mask_items": lists_from(g, "valid_eval", "mask_items")}, index=pd.Index(uid, name="user_id"))
    got = tr.evaluate(ev, "valid")
    assert np.allclose(got, g["valid_eval_metrics"], rtol=2e-2)
    ref = g["valid_eval_topk"]
    same = sum(np.array_equal(tr.last_topk[r].cpu().numpy(), ref[r]) for r in range(len(uid)))
    assert same >= len(uid) - max(1, len(uid) // 50)
    # single-row helper
    items = torch.arange(int(g["num_items"]))
    for r in (0, 3, 11):
        pred = tr.model(torch.full_like(items, int(uid[r])), items)
        top = tr._generate_top_k_recommendation(pred, ev.iloc[r]["mask_items"])
        assert np.array_equal(top, tr.last_topk[r].cpu().numpy())


def test_metric_module_vs_reference_cases():
    from yelprecommendation_b200 import metric
    for c in metric_cases():
        a, p, k = c["actual"], c["predicted"], c["k"]
        assert isclose(metric.precision_at_k(a, p, k), c["precision"], rel_tol=1e-12, abs_tol=1e-15)
        assert isclose(metric.recall_at_k(a, p, k), c["recall"], rel_tol=1e-12, abs_tol=1e-15)
        assert isclose(metric.map_at_k(a, p, k), c["map"], rel_tol=1e-12, abs_tol=1e-15)
        assert isclose(metric.ndcg_at_k(a, p, k), c["ndcg"], rel_tol=1e-12, abs_tol=1e-15)


def test_full_catalog_yelp_shape_properties():
    """BASELINE config 3 shape: all 31,668 users x 38,048 items. Size-independent checks + sampled oracle rows."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_eval_csr
    inter = syn.make_interactions()
    assert inter.user.size == 1_561_406 and inter.num_users == 31_668 and inter.num_items == 38_048
    split = syn.split_per_user(inter, seed=42)
    uid, pos, mask = syn.eval_lists(split, "valid")
    csr = build_eval_csr(uid, pos, mask, inter.num_items)
    U, V = syn.planted_embeddings(inter)
    topk, tsc, um, sums = run_gpu(U, V, csr, 10, "tc")
    n_fb = fallback_rows()
    topk_x, tsc_x, um_x, sums_x = run_gpu(U, V, csr, 10, "exact")
    assert np.array_equal(topk, topk_x) and np.array_equal(tsc, tsc_x) and np.array_equal(um, um_x)   # all 31,668 rows
    assert np.array_equal(sums, sums_x)
    assert n_fb <= len(uid) // 100
    assert topk.shape == (len(uid), 10) and topk.min() >= 0 and topk.max() < inter.num_items
    assert np.all(np.diff(tsc, axis=1) <= 0)                                  # best first
    assert np.all(np.sort(topk, axis=1)[:, 1:] != np.sort(topk, axis=1)[:, :-1])   # distinct ids
    rows = np.sort(np.random.default_rng(0).choice(len(uid), 2000, replace=False))     # 2,000 oracle rows (round 1: 96)
    for e in rows:                                                             # never a masked item
        assert not (set(topk[e].tolist()) & set(mask[e]))
    sub = build_eval_csr(uid[rows], [pos[e] for e in rows], [mask[e] for e in rows], inter.num_items)
    otopk, otsc, oum, _ = cport.eval_topk_metrics(U, V, sub.eval_uid, sub.mask_ptr, sub.mask_idx, sub.act_ptr, sub.act_idx, 10)
    assert np.array_equal(topk[rows], otopk) and np.array_equal(tsc[rows], otsc) and np.array_equal(um[rows], oum)
    # the NGCF form of the same evaluation: concatenated width d_eff = 256 (d = 64 x 4 layer outputs), all rows tensor-core vs
    # exact kernel, 400 rows against the oracle
    rng = np.random.default_rng(1)
    U4 = np.concatenate([U] + [(U * s_ + 0.05 * rng.standard_normal(U.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
    V4 = np.concatenate([V] + [(V * s_ + 0.05 * rng.standard_normal(V.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
    assert U4.shape[1] == 256
    t4, s4, m4, sums4 = run_gpu(U4, V4, csr, 10, "tc")
    t4x, s4x, m4x, sums4x = run_gpu(U4, V4, csr, 10, "exact")
    assert np.array_equal(t4, t4x) and np.array_equal(s4, s4x) and np.array_equal(m4, m4x) and np.array_equal(sums4, sums4x)
    r4 = rows[:400]
    sub4 = build_eval_csr(uid[r4], [pos[e] for e in r4], [mask[e] for e in r4], inter.num_items)
    o4, os4, om4, _ = cport.eval_topk_metrics(U4, V4, sub4.eval_uid, sub4.mask_ptr, sub4.mask_idx, sub4.act_ptr, sub4.act_idx, 10)
    assert np.array_equal(t4[r4], o4) and np.array_equal(s4[r4], os4) and np.array_equal(m4[r4], om4)
    # checksum of checksums: metric sums equal the sum of the per-row terms
    assert np.allclose(sums[:4], um.sum(axis=0), rtol=1e-12)
    assert sums[0] / len(uid) > 0.01                                          # planted structure is recoverable


def test_chunked_evaluator_for_non_dot_product_models_like_dcn():
    """the DCN trainer's chunked evaluator (trainers/dcn_trainer.py:145-203) on the fused kernels — any scoring
    model, catalog scored in chunks of cfg.batch_size items, train items masked with 0 (sigmoid outputs, :191), top-K by
    (score desc, id asc), the reference's metrics; 'valid' evaluates the first 1,000 rows of the frame (:152-153)."""
    import pandas as pd
    from oracle import torch_port as tp
    from yelprecommendation_b200.trainers import ChunkedTopKEvaluator
    rng = np.random.default_rng(5)
    nU, nI, K, d = 1500, 2311, 10, 16
    A = torch.from_numpy(rng.standard_normal((nU, d)).astype(np.float32)).cuda()
    Bm = torch.from_numpy(rng.standard_normal((nI, d)).astype(np.float32)).cuda()
    cat = torch.from_numpy(rng.integers(0, 7, nI)).cuda()
    bias = torch.from_numpy(rng.standard_normal(7).astype(np.float32)).cuda()

    def score_fn(uid, items):                               # a cross-feature model: not a dot product of two tables
        x = A[uid].unsqueeze(1) * Bm[items].unsqueeze(0)    # [R, C, d]
        out = torch.sigmoid(0.1 * (x.sum(-1) + 0.3 * (x * x).sum(-1)) + bias[cat[items]].unsqueeze(0))
        seen[(int(uid[0]), int(items[0]))] = (uid.cpu().numpy(), items.cpu().numpy(), out.cpu().numpy())   # what the model returned
        return out

    seen = {}

    users = rng.permutation(nU)[:1200]
    pos = [rng.permutation(nI)[: int(rng.integers(1, 9))].tolist() for _ in users]
    mask = [rng.permutation(nI)[: int(rng.integers(0, 30))].tolist() for _ in users]
    frame = pd.DataFrame({"pos_items": pos, "mask_items": mask}, index=pd.Index(users, name="user_id"))
    ev = ChunkedTopKEvaluator(nI, K, "cuda", chunk_size=256, mask_value=0.0, rows_per_block=50)
    for mode, n_rows in (("valid", 1000), ("test", 1200)):
        seen.clear()
        got = ev.evaluate(score_fn, frame, mode=mode)
        full = np.full((n_rows, nI), np.nan, np.float32)     # the catalog scores exactly as the chunked calls produced them
        row_of = {int(u): r for r, u in enumerate(users[:n_rows])}
        for uid_c, items_c, out_c in seen.values():
            full[np.array([row_of[int(u)] for u in uid_c])[:, None], items_c[None, :]] = out_c
        assert not np.isnan(full).any()
        predicted = []
        for r in range(n_rows):
            s = full[r].copy()
            s[np.asarray(mask[r], dtype=np.int64)] = 0.0
            order = np.lexsort((np.arange(nI), -s))[:K]
            predicted.append(order)
        assert np.array_equal(ev.last_topk.cpu().numpy(), np.stack(predicted)), mode
        want = tp.all_metrics(pos[:n_rows], predicted, K)
        assert np.allclose(got, want, rtol=1e-12), (mode, got, want)


@pytest.mark.gpu
def test_item_sliced_evaluation_is_identical():
    """A row shard that cannot fill the GPU (here 1/8 of the Yelp-shape users = 31 user tiles on 148 SMs, what a rank of an
    8-GPU run evaluates) is cut into S item slices evaluated by S concurrent tensor-core launches + yr_topk_merge: same top-K
    ids in the same order, same exact scores, same per-row metric terms and sums as the unsliced call — MF (d = 64) and the
    NGCF form (d_eff = 256)."""
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_eval_csr
    inter = syn.make_interactions()
    split = syn.split_per_user(inter, seed=42)
    uid, pos, mask = syn.eval_lists(split, "valid")
    lo, hi = 3 * 3959, 4 * 3959
    csr = build_eval_csr(uid[lo:hi], pos[lo:hi], mask[lo:hi], inter.num_items)
    U, V = syn.planted_embeddings(inter)
    rng = np.random.default_rng(2)
    U4 = np.concatenate([U] + [(U * s_ + 0.05 * rng.standard_normal(U.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
    V4 = np.concatenate([V] + [(V * s_ + 0.05 * rng.standard_normal(V.shape)).astype(np.float32) for s_ in (0.7, 0.4, 0.2)], axis=1)
    dev = torch.device("cuda")
    assert ops.eval_item_slices(hi - lo, inter.num_items, 10, dev) >= 4          # what the default would pick for this shard
    for Ue, Ve in ((U, V), (U4, V4)):
        ecsr = ops.DeviceEvalCSR(csr, dev, 10)
        Ud, Vd = torch.from_numpy(Ue).to(dev), torch.from_numpy(Ve).to(dev)
        ref = ops.eval_topk_metrics(Ud, Vd, ecsr, mode="tc", slices=1)
        assert ops.eval_topk_metrics.last_slices == 1
        for S in (2, 4, 6):
            got = ops.eval_topk_metrics(Ud, Vd, ecsr, mode="tc", slices=S)
            assert ops.eval_topk_metrics.last_slices == S
            for a, b in zip(ref[:4], got[:4]):
                assert torch.equal(a, b), S
            assert int(got[4].item()) == 0
        auto = ops.eval_topk_metrics(Ud, Vd, ecsr, mode="tc")
        assert ops.eval_topk_metrics.last_slices >= 4 and torch.equal(auto[0], ref[0]) and torch.equal(auto[3], ref[3])
