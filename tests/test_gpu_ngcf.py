"""GPU parity: NGCF propagation / backward / train step vs the oracle and the reference goldens."""
from math import isclose

import numpy as np
import pytest
import torch

from oracle import cport, torch_port as tp
from util import rel_fro, RTOL, batches_from, cfg, load_npz, rel_err

pytestmark = pytest.mark.gpu


def _golden():
    from yelprecommendation_b200.data.graph import coo_to_csr
    g = load_npz("ngcf_small.npz")
    nU, nI = int(g["bin_nU"]), int(g["bin_nI"])
    idx, val = g["bin_L_idx"], g["bin_L_val"]
    L = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), size=(nU + nI, nU + nI)).coalesce()
    csr = coo_to_csr(idx[0], idx[1], val, nU + nI)
    csrT = coo_to_csr(idx[1], idx[0], val, nU + nI)
    W1 = [g[f"ngcf_init_W1.{l}.weight"] for l in range(3)]
    W2 = [g[f"ngcf_init_W2.{l}.weight"] for l in range(3)]
    return g, nU, nI, L, csr, csrT, g["ngcf_init_embedding.weight"], W1, W2


def _trainer(g, nU, nI, L, name="sgd", lr=1e-2, wd=0.0, **extra):
    from yelprecommendation_b200.trainers import NGCFTrainer
    tr = NGCFTrainer(cfg(optimizer=name, lr=lr, weight_decay=wd, num_orders=3, **extra), nI, nU, L)
    sd = {k[len("ngcf_init_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("ngcf_init_")}
    tr.model.load_state_dict(sd)          # identical state_dict keys to the reference model
    return tr


def test_spmm_bit_exact_vs_oracle():
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian, laplacian_to_csr
    inter = syn.make_interactions(num_users=700, num_items=900, nnz=20000, seed=4, n_clusters=4, star_ratings=True)
    L = build_laplacian(inter.user, inter.item, inter.rating, 700, 900)
    csr = laplacian_to_csr(L, "cuda")
    rng = np.random.default_rng(0)
    for d in (32, 64, 128, 256):
        X = rng.standard_normal((1600, d)).astype(np.float32)
        Y0 = rng.standard_normal((1600, d)).astype(np.float32)
        assert csr.fwd.n_split_rows > 0 and not csr.symmetric        # long rows are cut into 128-nnz chunks
        c = (csr.fwd.rowptr.cpu().numpy(), csr.fwd.col.cpu().numpy(), csr.fwd.val.cpu().numpy())
        y = ops.spmm_csr(csr.fwd, torch.from_numpy(X).cuda())
        assert np.array_equal(y.cpu().numpy(), cport.spmm_csr(*c, X))
        ya = torch.from_numpy(Y0).cuda()
        ops.spmm_csr(csr.bwd, torch.from_numpy(X).cuda(), out=ya, accumulate=True)
        ct = (csr.bwd.rowptr.cpu().numpy(), csr.bwd.col.cpu().numpy(), csr.bwd.val.cpu().numpy())
        assert np.array_equal(ya.cpu().numpy(), cport.spmm_csr(*ct, X, Y0))


def test_layer_forward_bit_exact_vs_oracle_and_golden():
    """dense_mode 0 (YR_DENSE_FP32): FP32-pipe transforms, bit-comparable with the oracle. The mode is an argument of the
    call (per trainer: cfg.ngcf_dense_mode) — nothing process-wide."""
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data.graph import laplacian_to_csr
    g, nU, nI, L, csr, csrT, E0, W1, W2 = _golden()
    dcsr = laplacian_to_csr(L, "cuda")
    E = torch.from_numpy(E0).cuda()
    Ec = E0
    for l in range(3):
        En, LE = ops.ngcf_layer_fwd(dcsr, E, torch.from_numpy(W1[l]).cuda(), torch.from_numpy(W2[l]).cuda(), dense_mode=0)
        Eo, LEo = cport.ngcf_layer_fwd(csr, Ec, W1[l], W2[l])
        assert np.array_equal(LE.cpu().numpy(), LEo)
        assert np.array_equal(En.cpu().numpy(), Eo)                      # same fma chain order -> bit-exact
        assert rel_err(En.cpu().numpy(), g[f"ngcf_layer{l + 1}"]) < RTOL   # vs the reference (torch.eye and all)
        E, Ec = En, Eo


def test_layer_forward_tensor_core_3xtf32():
    """Default mode: tcgen05 TF32 MMAs with the hi/lo split. Must stay well inside the 1e-5 bar, on the golden graph
    and on a 20k-row graph with a ragged last tile."""
    from yelprecommendation_b200 import _cabi, ops
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian, laplacian_to_csr
    assert _cabi.YR_DENSE_TC_FWD == 1            # the default of ops.ngcf_layer_fwd (cfg.ngcf_dense_mode defaults to 2)
    g, nU, nI, L, csr, csrT, E0, W1, W2 = _golden()
    dcsr = laplacian_to_csr(L, "cuda")
    E, Ec = torch.from_numpy(E0).cuda(), E0
    for l in range(3):
        En, LE = ops.ngcf_layer_fwd(dcsr, E, torch.from_numpy(W1[l]).cuda(), torch.from_numpy(W2[l]).cuda())
        Eo, LEo = cport.ngcf_layer_fwd(csr, Ec, W1[l], W2[l])
        assert np.array_equal(LE.cpu().numpy(), LEo)
        assert rel_err(En.cpu().numpy(), Eo) < 2e-6
        assert rel_err(En.cpu().numpy(), g[f"ngcf_layer{l + 1}"]) < RTOL
        E, Ec = torch.from_numpy(Eo).cuda(), Eo          # same input for both on the next layer
    inter = syn.make_interactions(num_users=9000, num_items=11077, nnz=300_000, seed=12, n_clusters=8)
    Lb = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    big = laplacian_to_csr(Lb, "cuda")
    rng = np.random.default_rng(0)
    X = rng.standard_normal((20077, 64)).astype(np.float32)
    A, B = (rng.standard_normal((64, 64)).astype(np.float32) * 0.2 for _ in range(2))
    En, _ = ops.ngcf_layer_fwd(big, torch.from_numpy(X).cuda(), torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
    c = (big.fwd.rowptr.cpu().numpy(), big.fwd.col.cpu().numpy(), big.fwd.val.cpu().numpy())
    Eo, _ = cport.ngcf_layer_fwd(c, X, A, B)
    assert rel_err(En.cpu().numpy(), Eo) < 2e-6


@pytest.mark.parametrize("dense_mode", [1, 2])
def test_layer_backward_vs_oracle_and_autograd_golden(dense_mode):
    """dense_mode 1: tensor-core forward + FP32-pipe backward kernel; 2 (default of the trainers): tcgen05 backward (3xTF32, MN-major operands for dW)."""
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data.graph import laplacian_to_csr
    _layer_backward_case(ops, laplacian_to_csr, dense_mode)


def test_train_steps_with_tensor_core_backward():
    """Whole train steps (row-sparse last layer included) with the FP32-pipe backward (cfg.ngcf_dense_mode = 1) and with
    FP32-pipe transforms throughout (0) vs the reference golden — the default is 2, tensor cores for both; a second trainer
    built afterwards with the default config is unaffected (the mode is per trainer)."""
    test_train_steps_vs_reference_golden(1, ngcf_dense_mode=1)
    test_train_steps_vs_reference_golden(1, ngcf_dense_mode=0)
    g, nU, nI, L, *_ = _golden()
    assert _trainer(g, nU, nI, L)._state()[0].dense_mode == 2


def _layer_backward_case(ops, laplacian_to_csr, dense_mode=1):
    g, nU, nI, L, csr, csrT, E0, W1, W2 = _golden()
    dcsr = laplacian_to_csr(L, "cuda")
    layers, LEs = [E0], []
    for l in range(3):
        E, LE = cport.ngcf_layer_fwd(csr, layers[-1], W1[l], W2[l])
        layers.append(E)
        LEs.append(LE)
    u, p, n = g["tri_u"][:128], g["tri_p"][:128], g["tri_n"][:128]
    _, _, _, G = cport.ngcf_tail(layers, nU, u, p, n, want_grad=True)
    Gd = [torch.from_numpy(x).cuda() for x in G]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for l in (2, 1, 0):
        G[l], dW1o, dW2o = cport.ngcf_layer_bwd(csrT, layers[l], LEs[l], layers[l + 1], G[l + 1], W1[l], W2[l], G[l])
        dW1, dW2 = ops.ngcf_layer_bwd(dcsr, cu(layers[l]), cu(LEs[l]), cu(layers[l + 1]), Gd[l + 1], cu(W1[l]), cu(W2[l]), Gd[l],
                                      dense_mode=dense_mode)
        assert rel_err(Gd[l].cpu().numpy(), G[l]) < RTOL
        assert rel_err(dW1.cpu().numpy(), dW1o) < RTOL and rel_err(dW2.cpu().numpy(), dW2o) < RTOL
        assert rel_err(dW1.cpu().numpy(), g[f"ngcf_grad_W1.{l}.weight"]) < 2e-5
        assert rel_err(dW2.cpu().numpy(), g[f"ngcf_grad_W2.{l}.weight"]) < 2e-5
    assert rel_err(Gd[0].cpu().numpy(), g["ngcf_grad_embedding.weight"]) < 2e-5


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_train_steps_vs_reference_golden(ci, **extra):
    g, nU, nI, L, *_ = _golden()
    name, (lr, wd) = str(g[f"ngcf_{ci}_name"]), g[f"ngcf_{ci}_cfg"]
    tr = _trainer(g, nU, nI, L, name, float(lr), float(wd), **extra)
    batches = batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 128, limit=3)
    if ci == 0:
        assert isclose(tr.validate(batches), float(g["ngcf_valid0"]), rel_tol=RTOL)
    total = tr.train(batches)
    assert rel_err(tr.last_step_losses.cpu().numpy(), g[f"ngcf_{ci}_losses"]) < RTOL
    assert isclose(total, float(np.sum(g[f"ngcf_{ci}_losses"])), rel_tol=RTOL)
    sd = tr.model.state_dict()
    assert rel_err(sd["embedding.weight"].cpu().numpy(), g[f"ngcf_{ci}_final_embedding.weight"]) < RTOL
    for l in range(3):
        assert rel_err(sd[f"W1.{l}.weight"].cpu().numpy(), g[f"ngcf_{ci}_final_W1.{l}.weight"]) < 5e-5
        assert rel_err(sd[f"W2.{l}.weight"].cpu().numpy(), g[f"ngcf_{ci}_final_W2.{l}.weight"]) < 5e-5


@pytest.mark.parametrize("name,lr", [("sgd", 0.05), ("adam", 1e-3)])
def test_row_sparse_top_layer_equals_dense_step(name, lr):
    """yr_ngcf_train_step: last-layer forward/backward on the <= 3B batch rows + scatter form of L^T T (default) vs the
    dense layer everywhere (cfg.ngcf_top_rows_mode = 0, per trainer) — same sums in a different fp32 order: 1e-6
    norm-wise here. validate()/evaluate() after a row-sparse step must see a fully recomputed propagation (no stale rows)."""
    g, nU, nI, L, *_ = _golden()
    batches = batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 128, limit=4)
    out = {}
    for mode in (1, 0):
        tr = _trainer(g, nU, nI, L, name, lr, 0.0, ngcf_top_rows_mode=mode)
        total = tr.train(batches)
        sd = {k: v.detach().cpu().numpy().copy() for k, v in tr.model.state_dict().items()}
        out[mode] = (total, tr.last_step_losses.cpu().numpy().copy(), sd, tr.validate(batches))
    (ta, la, sa, va), (tb, lb, sb, vb) = out[1], out[0]
    assert isclose(ta, tb, rel_tol=1e-6) and rel_err(la, lb) < 1e-6 and isclose(va, vb, rel_tol=1e-6)
    for k in sb:
        assert rel_fro(sa[k], sb[k]) < 1e-6, k


def test_dropin_model_signatures_and_autograd():
    """NGCF.forward / bpr_forward / embedding_propagation with the reference's signatures + loss.backward()."""
    from yelprecommendation_b200.loss import BPRLoss
    g, nU, nI, L, *_ = _golden()
    tr = _trainer(g, nU, nI, L)
    m = tr.model
    assert sorted(m.state_dict().keys()) == sorted(["embedding.weight"] + [f"W{j}.{l}.weight" for j in (1, 2) for l in range(3)])
    Ld = L.cuda()
    u, p, n = (torch.from_numpy(g[k][:128]) for k in ("tri_u", "tri_p", "tri_n"))
    pos, neg = m.bpr_forward(u, p, n, Ld)
    assert rel_err(pos.detach().cpu().numpy(), g["ngcf_pos0"]) < RTOL
    assert rel_err(neg.detach().cpu().numpy(), g["ngcf_neg0"]) < RTOL
    BPRLoss()(pos, neg).backward()
    for k, prm in m.named_parameters():
        assert rel_err(prm.grad.cpu().numpy(), g["ngcf_grad_" + k]) < 2e-5, k
    items = torch.arange(nI)
    for j, usr in enumerate(g["ngcf_eval_users"]):
        s = m(torch.full_like(items, int(usr)), items, Ld)
        assert rel_err(s.detach().cpu().numpy(), g["ngcf_eval_scores"][j]) < RTOL
    e1 = m.embedding_propagation(m.embedding.weight, m.W1[0], m.W2[0], Ld)
    assert rel_err(e1.detach().cpu().numpy(), g["ngcf_layer1"]) < RTOL


def test_evaluate_vs_port_and_sample100_mode():
    import pandas as pd
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_eval_csr
    g, nU, nI, L, csr, csrT, E0, W1, W2 = _golden()
    inter = syn.Interactions(nU, nI, g["bin_user"], g["bin_item"], g["bin_rating"], None, None)
    split = syn.split_per_user(inter, seed=42)
    uid, pos, mask = syn.eval_lists(split, "valid")
    tr = _trainer(g, nU, nI, L)
    got = tr.evaluate(build_eval_csr(uid, pos, mask, nI))
    port = tp.NGCFPort(torch.from_numpy(E0), [torch.from_numpy(w) for w in W1], [torch.from_numpy(w) for w in W2], nU, L)
    want, ppred = port.evaluate(uid, mask, pos)
    same = sum(np.array_equal(tr.last_topk[r].cpu().numpy(), ppred[r]) for r in range(len(uid)))
    assert same >= len(uid) - max(1, len(uid) // 25)
    assert np.allclose(got, want, rtol=5e-2)
    # C oracle on the concatenated embeddings: bit-exact ids
    layers = [E0]
    for l in range(3):
        layers.append(cport.ngcf_layer_fwd(csr, layers[-1], W1[l], W2[l])[0])
    cat = np.concatenate(layers, axis=1)
    ec = build_eval_csr(uid, pos, mask, nI)
    otopk, _, _, osums = cport.eval_topk_metrics(cat[:nU], cat[nU:], ec.eval_uid, ec.mask_ptr, ec.mask_idx, ec.act_ptr, ec.act_idx, 10)
    assert np.array_equal(tr.last_topk.cpu().numpy(), otopk)
    assert np.allclose(got, cport.metrics_from_sums(osums, ec.n_eval), rtol=1e-12)
    # Q12 compat: 100 rows drawn from the global NumPy RNG, with replacement
    ev = pd.DataFrame({"pos_items": pos, "mask_items": mask}, index=pd.Index(uid, name="user_id"))
    tr.cfg.ngcf_eval_mode = "sample100"
    np.random.seed(3)
    r1 = tr.evaluate(ev)
    np.random.seed(3)
    rows = np.random.randint(ev.shape[0], size=100)
    assert tr.last_topk.shape[0] == 100
    sub = build_eval_csr(uid[rows], [pos[r] for r in rows], [mask[r] for r in rows], nI)
    _, _, _, s2 = cport.eval_topk_metrics(cat[:nU], cat[nU:], sub.eval_uid, sub.mask_ptr, sub.mask_idx, sub.act_ptr, sub.act_idx, 10)
    assert np.allclose(r1, cport.metrics_from_sums(s2, 100), rtol=1e-12)


def test_medium_graph_train_step_vs_port():
    """A 6k-node degree-skewed graph, one Adam step: CUDA vs the torch-CPU port (identity hoisted)."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions(num_users=2500, num_items=3500, nnz=90_000, seed=8, n_clusters=8)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
    batches = syn.to_batches(tu, tpos, tneg, 2048)[:2]
    torch.manual_seed(5)
    tr = NGCFTrainer(cfg(optimizer="adam", lr=1e-3, num_orders=3, batch_size=2048), inter.num_items, inter.num_users, L)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(3)],
                       [sd[f"W2.{l}.weight"] for l in range(3)], inter.num_users, L, "adam", 1e-3, 0.0)
    total = tr.train(batches)
    ptotal, psteps = port.train(batches)
    assert rel_err(tr.last_step_losses.cpu().numpy(), psteps) < RTOL
    assert rel_err(tr.model.embedding.weight.detach().cpu().numpy(), port.emb.detach().numpy()) < 2e-5
    for l in range(3):
        assert rel_err(tr.model.W1[l].weight.detach().cpu().numpy(), port.W1[l].detach().numpy()) < 1e-4


def test_full_size_yelp_shape_train_steps_vs_port():
    """BASELINE configs[1] at full size (69,716 nodes, 3.12 M non-zeros, hub rows of 10,696 entries, batch 2,048): two
    SGD steps of the fused train step (tensor-core forward, row-sparse last layer, scatter form of L^T T) against the
    torch-CPU port. 1e-5 norm-wise on every parameter and on the step losses."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions()
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
    batches = syn.to_batches(tu, tpos, tneg, 2048)[:2]
    torch.manual_seed(5)
    tr = NGCFTrainer(cfg(optimizer="sgd", lr=0.05, num_orders=3, batch_size=2048), inter.num_items, inter.num_users, L)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(3)],
                       [sd[f"W2.{l}.weight"] for l in range(3)], inter.num_users, L, "sgd", 0.05, 0.0)
    total = tr.train(batches)
    ptotal, psteps = port.train(batches)
    assert rel_err(tr.last_step_losses.cpu().numpy(), psteps) < RTOL
    assert isclose(total, ptotal, rel_tol=RTOL)
    assert rel_fro(tr.model.embedding.weight.detach().cpu().numpy(), port.emb.detach().numpy()) < RTOL
    # the update itself (what the step computed), not just the barely-moved table. A typical element moves by 8e-6
    # while the table's fp32 ulp is 1.2e-7 (1.5 % of that), so the recovered update carries quantisation noise:
    # measured 2.4e-5 with the FP32-pipe forward, 6.4e-5 with the 3xTF32 forward, identical for the dense and the
    # row-sparse last layer (scripts/diag_ngcf_full.py).
    d_ours = tr.model.embedding.weight.detach().cpu().numpy() - sd["embedding.weight"].numpy()
    d_port = port.emb.detach().numpy() - sd["embedding.weight"].numpy()
    assert rel_fro(d_ours, d_port) < 2e-4
    for l in range(3):
        assert rel_fro(tr.model.W1[l].weight.detach().cpu().numpy(), port.W1[l].detach().numpy()) < RTOL
        assert rel_fro(tr.model.W2[l].weight.detach().cpu().numpy(), port.W2[l].detach().numpy()) < RTOL


def test_batch_larger_than_row_scratch_falls_back_to_dense_last_layer():
    """A batch beyond the trainer's row-scratch capacity (3 * 4096 rows) runs the last layer densely — same result."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions(num_users=2500, num_items=3500, nnz=90_000, seed=8, n_clusters=8)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
    batches = syn.to_batches(tu, tpos, tneg, 5000)[:2]
    torch.manual_seed(5)
    tr = NGCFTrainer(cfg(optimizer="sgd", lr=0.05, num_orders=3, batch_size=2048), inter.num_items, inter.num_users, L)
    assert tr._row_cap < 5000
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(3)],
                       [sd[f"W2.{l}.weight"] for l in range(3)], inter.num_users, L, "sgd", 0.05, 0.0)
    total = tr.train(batches)
    ptotal, psteps = port.train(batches)
    assert rel_err(tr.last_step_losses.cpu().numpy(), psteps) < RTOL
    assert rel_fro(tr.model.embedding.weight.detach().cpu().numpy(), port.emb.detach().numpy()) < RTOL


def test_ngcf_trainer_consumes_device_loader():
    """NGCFTrainer.train(DeviceTripleLoader) == the same triples fed as host batches (1e-6: atomics order only)."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.data.sampler import DeviceTripleLoader
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions(num_users=900, num_items=700, nnz=20_000, seed=6, n_clusters=4)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    ld = DeviceTripleLoader.from_split(split, inter.num_items, batch_size=1024, seed=3)
    mk = lambda: NGCFTrainer(cfg(optimizer="sgd", lr=0.05, num_orders=3, batch_size=1024), inter.num_items, inter.num_users, L)
    torch.manual_seed(1)
    a = mk()
    torch.manual_seed(1)
    b = mk()
    la = a.train(ld)
    ld.set_epoch(0)
    u, p, n = (x.cpu().numpy() for x in ld.epoch_triples())
    lb = b.train(syn.to_batches(u, p, n, 1024))
    assert isclose(la, lb, rel_tol=1e-6)
    assert rel_fro(a.model.embedding.weight.detach().cpu().numpy(), b.model.embedding.weight.detach().cpu().numpy()) < 1e-6


def test_per_batch_train_calls_with_prefetched_prefix_equal_one_call():
    """train([b]) per batch (what train.py's epoch loop amounts to with a 1-batch loader) pre-propagates the
    batch-independent layers of the next step behind the loss read-back; the result must equal one train(batches) call,
    also when validate() / load_state_dict() happen in between (the latter must invalidate the prefetched layers)."""
    g, nU, nI, L, *_ = _golden()
    batches = batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 128, limit=5)
    a = _trainer(g, nU, nI, L, "adam", 1e-3, 0.0)
    b = _trainer(g, nU, nI, L, "adam", 1e-3, 0.0)
    total_a = a.train(batches)
    total_b = 0.0
    for i, bt in enumerate(batches):
        total_b += b.train([bt])
        assert b._prefix_tag is not None                       # the next step's first layers are in flight / done
        if i == 1:
            b.validate(batches[:1])                            # recomputes every layer; prefix stays valid
        if i == 2:                                             # a Python-side parameter write must invalidate it
            sd = {k: v.clone() for k, v in b.model.state_dict().items()}
            b.model.load_state_dict(sd)
            assert b._take_prefix() == 0
    assert isclose(total_a, total_b, rel_tol=1e-6)
    for (ka, va), (kb, vb) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
        assert rel_fro(vb.cpu().numpy(), va.cpu().numpy()) < 1e-6, ka


@pytest.mark.parametrize("layers", [1, 2, 4])
def test_other_layer_counts_vs_port(layers):
    """num_orders != 3: the row-sparse last layer is the ONLY layer at 1, the prefetched prefix has 0 / 1 / 3 layers."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions(num_users=700, num_items=900, nnz=20_000, seed=4, n_clusters=4)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
    batches = syn.to_batches(tu, tpos, tneg, 1024)[:3]
    torch.manual_seed(7)
    tr = NGCFTrainer(cfg(optimizer="sgd", lr=0.05, num_orders=layers, batch_size=1024), inter.num_items, inter.num_users, L)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(layers)],
                       [sd[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, "sgd", 0.05, 0.0)
    total = sum(tr.train([b]) for b in batches)               # per-batch calls: exercises the prefetched prefix too
    ptotal, psteps = port.train(batches)
    assert isclose(total, ptotal, rel_tol=RTOL)
    assert rel_fro(tr.model.embedding.weight.detach().cpu().numpy(), port.emb.detach().numpy()) < RTOL
    for l in range(layers):
        assert rel_fro(tr.model.W1[l].weight.detach().cpu().numpy(), port.W1[l].detach().numpy()) < RTOL
        assert rel_fro(tr.model.W2[l].weight.detach().cpu().numpy(), port.W2[l].detach().numpy()) < RTOL
    assert isclose(tr.validate(batches[:1]), port.validate(batches[:1]), rel_tol=RTOL)


@pytest.mark.parametrize("d", [32, 128])
def test_other_embedding_widths_layer_fwd_bwd_vs_oracle(d):
    """embed_size 32 / 128. FP32-pipe dense transforms (dense_mode 0; the only mode at d = 32): forward bit-exact against
    the C oracle (same fma chain), backward at 1e-5. d = 128 also runs the tcgen05 kernels (default mode: forward < 2e-6;
    dense_mode 2: backward on tensor cores at 1e-5)."""
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian, coo_to_csr, laplacian_to_csr
    inter = syn.make_interactions(num_users=500, num_items=700, nnz=12_000, seed=d, n_clusters=4)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    n = inter.num_users + inter.num_items
    Lc = L.coalesce()
    idx, val = Lc.indices().numpy(), Lc.values().numpy()
    csr, csrT = coo_to_csr(idx[0], idx[1], val, n), coo_to_csr(idx[1], idx[0], val, n)
    dcsr = laplacian_to_csr(L, "cuda")
    rng = np.random.default_rng(d)
    E0 = rng.standard_normal((n, d)).astype(np.float32)
    W1 = (rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    W2 = (rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    Gn = rng.standard_normal((n, d)).astype(np.float32)
    G0 = rng.standard_normal((n, d)).astype(np.float32)
    cu = lambda a: torch.from_numpy(a).cuda()
    En, LE = ops.ngcf_layer_fwd(dcsr, cu(E0), cu(W1), cu(W2), dense_mode=0)
    Eo, LEo = cport.ngcf_layer_fwd(csr, E0, W1, W2)
    assert np.array_equal(LE.cpu().numpy(), LEo)
    assert np.array_equal(En.cpu().numpy(), Eo)
    Go, dW1o, dW2o = cport.ngcf_layer_bwd(csrT, E0, LEo, Eo, Gn, W1, W2, G0)
    for mode in ((0, 1, 2) if d == 128 else (0,)):
        if mode:
            En_tc, LE_tc = ops.ngcf_layer_fwd(dcsr, cu(E0), cu(W1), cu(W2), dense_mode=mode)
            assert np.array_equal(LE_tc.cpu().numpy(), LEo)
            assert rel_err(En_tc.cpu().numpy(), Eo) < 2e-6
        G = cu(G0.copy())
        dW1, dW2 = ops.ngcf_layer_bwd(dcsr, cu(E0), LE, En, cu(Gn), cu(W1), cu(W2), G, dense_mode=mode)
        assert rel_fro(G.cpu().numpy(), Go) < RTOL, mode
        assert rel_fro(dW1.cpu().numpy(), dW1o) < RTOL and rel_fro(dW2.cpu().numpy(), dW2o) < RTOL, mode


@pytest.mark.parametrize("d,layers,name,lr", [(32, 3, "sgd", 0.05), (32, 2, "adam", 1e-3), (128, 1, "sgd", 0.05),
                                              (128, 2, "adam", 1e-3)])
def test_other_embedding_widths_trainer_vs_port(d, layers, name, lr):
    """NGCFTrainer.train / validate / evaluate at embed_size 32 / 128 (BASELINE config 5 asks d = 128) against the
    torch-CPU port; evaluation ids bit-exact against the C oracle on the concatenated embeddings."""
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_eval_csr, build_laplacian
    from yelprecommendation_b200.trainers import NGCFTrainer
    inter = syn.make_interactions(num_users=700, num_items=900, nnz=20_000, seed=4, n_clusters=4)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    split = syn.split_per_user(inter, seed=42)
    tu, tpos, tneg = syn.sample_triples(split, inter.num_items, seed=42)
    batches = syn.to_batches(tu, tpos, tneg, 1024)[:3]
    torch.manual_seed(d + layers)
    tr = NGCFTrainer(cfg(optimizer=name, lr=lr, num_orders=layers, batch_size=1024, embed_size=d), inter.num_items,
                     inter.num_users, L)
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    port = tp.NGCFPort(sd["embedding.weight"], [sd[f"W1.{l}.weight"] for l in range(layers)],
                       [sd[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, name, lr, 0.0)
    total = tr.train(batches)
    ptotal, psteps = port.train(batches)
    assert isclose(total, ptotal, rel_tol=RTOL)
    assert rel_err(tr.last_step_losses.cpu().numpy(), psteps) < RTOL
    tol = RTOL if name == "sgd" else 2e-5
    assert rel_fro(tr.model.embedding.weight.detach().cpu().numpy(), port.emb.detach().numpy()) < tol
    for l in range(layers):
        assert rel_fro(tr.model.W1[l].weight.detach().cpu().numpy(), port.W1[l].detach().numpy()) < 10 * tol
        assert rel_fro(tr.model.W2[l].weight.detach().cpu().numpy(), port.W2[l].detach().numpy()) < 10 * tol
    assert isclose(tr.validate(batches[:1]), port.validate(batches[:1]), rel_tol=RTOL)
    uid, pos, mask = syn.eval_lists(split, "valid")
    ec = build_eval_csr(uid, pos, mask, inter.num_items)
    from yelprecommendation_b200 import ops
    # d * (layers + 1) = 384 (128 x 3): beyond the tensor-core kernel -> FP32-pipe kernel with the user-tile tail in HBM
    got = tr.evaluate(ec)
    cat = ops.ngcf_concat(tr.propagate()[0]).cpu().numpy()
    nU = inter.num_users
    otopk, _, _, osums = cport.eval_topk_metrics(cat[:nU], cat[nU:], ec.eval_uid, ec.mask_ptr, ec.mask_idx, ec.act_ptr,
                                                 ec.act_idx, 10)
    assert np.array_equal(tr.last_topk.cpu().numpy(), otopk)
    assert np.allclose(got, cport.metrics_from_sums(osums, ec.n_eval), rtol=1e-12)


def test_spmm_hub_rows_bit_exact_vs_oracle():
    """Rows with more than YR_SPMM_BIG_CHUNKS = 256 chunks (hub rows of the scaled config-5 graph) take the follow-up kernel
    that stages the chunk partials through shared memory: same left-to-right order, still bit-exact against the oracle."""
    from yelprecommendation_b200 import ops
    from yelprecommendation_b200.data.graph import CSRMatrix
    rng = np.random.default_rng(7)
    n = 90_000
    lens = rng.integers(0, 60, n)
    lens[5], lens[77], lens[4000], lens[n - 1] = 70_001, 33_000, 20_000, 40_321      # 547 / 258 / 157 / 316 chunks
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    col = rng.integers(0, n, int(rowptr[-1])).astype(np.int32)
    val = (rng.standard_normal(int(rowptr[-1])) * 0.05).astype(np.float32)
    A = CSRMatrix(rowptr, col, val, "cuda")
    assert A.n_big_rows == 3 and A.n_split_rows >= 4
    for d in (64, 128):
        X = rng.standard_normal((n, d)).astype(np.float32)
        Y0 = rng.standard_normal((n, d)).astype(np.float32)
        y = ops.spmm_csr(A, torch.from_numpy(X).cuda())
        assert np.array_equal(y.cpu().numpy(), cport.spmm_csr(rowptr, col, val, X))
        ya = torch.from_numpy(Y0).cuda()
        ops.spmm_csr(A, torch.from_numpy(X).cuda(), out=ya, accumulate=True)
        assert np.array_equal(ya.cpu().numpy(), cport.spmm_csr(rowptr, col, val, X, Y0))
