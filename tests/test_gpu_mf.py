"""GPU parity: BPR-MF kernels (through the Python mirror -> C ABI) vs the oracle and the reference goldens."""
from math import isclose

import numpy as np
import pytest
import torch

from oracle import cport
from util import RTOL, assert_update_close, batches_from, cfg, load_npz, rel_err

pytestmark = pytest.mark.gpu


def _trainer(g, name, lr, wd, **kw):
    from yelprecommendation_b200.trainers import MFTrainer
    tr = MFTrainer(cfg(optimizer=name, lr=lr, weight_decay=wd, **kw), int(g["num_items"]), int(g["num_users"]))
    with torch.no_grad():
        tr.model.user_embedding.weight.copy_(torch.from_numpy(g["mf_U0"]))
        tr.model.item_embedding.weight.copy_(torch.from_numpy(g["mf_V0"]))
    return tr


def test_native_library_is_loaded():
    from yelprecommendation_b200 import _cabi
    lib = _cabi.load()
    import ctypes
    n = ctypes.c_int(0)
    assert lib.yr_device_sm_count(ctypes.byref(n)) == 0 and n.value > 0
    assert any("libyelprec_b200.so" in l for l in open("/proc/self/maps"))


def test_mf_score_bit_exact_vs_oracle_and_golden():
    from yelprecommendation_b200 import ops
    g = load_npz("mf_small.npz")
    U, V = torch.from_numpy(g["mf_U0"]).cuda(), torch.from_numpy(g["mf_V0"]).cuda()
    u, p = g["tri_u"], g["tri_p"]
    s = ops.mf_score(U, V, torch.from_numpy(u), torch.from_numpy(p)).cpu().numpy()
    assert np.array_equal(s, cport.mf_score(g["mf_U0"], g["mf_V0"], u, p))          # canonical fma chain
    assert rel_err(s[:300], g["mf_score0"]) < RTOL
    for d in (8, 20, 64, 100, 256):                                                   # generic widths
        rng = np.random.default_rng(d)
        A, B = rng.standard_normal((50, d)).astype(np.float32), rng.standard_normal((70, d)).astype(np.float32)
        a, b = rng.integers(0, 50, 333), rng.integers(0, 70, 333)
        s = ops.mf_score(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(a), torch.from_numpy(b))
        assert np.array_equal(s.cpu().numpy(), cport.mf_score(A, B, a, b))


def test_mf_score_rejects_bad_ids_like_embedding():
    from yelprecommendation_b200 import ops
    U, V = torch.randn(10, 64).cuda(), torch.randn(12, 64).cuda()
    with pytest.raises(IndexError):
        ops.mf_score(U, V, torch.tensor([0, 10]), torch.tensor([0, 1]))
    with pytest.raises(IndexError):
        ops.mf_score(U, V, torch.tensor([0, 1]), torch.tensor([-1, 1]))


def test_bpr_loss_fwd_bwd():
    from yelprecommendation_b200.loss import BPRLoss
    rng = np.random.default_rng(0)
    for B in (1, 7, 2048, 5000):
        pos = torch.from_numpy(rng.standard_normal(B).astype(np.float32) * 4)
        neg = torch.from_numpy(rng.standard_normal(B).astype(np.float32) * 4)
        pc, nc = pos.clone().requires_grad_(), neg.clone().requires_grad_()
        ref = torch.mean(-torch.nn.functional.logsigmoid(pc - nc))
        ref.backward()
        pg, ng = pos.cuda().requires_grad_(), neg.cuda().requires_grad_()
        out = BPRLoss()(pg, ng)
        (out * 3.0).backward()
        assert isclose(out.item(), ref.item(), rel_tol=RTOL)
        assert isclose(out.item(), cport.bpr_loss(pos.numpy(), neg.numpy()), rel_tol=1e-6)
        assert rel_err(pg.grad.cpu().numpy(), 3.0 * pc.grad.numpy()) < RTOL
        assert rel_err(ng.grad.cpu().numpy(), 3.0 * nc.grad.numpy()) < RTOL


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_fused_training_vs_reference_golden(ci):
    g = load_npz("mf_small.npz")
    name, (lr, wd) = str(g[f"mf_{ci}_name"]), g[f"mf_{ci}_cfg"]
    tr = _trainer(g, name, float(lr), float(wd))
    batches = batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 256, limit=6)
    total = tr.train(batches)
    steps = tr.last_step_losses.cpu().numpy()
    assert rel_err(steps, g[f"mf_{ci}_losses"]) < RTOL
    assert isclose(total, float(np.sum(g[f"mf_{ci}_losses"])), rel_tol=RTOL)           # Q1: sum of batch means
    assert rel_err(tr.model.user_embedding.weight.detach().cpu().numpy(), g[f"mf_{ci}_U"]) < RTOL
    assert rel_err(tr.model.item_embedding.weight.detach().cpu().numpy(), g[f"mf_{ci}_V"]) < RTOL
    if name == "sgd":          # the SGD update is ~2e-4 of the table: compare the UPDATES themselves against the reference
        tu_, tv_ = _touches(batches, int(g["num_users"]), int(g["num_items"]))
        assert_update_close(tr.model.user_embedding.weight.detach().cpu().numpy(), g[f"mf_{ci}_U"], g["mf_U0"], "U", touches=tu_)
        assert_update_close(tr.model.item_embedding.weight.detach().cpu().numpy(), g[f"mf_{ci}_V"], g["mf_V0"], "V", touches=tv_)
    sc = tr._scratch                                                                      # all-zero invariant
    for k in ("flagU", "flagV"):
        assert int(torch.count_nonzero(sc[k]).item()) == 0, k
    # one launch per batch (steps_per_launch=1) must give the same result as one launch for all
    tr2 = _trainer(g, name, float(lr), float(wd), steps_per_launch=1)
    total2 = tr2.train(batches)
    assert isclose(total2, total, rel_tol=1e-6)
    assert rel_err(tr2.model.user_embedding.weight.detach().cpu().numpy(), g[f"mf_{ci}_U"]) < RTOL


def test_validate_and_short_last_batch():
    g = load_npz("mf_small.npz")
    tr = _trainer(g, "adam", 1e-2, 0.0)
    vb = batches_from(g["vtri_u"], g["vtri_p"], g["vtri_n"], 256)
    assert vb[-1]["user_id"].numel() < 256                                               # DataLoader keeps it
    assert isclose(tr.validate(vb), float(g["mf_valid0"]), rel_tol=RTOL)
    tr.train(batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 256, limit=6))
    assert isclose(tr.validate(vb), float(g["mf_valid_after"]), rel_tol=RTOL)
    # a training pass that ends with a short batch, checked against the C oracle
    tr = _trainer(g, "sgd", 1e-2, 0.0)
    n = 256 * 3 + 77
    b = batches_from(g["tri_u"][:n], g["tri_p"][:n], g["tri_n"][:n], 256)
    total = tr.train(b)
    orc = cport.MFTrainerOracle(g["mf_U0"], g["mf_V0"], "sgd", 1e-2, 0.0)
    ototal, _ = orc.train([{k: v.numpy() for k, v in x.items()} for x in b])
    assert isclose(total, ototal, rel_tol=RTOL)
    assert rel_err(tr.model.item_embedding.weight.detach().cpu().numpy(), orc.V) < RTOL
    tu_, tv_ = _touches(b, int(g["num_users"]), int(g["num_items"]))
    assert_update_close(tr.model.item_embedding.weight.detach().cpu().numpy(), orc.V, g["mf_V0"], "V", touches=tv_)
    assert_update_close(tr.model.user_embedding.weight.detach().cpu().numpy(), orc.U, g["mf_U0"], "U", touches=tu_)


def test_train_bad_id_raises_index_error():
    g = load_npz("mf_small.npz")
    tr = _trainer(g, "sgd", 1e-2, 0.0)
    b = batches_from(g["tri_u"][:256].copy(), g["tri_p"][:256].copy(), g["tri_n"][:256].copy(), 256)
    b[0]["pos_item"][3] = int(g["num_items"])
    with pytest.raises(IndexError):
        tr.train(b)


def test_reference_style_autograd_loop_on_dropin_modules():
    """An unmodified reference trainer body (mf_trainer.py:104-114) on the drop-in model + loss + torch.optim."""
    from yelprecommendation_b200.loss import BPRLoss
    from yelprecommendation_b200.models.mf import MatrixFactorization
    g = load_npz("mf_small.npz")
    model = MatrixFactorization(cfg(), int(g["num_users"]), int(g["num_items"])).cuda()
    assert sorted(model.state_dict().keys()) == ["item_embedding.weight", "user_embedding.weight"]
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(g["mf_U0"]))
        model.item_embedding.weight.copy_(torch.from_numpy(g["mf_V0"]))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=0.0)
    loss_fn, losses = BPRLoss(), []
    for data in batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 256, limit=6):
        u, p, n = data["user_id"].cuda(), data["pos_item"].cuda(), data["neg_item"].cuda()
        pos_pred, neg_pred = model(u, p), model(u, n)
        opt.zero_grad()
        loss = loss_fn(pos_pred, neg_pred)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert rel_err(losses, g["mf_2_losses"]) < RTOL
    assert rel_err(model.user_embedding.weight.detach().cpu().numpy(), g["mf_2_U"]) < 2e-5


def _touches(batches, nU, nI):
    """how many gradient rows land in every table row over the batches (register-path SGD: one RED, i.e. one rounding, each)"""
    tu, tv = np.zeros(nU), np.zeros(nI)
    for x in batches:
        np.add.at(tu, x["user_id"].numpy(), 1)
        np.add.at(tv, x["pos_item"].numpy(), 1)
        np.add.at(tv, x["neg_item"].numpy(), 1)
    return tu, tv


def _same_losses(got, want):
    """Step losses: double sums in two different (fixed) orders rounded to fp32 — equal up to the last bit."""
    np.testing.assert_array_max_ulp(np.asarray(got, np.float32), np.asarray(want, np.float32), maxulp=1)


@pytest.mark.parametrize("name,wd,det", [("sgd", 0.0, False), ("sgd", 0.0, True), ("adam", 0.0, False), ("adamw", 1e-2, False)])
def test_full_size_yelp_shape_vs_oracle(name, wd, det):
    """BASELINE config 1 shape (31,668 x 38,048, d=64, B=2048): 12 fused steps vs the C oracle.
    Ordered path (Adam / AdamW / weight decay, or cfg.deterministic): BIT-EXACT tables — the kernel and the oracle sum
    duplicate rows in the reference's own order and share one definition of exp. Register path (plain SGD, default):
    duplicates land as REDs in arrival order -> 1e-5 on the tables and 1e-4 on the UPDATES (p - p0)."""
    from yelprecommendation_b200.trainers import MFTrainer
    rng = np.random.default_rng(1)
    nU, nI, B, steps = 31_668, 38_048, 2048, 12
    tr = MFTrainer(cfg(optimizer=name, lr=1e-2, weight_decay=wd, batch_size=B, deterministic=det), nI, nU)
    U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy()
    V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
    u = rng.integers(0, nU, B * steps)
    u[:64] = 7                                                    # heavy duplicate rows inside one batch
    p, n = rng.integers(0, nI, B * steps), rng.integers(0, nI, B * steps)
    p[100:140] = 11                                               # a hot item, and the same item as positive and negative
    n[140:160] = 11
    b = batches_from(u, p, n, B)
    total = tr.train(b)
    orc = cport.MFTrainerOracle(U0, V0, name, 1e-2, wd)
    ototal, osteps = orc.train([{k: v.numpy() for k, v in x.items()} for x in b])
    assert isclose(total, ototal, rel_tol=1e-6)
    if name != "sgd" or det:
        _same_losses(tr.last_step_losses.cpu().numpy(), osteps)
    else:                              # register path: expf / log1pf on the fp32 pipe (<= 1 ulp per term)
        assert rel_err(tr.last_step_losses.cpu().numpy(), osteps) < 1e-6
    Ug, Vg = (w.detach().cpu().numpy() for w in (tr.model.user_embedding.weight, tr.model.item_embedding.weight))
    if name != "sgd" or det:
        assert np.array_equal(Ug, orc.U) and np.array_equal(Vg, orc.V)
    else:
        assert rel_err(Ug, orc.U) < RTOL and rel_err(Vg, orc.V) < RTOL
        tu_, tv_ = _touches(b, nU, nI)
        assert_update_close(Ug, orc.U, U0, "U", touches=tu_)
        assert_update_close(Vg, orc.V, V0, "V", touches=tv_)
    # linearity property of the SGD step: untouched rows are bit-identical to the initial table
    if name == "sgd":
        touched = np.zeros(nU, bool)
        touched[u] = True
        assert np.array_equal(Ug[~touched], U0[~touched])
    assert all(int(torch.count_nonzero(tr._scratch[k]).item()) == 0 for k in ("flagU", "flagV"))


@pytest.mark.parametrize("name,wd", [("adam", 0.0), ("sgd", 1e-3)])
def test_two_runs_are_bit_identical(name, wd):
    """No floating-point atomics on the dense-semantics path: the same inputs give the same bits, run after run, whether
    the batches go through one launch or one launch per batch."""
    from yelprecommendation_b200.trainers import MFTrainer
    rng = np.random.default_rng(3)
    nU, nI, B, steps = 5000, 4000, 2048, 8
    u, p, n = rng.integers(0, 50, B * steps), rng.integers(0, 40, B * steps), rng.integers(0, nI, B * steps)   # many duplicates
    b = batches_from(u, p, n, B)
    outs = []
    for spl in (64, 64, 1):
        torch.manual_seed(0)
        tr = MFTrainer(cfg(optimizer=name, lr=1e-2, weight_decay=wd, batch_size=B, steps_per_launch=spl), nI, nU)
        total = tr.train(b)
        outs.append((total, tr.last_step_losses.cpu().numpy().copy(), tr.model.user_embedding.weight.detach().cpu().numpy().copy(),
                     tr.model.item_embedding.weight.detach().cpu().numpy().copy()))
    for o in outs[1:]:
        assert o[0] == outs[0][0] and np.array_equal(o[1], outs[0][1])
        assert np.array_equal(o[2], outs[0][2]) and np.array_equal(o[3], outs[0][3])


@pytest.mark.parametrize("d", [32, 128, 256, 512, 1024])
@pytest.mark.parametrize("name,wd", [("sgd", 0.0), ("sgd", 1e-3), ("adam", 0.0)])
def test_other_embedding_widths_vs_oracle(d, name, wd):
    """embed_size 32 ... 1,024 (the values of the reference's mf_sweep_config.yaml; 1 ... 32 floats per lane):
    register-resident SGD (up to 256), the ordered path for everything else, 6 steps of 1,024 triples with duplicates,
    against the C oracle (bit-exact on the ordered path); then validate and a full evaluation at that width."""
    from yelprecommendation_b200.trainers import MFTrainer
    rng = np.random.default_rng(d)
    nU, nI, B, steps = 3000, 2500, 1024, 6
    tr = MFTrainer(cfg(optimizer=name, lr=1e-2, weight_decay=wd, batch_size=B, embed_size=d), nI, nU)
    U0 = tr.model.user_embedding.weight.detach().cpu().numpy().copy()
    V0 = tr.model.item_embedding.weight.detach().cpu().numpy().copy()
    u, p, n = rng.integers(0, nU, B * steps), rng.integers(0, nI, B * steps), rng.integers(0, nI, B * steps)
    u[:32] = 5
    b = batches_from(u, p, n, B)
    total = tr.train(b)
    orc = cport.MFTrainerOracle(U0, V0, name, 1e-2, wd)
    ototal, osteps = orc.train([{k: v.numpy() for k, v in x.items()} for x in b])
    assert isclose(total, ototal, rel_tol=1e-6)
    register_path = name == "sgd" and wd == 0.0 and d <= 256
    if register_path:
        assert rel_err(tr.last_step_losses.cpu().numpy(), osteps) < 1e-6
    else:
        _same_losses(tr.last_step_losses.cpu().numpy(), osteps)
    Ug, Vg = (w.detach().cpu().numpy() for w in (tr.model.user_embedding.weight, tr.model.item_embedding.weight))
    if register_path:
        assert rel_err(Ug, orc.U) < RTOL and rel_err(Vg, orc.V) < RTOL
        tu_, tv_ = _touches(b, nU, nI)
        assert_update_close(Ug, orc.U, U0, "U", touches=tu_)
        assert_update_close(Vg, orc.V, V0, "V", touches=tv_)
    else:
        assert np.array_equal(Ug, orc.U) and np.array_equal(Vg, orc.V)
    # the kernel's invariant: the row flags are all-zero again after every call (a row the sweep missed — the d = 32
    # grid-barrier bug, notes/README.md — would leave its flag behind)
    sc = tr._scratch
    assert all(int(torch.count_nonzero(sc[k]).item()) == 0 for k in ("flagU", "flagV"))
    if wd == 0.0:
        return
    # validate + evaluate at this width (ids bit-exact against the oracle on the trainer's own tables)
    assert isclose(tr.validate(b[:2]), orc_validate(Ug, Vg, b[:2]), rel_tol=RTOL)
    from yelprecommendation_b200.data.graph import build_eval_csr
    uid = rng.integers(0, nU, 150)
    pos = [rng.permutation(nI)[: int(rng.integers(1, 20))].tolist() for _ in uid]
    mask = [rng.permutation(nI)[: int(rng.integers(0, 40))].tolist() for _ in uid]
    ec = build_eval_csr(uid, pos, mask, nI)
    got = tr.evaluate(ec)
    otopk, _, _, osums = cport.eval_topk_metrics(Ug, Vg, ec.eval_uid, ec.mask_ptr, ec.mask_idx, ec.act_ptr, ec.act_idx, 10)
    assert np.array_equal(tr.last_topk.cpu().numpy(), otopk)
    assert np.allclose(got, cport.metrics_from_sums(osums, ec.n_eval), rtol=1e-12)


def orc_validate(U, V, batches):
    """Sum of batch-mean BPR losses on fixed tables (MFTrainer.validate, mf_trainer.py:118-132), float64 on the host."""
    total = 0.0
    for x in batches:
        u, p, n = (x[k].numpy() for k in ("user_id", "pos_item", "neg_item"))
        d = (U[u].astype(np.float64) * (V[p].astype(np.float64) - V[n].astype(np.float64))).sum(1)
        total += float(np.mean(np.logaddexp(0.0, -d)))
    return total
