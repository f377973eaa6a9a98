"""CDAE (BASELINE config 4): the torch-CPU port is pinned to the real reference (CPU test); the CUDA path is checked
against the reference fixture and the port (GPU tests)."""
from math import isclose

import numpy as np
import pytest
import torch

from oracle import torch_port as tp
from util import RTOL, assert_adam_close, cfg, load_npz, rel_err

KEYS = ["hidden_layer.weight", "hidden_layer.bias", "user_nodes.weight", "output_layer.weight", "output_layer.bias"]


def _fixture():
    g = load_npz("cdae_small.npz")
    nU, nI, B = int(g["nU"]), int(g["nI"]), int(g["B"])
    t = torch.from_numpy
    users = np.arange(nU)

    def batches(masks, n=None):
        out = []
        for s in range(0, nU, B):
            sl = slice(s, min(s + B, nU))
            out.append({k: (t(users[sl].copy()) if v is None else t(v[sl].copy())) for k, v in masks.items()})
        return out[:n] if n else out

    tb = batches({"user_id": None, "input_mask": g["train_mask"], "negative_mask": g["neg_train"]}, 3)
    vb = batches({"user_id": None, "input_mask": g["train_mask"], "valid_mask": g["valid_mask"], "negative_mask": g["neg_valid"]})
    eb = batches({"user_id": None, "input_mask": g["train_mask"] + g["valid_mask"], "test_mask": g["test_mask"]})
    keeps = [t(k.copy()) for k in g["keeps"]]
    init = {k: t(g["init_" + k].copy()) for k in KEYS}
    return g, nU, nI, B, tb, vb, eb, keeps, init


@pytest.mark.parametrize("ci", [0, 1])
def test_port_matches_reference(ci):
    g, nU, nI, B, tb, vb, eb, keeps, init = _fixture()
    port = tp.CDAEPort(init, str(g[f"c{ci}_name"]), float(g[f"c{ci}_lr"]))
    if ci == 0:
        pred = port.forward(tb[0]["user_id"], tb[0]["input_mask"]).detach().numpy()
        assert rel_err(pred, g["pred_eval0"]) < 1e-6
        assert np.allclose(port.validate(vb), g["valid0"], rtol=1e-6)
    _, steps = port.train(tb, keeps)
    assert rel_err(steps, g[f"c{ci}_losses"]) < 1e-6
    for k, v in port.state_dict().items():
        assert rel_err(v.detach().numpy(), g[f"c{ci}_final_" + k]) < 1e-5, k
    if ci == 0:
        assert np.allclose(port.validate(vb), g["valid_after"], rtol=1e-5)
        assert np.allclose(port.evaluate(eb), g["test_after"], rtol=1e-6)


def _trainer(g, nU, nI, init, name, lr):
    from yelprecommendation_b200.trainers import CDAETrainer
    tr = CDAETrainer(cfg(optimizer=name, lr=lr, hidden_size=64, corruption_level=0.6, hidden_activation="sigmoid",
                         output_activation="sigmoid", negative_sampling=True, loss_name="bce"), nI, nU)
    assert sorted(tr.model.state_dict().keys()) == sorted(KEYS)
    tr.model.load_state_dict(init)
    return tr


@pytest.mark.gpu
def test_forward_and_loss_vs_reference():
    from yelprecommendation_b200.loss import NSBCELoss
    g, nU, nI, B, tb, vb, eb, keeps, init = _fixture()
    tr = _trainer(g, nU, nI, init, "adam", 1e-2)
    tr.model.eval()
    pred = tr.model(tb[0]["user_id"], tb[0]["input_mask"])
    assert rel_err(pred.cpu().numpy(), g["pred_eval0"]) < RTOL
    ref = tp.nsbce_loss(torch.from_numpy(g["pred_eval0"]), tb[0]["input_mask"], tb[0]["negative_mask"]).item()
    got = NSBCELoss()(pred, tb[0]["input_mask"].cuda(), tb[0]["negative_mask"].cuda()).item()
    assert isclose(got, ref, rel_tol=RTOL)
    # training mode with a supplied dropout multiplier
    tr.model.train()
    port = tp.CDAEPort(init)
    want = port.forward(tb[0]["user_id"], tb[0]["input_mask"], keeps[0]).detach().numpy()
    got = tr.model(tb[0]["user_id"], tb[0]["input_mask"], keep=keeps[0].cuda()).cpu().numpy()
    assert rel_err(got, want) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("ci", [0, 1])
def test_train_validate_evaluate_vs_reference(ci):
    g, nU, nI, B, tb, vb, eb, keeps, init = _fixture()
    tr = _trainer(g, nU, nI, init, str(g[f"c{ci}_name"]), float(g[f"c{ci}_lr"]))
    if ci == 0:
        v0 = tr.validate(vb)
        assert isclose(v0[0], float(g["valid0"][0]), rel_tol=RTOL)
        assert np.allclose(v0[1:], g["valid0"][1:], rtol=0.1, atol=2e-3)      # ranking metrics: ties aside
    total = tr.train(tb, keeps=keeps)
    assert rel_err(tr.last_step_losses.cpu().numpy(), g[f"c{ci}_losses"]) < RTOL
    assert isclose(total, float(np.sum(g[f"c{ci}_losses"])), rel_tol=RTOL)
    for k, v in tr.model.state_dict().items():
        assert rel_err(v.cpu().numpy(), g[f"c{ci}_final_" + k]) < 2e-5, k
    b = tr._bufs
    assert all(int(torch.count_nonzero(t).item()) == 0 for t in b["grads"])            # grads re-zeroed
    if ci == 0:
        va = tr.validate(vb)
        assert isclose(va[0], float(g["valid_after"][0]), rel_tol=2e-5)
        te = tr.evaluate(eb)
        # exact ranking check against the port's per-row top-10 (ties in the port follow NumPy)
        port = tp.CDAEPort({k: v.detach().cpu() for k, v in tr.model.state_dict().items()})
        (pm, ppred) = port._rank(eb, "test_mask")
        same = sum(np.array_equal(tr.last_topk[r].cpu().numpy(), ppred[r]) for r in range(nU))
        assert same >= nU - max(2, nU // 20)
        assert np.allclose(te, pm, rtol=0.08, atol=2e-3)


@pytest.mark.gpu
def test_default_dropout_draw_and_bad_ids():
    g, nU, nI, B, tb, vb, eb, keeps, init = _fixture()
    tr = _trainer(g, nU, nI, init, "adam", 1e-3)
    loss = tr.train(tb)                                    # masks drawn on the device
    assert np.isfinite(loss) and loss > 0
    k = tr.model.draw_keep(torch.ones(64, nI, device="cuda"))
    vals = torch.unique(k).cpu().numpy()
    assert set(np.round(vals, 4)) <= {0.0, 2.5} and 0.3 < float((k > 0).float().mean()) < 0.5
    bad = dict(tb[0])
    bad["user_id"] = bad["user_id"].clone()
    bad["user_id"][0] = nU
    with pytest.raises(IndexError):
        tr.train([bad])


# ---- hidden_size / hidden_activation values of the reference's cdae_sweep_config.yaml ---------------------------------
def _width_case(wi):
    g, nU, nI, B, tb, vb, eb, keeps, _ = _fixture()
    w = load_npz("cdae_widths.npz")
    init = {k: torch.from_numpy(w[f"w{wi}_init_" + k].copy()) for k in KEYS}
    return w, nU, nI, B, tb, vb, eb, keeps, init, int(w[f"w{wi}_h"]), str(w[f"w{wi}_act"]), str(w[f"w{wi}_name"]), float(w[f"w{wi}_lr"])


@pytest.mark.parametrize("wi", [0, 1, 2])
def test_port_matches_reference_other_widths(wi):
    """The port with hidden_size 32 / 128 / 256 and hidden_activation identity / sigmoid against the REAL reference
    (fixture cdae_widths.npz, tests/golden/make_golden.py cdae_widths)."""
    w, nU, nI, B, tb, vb, eb, keeps, init, h, act, name, lr = _width_case(wi)
    port = tp.CDAEPort(init, name, lr, hidden_activation=act)
    pred = port.forward(tb[0]["user_id"], tb[0]["input_mask"]).detach().numpy()
    assert rel_err(pred, w[f"w{wi}_pred_eval0"]) < 1e-6
    _, steps = port.train(tb, keeps)
    assert rel_err(steps, w[f"w{wi}_losses"]) < 1e-6
    for k, v in port.state_dict().items():
        assert rel_err(v.detach().numpy(), w[f"w{wi}_final_" + k]) < 1e-5, k
    assert np.allclose(port.validate(vb), w[f"w{wi}_valid_after"], rtol=1e-5)
    assert np.allclose(port.evaluate(eb), w[f"w{wi}_test_after"], rtol=1e-6)


def _width_trainer(nU, nI, init, h, act, name, lr):
    from yelprecommendation_b200.trainers import CDAETrainer
    tr = CDAETrainer(cfg(optimizer=name, lr=lr, hidden_size=h, corruption_level=0.6, hidden_activation=act,
                         output_activation="sigmoid", negative_sampling=True, loss_name="bce"), nI, nU)
    tr.model.load_state_dict(init)
    return tr


@pytest.mark.gpu
@pytest.mark.parametrize("wi", [0, 1, 2])
def test_train_validate_evaluate_other_widths_vs_reference(wi):
    w, nU, nI, B, tb, vb, eb, keeps, init, h, act, name, lr = _width_case(wi)
    tr = _width_trainer(nU, nI, init, h, act, name, lr)
    tr.model.eval()
    pred = tr.model(tb[0]["user_id"], tb[0]["input_mask"])
    assert rel_err(pred.cpu().numpy(), w[f"w{wi}_pred_eval0"]) < RTOL
    total = tr.train(tb, keeps=keeps)
    assert rel_err(tr.last_step_losses.cpu().numpy(), w[f"w{wi}_losses"]) < RTOL
    assert isclose(total, float(np.sum(w[f"w{wi}_losses"])), rel_tol=RTOL)
    for k, v in tr.model.state_dict().items():
        if name == "sgd":
            assert rel_err(v.cpu().numpy(), w[f"w{wi}_final_" + k]) < 2e-5, k
        else:       # Adam: norm-wise 1e-5 with bounded eps-amplified outliers (util.assert_adam_close)
            assert_adam_close(v.cpu().numpy(), w[f"w{wi}_final_" + k], k, touched=v.numel() * 3)
    va = tr.validate(vb)
    assert isclose(va[0], float(w[f"w{wi}_valid_after"][0]), rel_tol=2e-5)
    te = tr.evaluate(eb)
    port = tp.CDAEPort({k: v.detach().cpu() for k, v in tr.model.state_dict().items()}, hidden_activation=act)
    (pm, ppred) = port._rank(eb, "test_mask")
    same = sum(np.array_equal(tr.last_topk[r].cpu().numpy(), ppred[r]) for r in range(nU))
    assert same >= nU - max(2, nU // 20)
    assert np.allclose(te, pm, rtol=0.08, atol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("h,act", [(512, "sigmoid"), (1024, "identity")])
def test_widest_hidden_sizes_vs_port(h, act):
    """hidden_size 512 / 1,024 (16 / 32 floats per lane; the user tile of the evaluation spills to HBM at 1,056 columns)."""
    g, nU, nI, B, tb, vb, eb, keeps, _ = _fixture()
    gen = torch.Generator().manual_seed(h)
    r = lambda *s, a=1.0: (torch.rand(*s, generator=gen) * 2 - 1) * a
    init = {"hidden_layer.weight": r(h, nI, a=(6.0 / (h + nI)) ** 0.5), "hidden_layer.bias": torch.rand(h, generator=gen),
            "user_nodes.weight": torch.rand(nU, h, generator=gen) * 0.1, "output_layer.weight": r(nI, h, a=(6.0 / (h + nI)) ** 0.5),
            "output_layer.bias": torch.rand(nI, generator=gen)}
    tr = _width_trainer(nU, nI, init, h, act, "adam", 1e-3)
    port = tp.CDAEPort(init, "adam", 1e-3, hidden_activation=act)
    total = tr.train(tb, keeps=keeps)
    ptotal, psteps = port.train(tb, keeps)
    assert rel_err(tr.last_step_losses.cpu().numpy(), psteps) < RTOL
    for k, v in tr.model.state_dict().items():       # Adam: norm-wise 1e-5, rare eps-amplified elements bounded (util.assert_adam_close)
        assert_adam_close(v.cpu().numpy(), port.state_dict()[k].detach().numpy(), k, touched=v.numel() * 3)
    va, pv = tr.validate(vb), port.validate(vb)
    assert isclose(va[0], pv[0], rel_tol=2e-5)
    te = tr.evaluate(eb)
    (pm, ppred) = tp.CDAEPort({k: v.detach().cpu() for k, v in tr.model.state_dict().items()}, hidden_activation=act)._rank(eb, "test_mask")
    same = sum(np.array_equal(tr.last_topk[r].cpu().numpy(), ppred[r]) for r in range(nU))
    assert same >= nU - max(2, nU // 20)


@pytest.mark.gpu
@pytest.mark.parametrize("name,lr", [("sgd", 0.05), ("adam", 1e-3)])
def test_index_list_batches_equal_dense_batches(name, lr):
    """yr_cdae_step_idx: the same batches shipped as index lists (data/cdae_sparse.py: a few hundred bytes per user instead of
    two dense [B x num_items] masks) give the same losses and parameters as the dense form, with the same dropout values."""
    from yelprecommendation_b200.data.cdae_sparse import sparse_cdae_batch
    g, nU, nI, B, tb, vb, eb, keeps, init = _fixture()
    dense = _trainer(g, nU, nI, init, name, lr)
    ld = dense.train(tb, keeps)
    sparse = _trainer(g, nU, nI, init, name, lr)
    sb = [sparse_cdae_batch(b) for b in tb]
    kv = [k[b["input_mask"] != 0] for k, b in zip(keeps, tb)]           # the dropout multiplier of every listed input
    assert all(s["input_idx"].numel() == v.numel() for s, v in zip(sb, kv))
    ls = sparse.train(sb, kv)
    assert np.isclose(ls, ld, rtol=1e-6)
    assert rel_err(sparse.last_step_losses.cpu().numpy(), dense.last_step_losses.cpu().numpy()) < 1e-6
    for (k, a), (_, b_) in zip(sparse.model.state_dict().items(), dense.model.state_dict().items()):
        if name == "sgd":
            assert rel_err(a.cpu().numpy(), b_.cpu().numpy()) < 1e-6, k
        else:
            assert_adam_close(a.cpu().numpy(), b_.cpu().numpy(), k, touched=a.numel() * 3)
    # device-drawn dropout on the lists runs, and a bad item id raises like nn.Embedding would
    assert np.isfinite(sparse.train(sb)) 
    bad = dict(sb[0])
    bad["input_idx"] = bad["input_idx"].clone()
    bad["input_idx"][0] = nI
    with pytest.raises(IndexError):
        sparse.train([bad])
