"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference (twndus/YelpRecommendation, mounted at
/root/reference) on small seeded inputs. Run here (the build container); the fixtures are committed because the
GPU box has no /root/reference.

    python tests/golden/make_golden.py

What is imported from the reference (nothing is copied): metric.py, models/mf.py, models/ngcf.py, loss.py,
trainers/mf_trainer.py, trainers/ngcf_trainer.py, data/datasets/mf_data_pipeline.py, ngcf_data_pipeline.py.
Shims: `omegaconf` (type annotation only, trainers/base_trainer.py:10) and, for the Laplacian only, a
`Tensor.to('cuda')` no-op because ngcf_data_pipeline.py:38-39 hard-codes the device.
"""
import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("YR_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

oc, dc = types.ModuleType("omegaconf"), types.ModuleType("omegaconf.dictconfig")


class DictConfig(dict):
    pass


dc.DictConfig = oc.DictConfig = DictConfig
oc.dictconfig = dc
sys.modules["omegaconf"], sys.modules["omegaconf.dictconfig"] = oc, dc

import metric as ref_metric                                   # noqa: E402
from trainers.mf_trainer import MFTrainer as RefMFTrainer     # noqa: E402
from trainers.ngcf_trainer import NGCFTrainer as RefNGCFTrainer  # noqa: E402
from data.datasets.mf_data_pipeline import MFDataPipeline     # noqa: E402
from data.datasets.ngcf_data_pipeline import NGCFDataPipeline  # noqa: E402

from yelprecommendation_b200.data import synthetic as syn     # noqa: E402

from loguru import logger                                     # noqa: E402
logger.remove()

TMP = tempfile.mkdtemp()


def cfg(**kw):
    base = dict(device="cpu", model_dir=TMP, embed_size=64, optimizer="sgd", lr=1e-2, weight_decay=0.0, top_n=10,
                wandb=False, num_orders=3, epochs=1, patience=1, best_metric="loss", loss_name="bpr", seed=42,
                batch_size=256)
    base.update(kw)
    return SimpleNamespace(**base)


def frame(inter):
    return pd.DataFrame({"user_id": inter.user, "business_id": inter.item, "rating": inter.rating})


# ------------------------------------------------------------------------------------------------
def golden_metrics():
    rng = np.random.default_rng(11)
    cases = []

    def add(actual, predicted, k):
        cases.append(dict(actual=[list(map(int, a)) for a in actual], predicted=[list(map(int, p)) for p in predicted],
                          k=k, precision=ref_metric.precision_at_k(actual, predicted, k),
                          recall=ref_metric.recall_at_k(actual, predicted, k),
                          map=ref_metric.map_at_k(actual, predicted, k), ndcg=ref_metric.ndcg_at_k(actual, predicted, k)))

    # the reference's own test vectors (test/test_metric.py:9-47)
    for k in range(1, 6):
        add([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]], [[1, 6, 7, 11, 12], [6, 7, 14, 16, 20]], k)
    # order dependence of AP (Q7), NDCG window (Q8), empty users (Q6)
    for A in ([1, 5, 9], [9, 5, 1], [5, 9, 1]):
        add([A], [[5, 9, 1, 7, 8, 11, 12, 13, 14, 15]], 10)
    add([[1, 2]], [[7, 8, 1, 2, 3, 4, 5, 6, 9, 10]], 10)
    add([[1, 2]], [[1, 8, 7, 2, 3, 4, 5, 6, 9, 10]], 10)
    add([[1], []], [[1, 2], [3, 4]], 2)
    # random: shuffled actual, len(actual) < k, duplicates in actual, empty rows
    for t in range(40):
        n = int(rng.integers(1, 7))
        k = int(rng.integers(1, 13))
        actual, predicted = [], []
        for _ in range(n):
            la = int(rng.integers(0, 15))
            a = rng.integers(0, 30, size=la).tolist() if t % 3 == 0 else rng.permutation(30)[:la].tolist()
            actual.append(a)
            predicted.append(rng.permutation(30)[:max(k, 12)].tolist())
        if all(len(set(a)) == 0 for a in actual):
            actual[0] = [3]
        add(actual, predicted, k)
    with open(os.path.join(HERE, "metric_cases.json"), "w") as f:
        json.dump(cases, f)
    print("metric cases:", len(cases))


# ------------------------------------------------------------------------------------------------
def golden_split_and_mf():
    inter = syn.make_interactions(num_users=160, num_items=240, nnz=3200, seed=5, n_clusters=4)
    df = frame(inter)
    pipe = MFDataPipeline(cfg())
    pipe.num_users, pipe.num_items = inter.num_users, inter.num_items
    train_data, valid_data, valid_eval, test_eval = pipe.split(df)
    out = dict(user=inter.user, item=inter.item, rating=inter.rating, num_users=inter.num_users,
               num_items=inter.num_items)

    def pack(prefix, ev):
        out[f"{prefix}_uid"] = ev.index.to_numpy().astype(np.int64)
        for col in ("pos_items", "mask_items"):
            lens = np.array([len(x) for x in ev[col]], dtype=np.int64)
            out[f"{prefix}_{col}_ptr"] = np.concatenate([[0], np.cumsum(lens)])
            out[f"{prefix}_{col}"] = np.concatenate([np.asarray(x, dtype=np.int64) for x in ev[col]])

    pack("valid_eval", valid_eval)
    pack("test_eval", test_eval)
    out["train_user"] = train_data["user_id"].to_numpy().astype(np.int64)
    out["train_item"] = train_data["business_id"].to_numpy().astype(np.int64)
    out["valid_user"] = valid_data["user_id"].to_numpy().astype(np.int64)
    out["valid_item"] = valid_data["business_id"].to_numpy().astype(np.int64)

    # pre-sampled triples (ours — the reference samples inside Dataset.__getitem__ from the global RNG)
    split = syn.split_per_user(inter, seed=42)
    tu, tp, tn = syn.sample_triples(split, inter.num_items, seed=42)
    vu, vp, vn = syn.sample_triples(split, inter.num_items, seed=43, which="valid", reject="train+valid")
    out.update(tri_u=tu, tri_p=tp, tri_n=tn, vtri_u=vu, vtri_p=vp, vtri_n=vn)
    B = 256
    batches = syn.to_batches(tu, tp, tn, B)[:6]
    vbatches = syn.to_batches(vu, vp, vn, B)

    configs = [("sgd", 1e-2, 0.0), ("sgd", 1e-2, 1e-2), ("adam", 1e-2, 0.0), ("adam", 1e-3, 1e-3), ("adamw", 1e-2, 1e-2)]
    for ci, (name, lr, wd) in enumerate(configs):
        torch.manual_seed(42)
        tr = RefMFTrainer(cfg(optimizer=name, lr=lr, weight_decay=wd), inter.num_items, inter.num_users)
        if ci == 0:
            out["mf_U0"] = tr.model.user_embedding.weight.detach().numpy().copy()
            out["mf_V0"] = tr.model.item_embedding.weight.detach().numpy().copy()
            out["mf_valid0"] = tr.validate(vbatches)
            u = torch.from_numpy(tu[:300])
            out["mf_score0"] = tr.model(u, torch.from_numpy(tp[:300])).detach().numpy()
        # per-step losses need a second pass structure: call train() batch by batch
        losses = [tr.train([b]) for b in batches]
        out[f"mf_{ci}_cfg"] = np.array([lr, wd])
        out[f"mf_{ci}_name"] = name
        out[f"mf_{ci}_losses"] = np.array(losses)
        out[f"mf_{ci}_U"] = tr.model.user_embedding.weight.detach().numpy().copy()
        out[f"mf_{ci}_V"] = tr.model.item_embedding.weight.detach().numpy().copy()
        if ci == 2:
            out["mf_valid_after"] = tr.validate(vbatches)
    out["n_mf_cfg"] = len(configs)

    # evaluation: planted ("trained-ish") tables so that top-10 gaps and metrics are realistic
    Up, Vp = syn.planted_embeddings(inter, d=64, seed=7)
    tr = RefMFTrainer(cfg(), inter.num_items, inter.num_users)
    with torch.no_grad():
        tr.model.user_embedding.weight.copy_(torch.from_numpy(Up))
        tr.model.item_embedding.weight.copy_(torch.from_numpy(Vp))
    out["eval_U"], out["eval_V"] = Up, Vp
    for prefix, ev in (("valid_eval", valid_eval), ("test_eval", test_eval)):
        out[f"{prefix}_metrics"] = np.array(tr.evaluate(ev, "valid"))
        preds = []
        items = torch.arange(inter.num_items)
        for uid, row in ev.iterrows():
            pred = tr.model(torch.tensor([uid] * inter.num_items), items)
            preds.append(tr._generate_top_k_recommendation(pred, row["mask_items"]))
        out[f"{prefix}_topk"] = np.stack(preds).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "mf_small.npz"), **out)
    print("mf_small: valid users", len(valid_eval), "metrics", out["valid_eval_metrics"])
    return inter, split


# ------------------------------------------------------------------------------------------------
def golden_ngcf():
    out = {}
    for tag, stars in (("bin", False), ("star", True)):
        inter = syn.make_interactions(num_users=96, num_items=128, nnz=1500, seed=9, n_clusters=4, star_ratings=stars)
        df = frame(inter)
        pipe = NGCFDataPipeline(cfg())
        pipe.num_users, pipe.num_items = inter.num_users, inter.num_items
        orig_to = torch.Tensor.to

        def to_shim(self, *a, **k):   # ngcf_data_pipeline.py:38-39 hard-codes 'cuda'
            if a and a[0] == "cuda":
                return self
            return orig_to(self, *a, **k)

        torch.Tensor.to = to_shim
        try:
            pipe._set_laplacian_matrix(df)
        finally:
            torch.Tensor.to = orig_to
        L = pipe.laplacian_matrix.coalesce()
        out[f"{tag}_user"], out[f"{tag}_item"], out[f"{tag}_rating"] = inter.user, inter.item, inter.rating
        out[f"{tag}_L_idx"] = L.indices().numpy()
        out[f"{tag}_L_val"] = L.values().numpy()
        out[f"{tag}_nU"], out[f"{tag}_nI"] = inter.num_users, inter.num_items
        if tag == "star":
            continue
        split = syn.split_per_user(inter, seed=42)
        tu, tp, tn = syn.sample_triples(split, inter.num_items, seed=42)
        out.update(tri_u=tu, tri_p=tp, tri_n=tn)
        B = 128
        batches = syn.to_batches(tu, tp, tn, B)[:3]
        for ci, (name, lr, wd) in enumerate([("sgd", 1e-2, 0.0), ("adam", 1e-3, 0.0), ("adamw", 1e-3, 1e-2)]):
            torch.manual_seed(42)
            tr = RefNGCFTrainer(cfg(optimizer=name, lr=lr, weight_decay=wd, num_orders=3), inter.num_items,
                                inter.num_users, L)
            if ci == 0:
                sd = {k: v.detach().numpy().copy() for k, v in tr.model.state_dict().items()}
                for k, v in sd.items():
                    out["ngcf_init_" + k] = v
                with torch.no_grad():
                    e = tr.model.embedding.weight
                    for l, (w1, w2) in enumerate(zip(tr.model.W1, tr.model.W2)):
                        e = tr.model.embedding_propagation(e, w1, w2, L)   # verbatim, torch.eye and all
                        out[f"ngcf_layer{l + 1}"] = e.numpy().copy()
                    b0 = batches[0]
                    p, n = tr.model.bpr_forward(b0["user_id"], b0["pos_item"], b0["neg_item"], L)
                    out["ngcf_pos0"], out["ngcf_neg0"] = p.numpy().copy(), n.numpy().copy()
                    out["ngcf_valid0"] = tr.validate(batches)
                    # per-user forward exactly as trainers/ngcf_trainer.py:144 does it
                    items = torch.arange(inter.num_items)
                    users = [0, 5, 17, 40]
                    out["ngcf_eval_users"] = np.array(users)
                    out["ngcf_eval_scores"] = np.stack([
                        tr.model(torch.tensor([u] * inter.num_items), items, L).numpy() for u in users])
                # gradients of the first step (autograd), for the backward kernels
                b0 = batches[0]
                p, n = tr.model.bpr_forward(b0["user_id"], b0["pos_item"], b0["neg_item"], L)
                loss = tr.loss(p, n)
                tr.optimizer.zero_grad()
                loss.backward()
                for k, prm in tr.model.named_parameters():
                    out["ngcf_grad_" + k] = prm.grad.detach().numpy().copy()
                tr.optimizer.zero_grad()
            losses = [tr.train([b]) for b in batches]
            out[f"ngcf_{ci}_name"] = name
            out[f"ngcf_{ci}_cfg"] = np.array([lr, wd])
            out[f"ngcf_{ci}_losses"] = np.array(losses)
            for k, v in tr.model.state_dict().items():
                out[f"ngcf_{ci}_final_" + k] = v.detach().numpy().copy()
        out["n_ngcf_cfg"] = 3
    np.savez_compressed(os.path.join(HERE, "ngcf_small.npz"), **out)
    print("ngcf_small: losses", out["ngcf_0_losses"])


# ------------------------------------------------------------------------------------------------
class _SuppliedDropout(torch.nn.Module):
    """Stands in for model.dropout_layer so that the fixture records the exact nn.Dropout multipliers used."""

    def __init__(self, keeps):
        super().__init__()
        self.keeps, self.i = keeps, 0

    def forward(self, x):
        if not self.training:
            return x
        k = self.keeps[self.i]
        self.i += 1
        return x * k


def _cdae_problem():
    rng = np.random.default_rng(21)
    nU, nI, B, p = 80, 112, 16, 0.6
    inter = syn.make_interactions(num_users=nU, num_items=nI, nnz=1400, seed=6, n_clusters=4)
    dense = np.zeros((nU, nI), np.float32)
    dense[inter.user, inter.item] = 1.0
    # per-user split into train / valid / test masks in the spirit of cdae_data_pipeline.py:18-52
    train_m, valid_m, test_m = np.zeros_like(dense), np.zeros_like(dense), np.zeros_like(dense)
    for u in range(nU):
        h = rng.permutation(np.nonzero(dense[u])[0])
        tr, te = np.split(h, [int(0.8 * len(h))])
        tr, va = np.split(tr, [int(0.75 * len(tr))])
        train_m[u, tr], valid_m[u, va], test_m[u, te] = 1, 1, 1

    def neg_mask(pos):
        out = np.zeros_like(pos)
        for u in range(pos.shape[0]):
            cand = np.nonzero(1 - pos[u])[0]
            out[u, rng.choice(cand, min(len(cand), int(pos[u].sum()) * 5), replace=False)] = 1.0
        return out

    users = np.arange(nU)
    neg_train, neg_valid = neg_mask(train_m), neg_mask(train_m + valid_m)
    n_steps = 3
    keeps = [(rng.random((B, nI)) >= p).astype(np.float32) / np.float32(1 - p) for _ in range(n_steps)]
    t = torch.from_numpy

    def batches(masks, lo=0, hi=None):
        hi = hi or nU
        out = []
        for s in range(lo, hi, B):
            sl = slice(s, min(s + B, hi))
            out.append({k: (t(users[sl].copy()) if k == "user_id" else t(v[sl].copy())) for k, v in masks.items()})
        return out

    out = dict(nU=nU, nI=nI, B=B, train_mask=train_m, valid_mask=valid_m, test_mask=test_m, neg_train=neg_train,
               neg_valid=neg_valid, keeps=np.stack(keeps))
    tb = batches({"user_id": None, "input_mask": train_m, "negative_mask": neg_train})[:n_steps]
    vb = batches({"user_id": None, "input_mask": train_m, "valid_mask": valid_m, "negative_mask": neg_valid})
    eb = batches({"user_id": None, "input_mask": train_m + valid_m, "test_mask": test_m})
    return nU, nI, B, p, users, train_m, keeps, out, tb, vb, eb


def golden_cdae_widths():
    """hidden_size / hidden_activation values of the reference's cdae_sweep_config.yaml on the cdae_small problem
    (same masks, batches and dropout multipliers): init, 3 train-step losses, final state, validate / evaluate."""
    from trainers.cdae_trainer import CDAETrainer as RefCDAETrainer
    nU, nI, B, p, users, train_m, keeps, _, tb, vb, eb = _cdae_problem()
    t = torch.from_numpy
    out = {}
    for ci, (h, act, name, lr) in enumerate([(32, "identity", "adam", 1e-2), (128, "sigmoid", "sgd", 0.5),
                                             (256, "identity", "adam", 1e-3)]):
        torch.manual_seed(42 + ci)
        tr = RefCDAETrainer(cfg(optimizer=name, lr=lr, hidden_size=h, corruption_level=p, hidden_activation=act,
                                output_activation="sigmoid", negative_sampling=True, loss_name="bce"), nI, nU)
        tr.model.dropout_layer = _SuppliedDropout([t(k) for k in keeps])
        for k, v in tr.model.state_dict().items():
            out[f"w{ci}_init_" + k] = v.detach().numpy().copy()
        tr.model.eval()
        out[f"w{ci}_pred_eval0"] = tr.model(t(users[:B].copy()), t(train_m[:B].copy())).detach().numpy()
        losses = [tr.train([b]) for b in tb]
        out[f"w{ci}_h"], out[f"w{ci}_act"], out[f"w{ci}_name"], out[f"w{ci}_lr"] = h, act, name, lr
        out[f"w{ci}_losses"] = np.array(losses)
        for k, v in tr.model.state_dict().items():
            out[f"w{ci}_final_" + k] = v.detach().numpy().copy()
        out[f"w{ci}_valid_after"] = np.array(tr.validate(vb))
        out[f"w{ci}_test_after"] = np.array(tr.evaluate(eb))
        print(f"cdae_widths[{ci}] h={h} act={act}: losses", out[f"w{ci}_losses"], "valid", out[f"w{ci}_valid_after"])
    np.savez_compressed(os.path.join(HERE, "cdae_widths.npz"), **out)


def golden_cdae():
    from trainers.cdae_trainer import CDAETrainer as RefCDAETrainer
    nU, nI, B, p, users, train_m, keeps, out, tb, vb, eb = _cdae_problem()
    t = torch.from_numpy
    for ci, (name, lr) in enumerate([("adam", 1e-2), ("sgd", 0.5)]):
        torch.manual_seed(42)
        tr = RefCDAETrainer(cfg(optimizer=name, lr=lr, hidden_size=64, corruption_level=p, hidden_activation="sigmoid",
                                output_activation="sigmoid", negative_sampling=True, loss_name="bce"), nI, nU)
        tr.model.dropout_layer = _SuppliedDropout([t(k) for k in keeps])
        if ci == 0:
            for k, v in tr.model.state_dict().items():
                out["init_" + k] = v.detach().numpy().copy()
            tr.model.eval()
            out["pred_eval0"] = tr.model(t(users[:B].copy()), t(train_m[:B].copy())).detach().numpy()
            out["valid0"] = np.array(tr.validate(vb))
        losses = [tr.train([b]) for b in tb]
        out[f"c{ci}_name"], out[f"c{ci}_lr"], out[f"c{ci}_losses"] = name, lr, np.array(losses)
        for k, v in tr.model.state_dict().items():
            out[f"c{ci}_final_" + k] = v.detach().numpy().copy()
        if ci == 0:
            out["valid_after"] = np.array(tr.validate(vb))
            out["test_after"] = np.array(tr.evaluate(eb))
    np.savez_compressed(os.path.join(HERE, "cdae_small.npz"), **out)
    print("cdae_small: losses", out["c0_losses"], "valid", out["valid_after"])


# ------------------------------------------------------------------------------------------------
def golden_run_loop():
    """BaseTrainer.run (trainers/mf_trainer.py:34-97): epoch loop, best-metric selection, best_model.pt, early stop.
    Adam at lr 0.05 over-fits the 160-user problem within a few epochs, so the validation loss turns and patience = 1
    stops the loop before cfg.epochs — the recorded per-epoch history and the saved best model pin that behaviour."""
    inter = syn.make_interactions(num_users=160, num_items=240, nnz=3200, seed=5, n_clusters=4)
    pipe = MFDataPipeline(cfg())
    pipe.num_users, pipe.num_items = inter.num_users, inter.num_items
    _, _, valid_eval, _ = pipe.split(frame(inter))
    split = syn.split_per_user(inter, seed=42)
    tu, tp, tn = syn.sample_triples(split, inter.num_items, seed=42)
    vu, vp, vn = syn.sample_triples(split, inter.num_items, seed=43, which="valid", reject="train+valid")
    batches, vbatches = syn.to_batches(tu, tp, tn, 256), syn.to_batches(vu, vp, vn, 256)
    out = {}
    for tag, kw in (("loss", dict(best_metric="loss", patience=1)), ("recall", dict(best_metric="recall", patience=2))):
        mdir = tempfile.mkdtemp()
        torch.manual_seed(42)
        tr = RefMFTrainer(cfg(optimizer="adam", lr=0.05, epochs=12, model_dir=mdir, **kw), inter.num_items, inter.num_users)
        out[f"{tag}_U0"] = tr.model.user_embedding.weight.detach().numpy().copy()
        out[f"{tag}_V0"] = tr.model.item_embedding.weight.detach().numpy().copy()
        hist = {"train": [], "valid": [], "metrics": []}
        for name in ("train", "validate", "evaluate"):
            orig = getattr(tr, name)

            def wrapped(*a, _orig=orig, _name=name, **k):
                r = _orig(*a, **k)
                hist[{"train": "train", "validate": "valid", "evaluate": "metrics"}[_name]].append(r)
                return r
            setattr(tr, name, wrapped)
        tr.run(batches, vbatches, valid_eval)
        best = torch.load(os.path.join(mdir, "best_model.pt"))
        out[f"{tag}_train"] = np.array(hist["train"])
        out[f"{tag}_valid"] = np.array(hist["valid"])
        out[f"{tag}_metrics"] = np.array(hist["metrics"])
        out[f"{tag}_best_U"] = best["user_embedding.weight"].numpy()
        out[f"{tag}_best_V"] = best["item_embedding.weight"].numpy()
        out[f"{tag}_keys"] = np.array(sorted(best.keys()))
        print(f"run_loop[{tag}]: epochs run {len(hist['train'])} of 12; valid", np.round(hist["valid"], 4),
              "recall", np.round([m[1] for m in hist["metrics"]], 4))
    np.savez_compressed(os.path.join(HERE, "run_loop.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cdae_widths":
        golden_cdae_widths()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cdae":
        golden_cdae()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "run_loop":
        golden_run_loop()
        sys.exit(0)
    golden_metrics()
    golden_split_and_mf()
    golden_ngcf()
    golden_cdae()
    golden_run_loop()
    print("fixtures written to", HERE)
    os.system(f"ls -la {HERE}")
