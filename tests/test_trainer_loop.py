"""GPU: BaseTrainer.run — epoch loop, best-metric selection, best_model.pt round trip and early stopping — against the
history recorded from the REAL reference loop (trainers/mf_trainer.py:34-97) by tests/golden/make_golden.py::golden_run_loop
on the same batches. Tolerance: 1e-5 relative on losses and saved tables (fp32 Adam, lr 0.05, up to 9
epochs of 8 steps); metrics to 1e-5 unless a fp32-noise tie swaps an item (then within one hit: 1/(10*160))."""
import os
import tempfile
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from util import batches_from, lists_from, load_npz, rel_fro

pytestmark = pytest.mark.gpu


def _setup(tag, **kw):
    from yelprecommendation_b200.trainers import MFTrainer
    g, r = load_npz("mf_small.npz"), load_npz("run_loop.npz")
    mdir = tempfile.mkdtemp()
    cfg = SimpleNamespace(device="cuda", model_dir=mdir, embed_size=64, optimizer="adam", lr=0.05, weight_decay=0.0, top_n=10,
                          wandb=False, epochs=12, batch_size=256, **kw)
    tr = MFTrainer(cfg, int(g["num_items"]), int(g["num_users"]))
    with torch.no_grad():
        tr.model.user_embedding.weight.copy_(torch.from_numpy(r[f"{tag}_U0"]))
        tr.model.item_embedding.weight.copy_(torch.from_numpy(r[f"{tag}_V0"]))
    ev = pd.DataFrame({"pos_items": lists_from(g, "valid_eval", "pos_items"),
                       "mask_items": lists_from(g, "valid_eval", "mask_items")},
                      index=pd.Index(g["valid_eval_uid"], name="user_id"))
    batches = batches_from(g["tri_u"], g["tri_p"], g["tri_n"], 256)
    vbatches = batches_from(g["vtri_u"], g["vtri_p"], g["vtri_n"], 256)
    hist = {"train": [], "validate": [], "evaluate": []}
    for name in hist:
        orig = getattr(tr, name)

        def wrapped(*a, _orig=orig, _name=name, **k):
            out = _orig(*a, **k)
            hist[_name].append(out)
            return out
        setattr(tr, name, wrapped)
    return tr, r, ev, batches, vbatches, hist, mdir


@pytest.mark.parametrize("tag,kw", [("loss", dict(best_metric="loss", patience=1)),
                                    ("recall", dict(best_metric="recall", patience=2))])
def test_run_loop_matches_reference_history(tag, kw):
    tr, r, ev, batches, vbatches, hist, mdir = _setup(tag, **kw)
    tr.run(batches, vbatches, ev)
    n_ref = len(r[f"{tag}_train"])
    assert n_ref < 12                                                  # the reference stopped early ...
    assert len(hist["train"]) == n_ref                                 # ... and so did we, at the same epoch
    assert np.allclose(hist["train"], r[f"{tag}_train"], rtol=1e-5)
    assert np.allclose(hist["validate"], r[f"{tag}_valid"], rtol=1e-5)
    got_m, ref_m = np.array(hist["evaluate"]), r[f"{tag}_metrics"]
    assert got_m.shape == ref_m.shape and np.abs(got_m - ref_m).max() <= 1.0 / (10 * len(ev)) + 1e-9
    best = torch.load(os.path.join(mdir, "best_model.pt"))
    assert sorted(best.keys()) == [str(k) for k in r[f"{tag}_keys"]]   # same state_dict keys as the reference's file
    assert rel_fro(best["user_embedding.weight"].cpu().numpy(), r[f"{tag}_best_U"]) < 1e-5
    assert rel_fro(best["item_embedding.weight"].cpu().numpy(), r[f"{tag}_best_V"]) < 1e-5
    # the model kept training past the best epoch; load_best_model brings the saved one back
    assert rel_fro(tr.model.user_embedding.weight.data.cpu().numpy(), r[f"{tag}_best_U"]) > 1e-3
    tr.load_best_model()
    assert rel_fro(tr.model.user_embedding.weight.data.cpu().numpy(), r[f"{tag}_best_U"]) < 1e-5
    assert tr.model.user_embedding.weight.is_cuda


def test_reference_checkpoint_loads_into_dropin():
    """A best_model.pt written by the reference (CPU tensors, reference key names) loads into the drop-in model."""
    tr, r, ev, batches, vbatches, hist, mdir = _setup("loss", best_metric="loss", patience=1)
    torch.save({"user_embedding.weight": torch.from_numpy(r["loss_best_U"]),
                "item_embedding.weight": torch.from_numpy(r["loss_best_V"])}, os.path.join(mdir, "best_model.pt"))
    tr.load_best_model()
    assert np.array_equal(tr.model.item_embedding.weight.data.cpu().numpy(), r["loss_best_V"])
    assert abs(tr.validate(vbatches) - float(r["loss_valid"][0])) <= 1e-5 * float(r["loss_valid"][0])
