"""CPU: the C-ABI library builds, loads and exports every symbol include/yelprec_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yelprec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from yelprecommendation_b200 import _cabi
    import __graft_entry__ as ge
    ge.build()
    lib = _cabi.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but missing from the .so"
    assert sorted(_cabi.SYMBOLS) == syms, "ctypes binding and header disagree"
    assert lib.yr_version() >= 100


def test_sass_is_sm100a_only():
    import subprocess
    from yelprecommendation_b200 import _cabi
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    import torch
    from yelprecommendation_b200 import ops, _cabi
    with pytest.raises(_cabi.YelprecError):
        ops.bpr_loss(torch.zeros(4), torch.zeros(4))
    with pytest.raises(_cabi.YelprecError):
        _cabi.dptr(torch.zeros(4))


def test_optimizer_name_errors_like_reference():
    from yelprecommendation_b200.trainers.base_trainer import FusedOptimizer
    with pytest.raises(NotImplementedError, match="Optimizer Not Exists"):
        FusedOptimizer("rmsprop", 1e-3)
    assert FusedOptimizer("AdamW", 1e-3, 0.1).name == "adamw"


def test_pytorch_custom_ops_are_registered():
    """north_star: the path drops in via PyTorch custom ops bound through the C ABI — torch.ops.yelprec.* exist with
    tensor-only schemas, fake kernels (shape inference without a GPU) and autograd formulas."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from yelprecommendation_b200 import ops  # noqa: F401  (registers the ops)
    want = {"mf_score": "yelprec::mf_score(Tensor U, Tensor V, Tensor uid, Tensor iid) -> Tensor",
            "mf_score_bwd": "yelprec::mf_score_bwd(Tensor U, Tensor V, Tensor uid, Tensor iid, Tensor gout) -> (Tensor, Tensor)",
            "bpr_loss": "yelprec::bpr_loss(Tensor pos, Tensor neg) -> Tensor",
            "bpr_loss_bwd": "yelprec::bpr_loss_bwd(Tensor pos, Tensor neg, Tensor gloss) -> (Tensor, Tensor)",
            "ngcf_layer": "yelprec::ngcf_layer(Tensor E, Tensor W1, Tensor W2, SymInt csr_handle, float slope, SymInt dense_mode) -> (Tensor, Tensor)"}
    for name, schema in want.items():
        assert str(getattr(torch.ops.yelprec, name).default._schema) == schema
    with FakeTensorMode():
        U, V = torch.empty(10, 64, device="cuda"), torch.empty(12, 64, device="cuda")
        u = torch.empty(5, dtype=torch.int64, device="cuda")
        assert torch.ops.yelprec.mf_score(U, V, u, u).shape == (5,)
        assert torch.ops.yelprec.bpr_loss(torch.empty(5, device="cuda"), torch.empty(5, device="cuda")).shape == ()
    # no CPU kernels: the product has no CPU path
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.yelprec.bpr_loss(torch.zeros(4), torch.zeros(4))
