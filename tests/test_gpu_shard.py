"""GPU: row-sharded BPR-MF (BASELINE config 5). The reference cannot run this configuration, so the
checker is the oracle's MFPort on the unsharded tables: tolerance 1e-5 norm-wise (fp32; duplicate-row sums reorder)."""
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from util import rel_fro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("optname,lr,wd,d", [("sgd", 0.05, 0.0, 64), ("sgd", 0.05, 1e-3, 32), ("adam", 1e-2, 1e-4, 128),
                                             ("adamw", 1e-2, 1e-2, 256)])
def test_sharded_trainer_world1_matches_oracle(optname, lr, wd, d):
    from oracle.torch_port import MFPort
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.trainers.sharded_mf_trainer import ShardedMFTrainer
    inter = syn.make_interactions(num_users=2000, num_items=900, nnz=40000, seed=13, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    batches = syn.to_batches(u, p, n, 1023)[:6]
    U0, V0 = (torch.from_numpy(np.ascontiguousarray(a)) for a in syn.planted_embeddings(inter, d=d, seed=5))
    cfg = SimpleNamespace(embed_size=d, optimizer=optname, lr=lr, weight_decay=wd, seed=1)
    port = MFPort(U0, V0, optimizer=optname, lr=lr, weight_decay=wd)
    ref_loss, _ = port.train(batches)
    out = {}
    for exchange, adam_mode in (("all_to_all", "sparse"), ("all_to_all", "sparse"), ("all_to_all", "dense"), ("all_reduce", "dense")):
        tr = ShardedMFTrainer(cfg, inter.num_items, inter.num_users, init=(U0, V0), exchange=exchange, adam_mode=adam_mode)
        loss = tr.train(batches)
        U, V = tr.gather_tables()
        assert rel_fro(U.cpu(), port.user.weight.detach()) < 1e-5, (exchange, adam_mode)
        assert rel_fro(V.cpu(), port.item.weight.detach()) < 1e-5, (exchange, adam_mode)
        assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss)
        # scratch is left clean for the next step
        assert int(tr._sh["V"]["flags"].sum()) == 0 and float(tr._sh["V"]["g"].abs().sum()) == 0.0
        out.setdefault((exchange, adam_mode), []).append((U.clone(), V.clone(), loss))
    # ordered accumulate: two runs are bit-identical; the sparse-traffic (catch-up) Adam equals the dense sweep bit for bit
    (Ua, Va, la), (Ub, Vb, lb) = out[("all_to_all", "sparse")]
    assert torch.equal(Ua, Ub) and torch.equal(Va, Vb) and la == lb
    Ud, Vd, _ = out[("all_to_all", "dense")][0]
    assert torch.equal(Ua, Ud) and torch.equal(Va, Vd)


def test_sharded_trainer_oob_id_raises():
    from yelprecommendation_b200.trainers.sharded_mf_trainer import ShardedMFTrainer
    cfg = SimpleNamespace(embed_size=32, optimizer="sgd", lr=0.1, weight_decay=0.0, seed=1)
    tr = ShardedMFTrainer(cfg, 50, 40)
    b = {"user_id": torch.tensor([1, 40]), "pos_item": torch.tensor([2, 3]), "neg_item": torch.tensor([4, 5])}
    with pytest.raises(IndexError):
        tr.train([b])


def _ngcf_problem(seed=21, nu=1203, ni=958, nnz=30000, d=64, layers=3):
    from yelprecommendation_b200.data import synthetic as syn
    from yelprecommendation_b200.data.graph import build_laplacian
    inter = syn.make_interactions(num_users=nu, num_items=ni, nnz=nnz, seed=seed, n_clusters=4, star_ratings=True)
    split = syn.split_per_user(inter, seed=42)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    batches = syn.to_batches(u, p, n, 1023)[:4]
    g = torch.Generator().manual_seed(3)
    init = {"embedding.weight": torch.randn(nu + ni, d, generator=g) * 0.3}
    for l in range(layers):
        init[f"W1.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
        init[f"W2.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
    return inter, L, batches, init


@pytest.mark.parametrize("optname,lr,wd,layers,d,mode", [("sgd", 0.05, 0.0, 3, 64, 2), ("adam", 1e-2, 1e-4, 3, 64, 2),
                                                         ("adamw", 2e-3, 1e-2, 1, 64, 2), ("sgd", 0.05, 0.0, 1, 128, 2),
                                                         ("adam", 1e-2, 1e-4, 3, 32, 2), ("sgd", 0.05, 0.0, 3, 128, 2),
                                                         ("adam", 1e-3, 0.0, 3, 128, 2), ("adam", 1e-3, 0.0, 3, 128, 0),
                                                         ("adam", 1e-2, 1e-4, 3, 64, 1)])
def test_sharded_ngcf_world1_matches_oracle(optname, lr, wd, layers, d, mode):
    """Op-by-op sharded path (row panels of the SpMM block + yr_ngcf_dense_fwd/bwd + shard gather/scatter) vs the oracle;
    d = 128 with 3 layers (concatenated width 512) is BASELINE config 5's model. mode = yr_dense_mode (2 = tensor cores for
    both passes, the default). The tail sums duplicate rows in batch order without atomics, so a step is reproducible
    (test_sharded_ngcf_step_reproducible). The d = 128 x 3-layer Adam case runs at lr = 1e-3 (10 x the reference's NGCF
    learning rate): at lr = 1e-2 without weight decay that model is ill-conditioned — a weight-gradient element near zero
    changes sign with the last bits and Adam's first step moves the weight by 2 lr; profiles/r02_adam_dense_modes.txt shows
    round 2's base build itself (FP32 pipe, atomics) flipping between 5e-7 and 2e-5 from run to run on it, and the torch-CPU
    port of another host selecting the other outcome for the FP32-pipe mode."""
    from oracle.torch_port import NGCFPort
    from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer
    inter, L, batches, init = _ngcf_problem(layers=layers, d=d)
    cfg = SimpleNamespace(embed_size=d, num_orders=layers, optimizer=optname, lr=lr, weight_decay=wd, seed=1,
                          ngcf_dense_mode=mode)
    tr = ShardedNGCFTrainer(cfg, inter.num_items, inter.num_users, L, init=init, n_panels=3)
    assert len(tr.panels) == 3 and tr.k.dense_mode == mode
    loss = tr.train(batches)
    port = NGCFPort(init["embedding.weight"], [init[f"W1.{l}.weight"] for l in range(layers)],
                    [init[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, optname, lr, wd)
    ref_loss, _ = port.train(batches)
    assert rel_fro(tr.gather_embedding().cpu(), port.emb.detach()) < 1e-5
    for l in range(layers):
        assert rel_fro(tr.W1[l].cpu(), port.W1[l].detach()) < 1e-5
        assert rel_fro(tr.W2[l].cpu(), port.W2[l].detach()) < 1e-5
    assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss)


def test_sharded_ngcf_step_reproducible():
    """Two runs of the same sharded NGCF steps give bit-identical embeddings (Adam, d = 64, 3 layers): the tail accumulates
    in batch order (yr_shard_accumulate_sorted), the SpMM and the dW partials are reduced in fixed order."""
    from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer
    inter, L, batches, init = _ngcf_problem(layers=3, d=64)
    cfg = SimpleNamespace(embed_size=64, num_orders=3, optimizer="adam", lr=1e-2, weight_decay=1e-4, seed=1)
    outs = []
    for _ in range(2):
        tr = ShardedNGCFTrainer(cfg, inter.num_items, inter.num_users, L, init=init, n_panels=3)
        tr.train(batches)
        outs.append((tr.gather_embedding().clone(), [w.clone() for w in tr.W1]))
    assert torch.equal(outs[0][0], outs[1][0])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][1], outs[1][1]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_trainer_nccl_world2():
    script = os.path.join(ROOT, "tests", "_dist_shard_gpu_worker.py")
    port = 33500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_SHARD_GPU_OK" in r.stdout


# ---- BASELINE config 5 inputs built on the device (data/scaled.py) --------------------------------------------------
def _small_scaled(nU=3000, nI=700, nnz=60_000):
    from yelprecommendation_b200.data.scaled import make_scaled_graph
    return make_scaled_graph(nU, nI, nnz, seed=5, device="cuda")


def test_scaled_graph_generator_properties():
    """Philox generator kernel: sorted duplicate-free item lists, every user >= min_deg draws, requested size within 1 %,
    Zipf head (the most popular item is far above the mean), and the same graph on every call (counter-based)."""
    g = _small_scaled()
    ptr, items = g.user_ptr.cpu().numpy().astype(np.int64), g.user_items.cpu().numpy()
    assert ptr[0] == 0 and ptr[-1] == items.size and abs(items.size - 60_000) <= 0.01 * 60_000
    assert items.min() >= 0 and items.max() < 700
    for u in (0, 1, 17, 2999):
        row = items[ptr[u]:ptr[u + 1]]
        assert row.size >= 1 and np.all(np.diff(row) > 0)
    lens = np.diff(ptr)
    starts = ptr[:-1]
    inner = np.ones(items.size, bool)
    inner[starts] = False
    assert np.all(np.diff(items)[inner[1:]] > 0)                        # strictly ascending inside every row
    deg = np.bincount(items, minlength=700)
    assert deg.max() > 8 * deg.mean() and lens.max() > 3 * lens.mean()
    g2 = _small_scaled()
    assert torch.equal(g.user_ptr, g2.user_ptr) and torch.equal(g.user_items, g2.user_items)


@pytest.mark.parametrize("world,panels", [(1, 1), (3, 1), (8, 1), (1, 4), (4, 8)])
def test_shard_laplacian_blocks_equal_the_reference_laplacian(world, panels):
    """Row blocks cut on the device (shard_laplacian: binary ratings, yr_laplacian_binary_values) re-assembled in the
    reference's node order == data.graph.build_laplacian of the same interactions (pinned to the reference,
    tests/test_oracle_golden.py) bit for bit; the COO entry point gives the same blocks."""
    from yelprecommendation_b200.data.graph import build_laplacian, coo_to_csr
    from yelprecommendation_b200.data.scaled import ShardLayout, shard_laplacian, shard_laplacian_from_coo
    g = _small_scaled()
    nU, nI = g.num_users, g.num_items
    ptr, items = g.user_ptr.cpu().numpy().astype(np.int64), g.user_items.cpu().numpy().astype(np.int64)
    users = np.repeat(np.arange(nU), np.diff(ptr))
    L = build_laplacian(users, items, np.ones(items.size), nU, nI)
    idx, val = L.indices().numpy(), L.values().numpy()
    rp_ref, ci_ref, va_ref = coo_to_csr(idx[0], idx[1], val, nU + nI)
    lay = ShardLayout(nU, nI, world, panels)
    pos_of_node = lay.node_pos(torch.arange(nU + nI)).numpy()
    node_of_pos = np.full(world * lay.per, -1, np.int64)
    node_of_pos[pos_of_node] = np.arange(nU + nI)
    seen = 0
    for rank in range(world):
        rp, ci, va = (t.cpu().numpy() for t in shard_laplacian(g, lay, rank))
        rp2, ci2, va2 = (t.cpu().numpy() for t in shard_laplacian_from_coo(L, lay, rank, "cuda"))
        assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(va, va2)
        for r in range(lay.per):
            node = node_of_pos[rank * lay.per + r]
            if node < 0:
                assert rp[r + 1] == rp[r]
                continue
            a, b = rp_ref[node], rp_ref[node + 1]
            assert np.array_equal(node_of_pos[ci[rp[r]:rp[r + 1]]], ci_ref[a:b])
            assert np.array_equal(va[rp[r]:rp[r + 1]], va_ref[a:b])
            seen += b - a
    assert seen == ci_ref.size


def test_sharded_ngcf_on_scaled_graph_world1_matches_port():
    """ShardedNGCFTrainer fed with a ScaledGraph (device-built blocks, row panels) vs the torch-CPU port on the same graph."""
    from oracle.torch_port import NGCFPort
    from yelprecommendation_b200.data.graph import build_laplacian
    from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer
    g = _small_scaled()
    nU, nI, d, layers = g.num_users, g.num_items, 128, 3
    ptr, items = g.user_ptr.cpu().numpy().astype(np.int64), g.user_items.cpu().numpy().astype(np.int64)
    users = np.repeat(np.arange(nU), np.diff(ptr))
    L = build_laplacian(users, items, np.ones(items.size), nU, nI)
    rng = np.random.default_rng(0)
    B = 4096
    batches = [{"user_id": torch.from_numpy(rng.integers(0, nU, B)), "pos_item": torch.from_numpy(rng.integers(0, nI, B)),
                "neg_item": torch.from_numpy(rng.integers(0, nI, B))} for _ in range(2)]
    gen = torch.Generator().manual_seed(3)
    init = {"embedding.weight": torch.randn(nU + nI, d, generator=gen) * 0.3}
    for l in range(layers):
        init[f"W1.{l}.weight"] = (torch.rand(d, d, generator=gen) * 2 - 1) / 11
        init[f"W2.{l}.weight"] = (torch.rand(d, d, generator=gen) * 2 - 1) / 11
    cfg = SimpleNamespace(embed_size=d, num_orders=layers, optimizer="adam", lr=1e-2, weight_decay=0.0, seed=1)
    tr = ShardedNGCFTrainer(cfg, nI, nU, g, init=init, n_panels=4)
    loss = tr.train(batches)
    port = NGCFPort(init["embedding.weight"], [init[f"W1.{l}.weight"] for l in range(layers)],
                    [init[f"W2.{l}.weight"] for l in range(layers)], nU, L, "adam", 1e-2, 0.0)
    ref_loss, _ = port.train(batches)
    assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss)
    assert rel_fro(tr.gather_embedding().cpu(), port.emb.detach()) < 2e-5
