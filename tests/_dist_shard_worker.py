"""world_size-2 gloo worker for test_host_logic: drives ShardedMFTrainer's collective choreography (owner gather ->
all-reduce -> sliced gradient rows -> all-gather -> owner accumulate + step) with a CPU restatement of the four
device-side pieces, and checks the gathered tables and the loss against the oracle's single-process MFPort.
On GPUs the same class runs with CabiShardKernels (tests/test_gpu_shard.py)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.torch_port import MFPort, NGCFPort                                    # noqa: E402
from yelprecommendation_b200 import _cabi                                          # noqa: E402
from yelprecommendation_b200.data import synthetic as syn                          # noqa: E402
from yelprecommendation_b200.data.graph import build_laplacian                     # noqa: E402
from yelprecommendation_b200.trainers.sharded_mf_trainer import ShardedMFTrainer   # noqa: E402
from yelprecommendation_b200.trainers.sharded_ngcf_trainer import ShardedNGCFTrainer  # noqa: E402


class CpuShardKernels:
    """Semantics of yr_shard_gather_rows / yr_bpr_rows_grad / yr_shard_accumulate / yr_shard_step on CPU tensors
    (include/yelprec_b200.h). Test infrastructure only."""

    def gather_rows(self, s, total, ids, R, which, err):
        if bool(((ids < 0) | (ids >= total)).any()):
            err.fill_(1)
        own = (ids >= s["lo"]) & (ids < s["hi"])
        R[:, which, :] = 0
        R[own, which, :] = s["T"][ids[own] - s["lo"]]

    def rows_grad(self, R, B, b0, b1, Gs, loss_acc):
        u, p, n = R[b0:b1, 0], R[b0:b1, 1], R[b0:b1, 2]
        x = (u * p).sum(1) - (u * n).sum(1)
        loss_acc += (-F.logsigmoid(x)).double().sum()
        g = (-torch.sigmoid(-x) / B).unsqueeze(1)
        Gs[: b1 - b0, 0] = g * p - g * n
        Gs[: b1 - b0, 1] = g * u
        Gs[: b1 - b0, 2] = -(g * u)

    def accumulate(self, s, opt, ids, G, which):
        own = (ids >= s["lo"]) & (ids < s["hi"])
        j = torch.nonzero(own).flatten()
        r = ids[j] - s["lo"]
        s["g"].index_add_(0, r, G[j, which])
        s["flags"][r] = 1

    def step(self, s, opt, max_rows, d):
        kind = {v: k for k, v in _cabi.OPT_KINDS.items()}[opt.kind]
        n = s["hi"] - s["lo"]
        T, g = s["T"], s["g"][:n]
        if kind == "sgd":
            T -= opt.lr * (g + opt.weight_decay * T)
        else:
            if kind == "adam":
                g = g + opt.weight_decay * T
            else:
                T *= 1 - opt.lr * opt.weight_decay
            m, v = s["m"][:n], s["v"][:n]
            m.mul_(opt.beta1).add_(g, alpha=1 - opt.beta1)
            v.mul_(opt.beta2).addcmul_(g, g, value=1 - opt.beta2)
            bc1, bc2 = 1 - opt.beta1 ** opt.step, 1 - opt.beta2 ** opt.step
            T.addcdiv_(m, (v.sqrt() / np.sqrt(bc2)).add_(opt.eps), value=-opt.lr / bc1)
        s["g"].zero_()
        s["flags"].zero_()


def _cpu_a2a_methods():
    """the all-to-all choreography's device pieces on CPU tensors (yr_shard_gather_local / yr_bpr_rows_grad on a slice /
    yr_shard_accumulate_sorted / yr_shard_step_sparse_adam)"""

    def gather_local(self, su, sv, sel, row, out):
        r = row.long()
        out.copy_(torch.where((sel != 0).unsqueeze(1), sv["T"][r.clamp(max=sv["T"].shape[0] - 1)], su["T"][r.clamp(max=su["T"].shape[0] - 1)]))

    def rows_grad_slice(self, R, B, b0, b1, G, loss_acc):
        n = b1 - b0
        u, p, q = R[:n, 0], R[:n, 1], R[:n, 2]
        x = (u * p).sum(1) - (u * q).sum(1)
        loss_acc += (-F.logsigmoid(x)).double().sum()
        g = (-torch.sigmoid(-x) / B).unsqueeze(1)
        G[:n, 0] = g * p - g * q
        G[:n, 1] = g * u
        G[:n, 2] = -(g * u)

    def accumulate_sorted(self, s, opt, rows_sorted, src, G, list_rows):
        s["g"].index_add_(0, rows_sorted.long(), G[src.long()])
        s["flags"][rows_sorted.long()] = 1

    def adam_scalars(self, opt, n_steps, scal):
        pass

    def catch_up(self, s, opt, scal, n_scal, rows_sorted, d):
        rows = torch.unique(rows_sorted.long())
        last = s["last"][rows].long()
        for t in range(int(last.min()) + 1, opt.step):
            sel = rows[last < t]
            if sel.numel() == 0:
                continue
            T, m, v = s["T"][sel], s["m"][sel], s["v"][sel]
            _opt_step_cpu(T, torch.zeros(sel.numel(), d), m, v, type(opt)(opt.kind, t, opt.lr, opt.weight_decay, opt.beta1, opt.beta2, opt.eps))
            s["T"][sel], s["m"][sel], s["v"][sel] = T, m, v
        s["last"][rows] = torch.maximum(s["last"][rows], torch.full_like(s["last"][rows], opt.step - 1))

    def step_sparse_adam(self, s, opt, scal, n_scal, max_rows, d, flush=False):
        n = s["hi"] - s["lo"]
        t_now = opt.step
        rows = torch.arange(n) if flush else torch.nonzero(s["flags"][:n]).flatten()
        if rows.numel() == 0:
            return
        last = s["last"][rows].long()
        for t in range(int(last.min()) + 1, t_now + 1):
            sel = rows[last < t]
            if sel.numel() == 0:
                continue
            g = s["g"][sel] if (t == t_now and not flush) else torch.zeros(sel.numel(), d)
            T, m, v = s["T"][sel], s["m"][sel], s["v"][sel]
            o = type(opt)(opt.kind, t, opt.lr, opt.weight_decay, opt.beta1, opt.beta2, opt.eps)
            _opt_step_cpu(T, g, m, v, o)
            s["T"][sel], s["m"][sel], s["v"][sel] = T, m, v
        s["last"][rows] = t_now
        if not flush:
            s["g"][rows] = 0
            s["flags"][rows] = 0

    return dict(gather_local=gather_local, rows_grad_slice=rows_grad_slice, accumulate_sorted=accumulate_sorted,
                adam_scalars=adam_scalars, step_sparse_adam=step_sparse_adam, catch_up=catch_up)


def _opt_step_cpu(T, g, m, v, opt):
    kind = {v_: k for k, v_ in _cabi.OPT_KINDS.items()}[opt.kind]
    if kind == "sgd":
        T -= opt.lr * (g + opt.weight_decay * T)
        return
    if kind == "adam":
        g = g + opt.weight_decay * T
    else:
        T *= 1 - opt.lr * opt.weight_decay
    m.mul_(opt.beta1).add_(g, alpha=1 - opt.beta1)
    v.mul_(opt.beta2).addcmul_(g, g, value=1 - opt.beta2)
    bc1, bc2 = 1 - opt.beta1 ** opt.step, 1 - opt.beta2 ** opt.step
    T.addcdiv_(m, (v.sqrt() / np.sqrt(bc2)).add_(opt.eps), value=-opt.lr / bc1)


for _name, _fn in _cpu_a2a_methods().items():
    setattr(CpuShardKernels, _name, _fn)


class CpuNgcfShardKernels:
    """Semantics of yr_spmm_csr / yr_ngcf_dense_fwd / yr_ngcf_dense_bwd / yr_shard_gather_rows / yr_bpr_rows_grad /
    yr_shard_accumulate / yr_dense_opt_step on CPU tensors. Test infrastructure only."""

    def make_csr(self, rowptr, col, val):
        """rowptr: absolute offsets into col / val (row panels share the block's arrays)"""
        rowptr = np.asarray(rowptr, dtype=np.int64)
        n = len(rowptr) - 1
        rows = np.repeat(np.arange(n), np.diff(rowptr))
        sl = slice(int(rowptr[0]), int(rowptr[-1]))
        return (n, torch.from_numpy(rows.astype(np.int64)), torch.as_tensor(col)[sl].to(torch.int64), torch.as_tensor(val)[sl])

    def spmm(self, A, X, out, accumulate):
        n, rows, cols, vals = A
        y = torch.zeros(n, X.shape[1]).index_add_(0, rows, X[cols] * vals.unsqueeze(1))
        if accumulate:
            out += y
        else:
            out.copy_(y)

    def dense_fwd(self, E, LE, W1, W2, out):
        out.copy_(F.leaky_relu(F.linear(LE + E, W1) + F.linear(E * LE, W2), 0.01))

    def dense_bwd(self, E, LE, En, Gn, W1, W2, G, T, dW1, dW2):
        dZ = Gn * torch.where(En > 0, torch.ones_like(En), torch.full_like(En, 0.01))
        dS, dP = dZ @ W1, dZ @ W2
        G += dS + dP * LE
        T.copy_(dS + dP * E)
        dW1.copy_(dZ.t() @ (LE + E))
        dW2.copy_(dZ.t() @ (E * LE))

    def gather_rows(self, T, lo, hi, total, ids, R, col_off, err):
        d = T.shape[1]
        own = (ids >= lo) & (ids < hi)
        R[:, col_off:col_off + d] = 0
        R[own, col_off:col_off + d] = T[ids[own] - lo]

    def rows_grad(self, R, B, width, Gr, loss_acc):
        Rv, Gv = R.view(B, 3, width), Gr.view(B, 3, width)
        u, p, n = Rv[:, 0], Rv[:, 1], Rv[:, 2]
        x = (u * p).sum(1) - (u * n).sum(1)
        loss_acc += (-F.logsigmoid(x)).double().sum()
        g = (-torch.sigmoid(-x) / B).unsqueeze(1)
        Gv[:, 0] = g * p - g * n
        Gv[:, 1] = g * u
        Gv[:, 2] = -(g * u)

    def scatter_rows(self, G, lo, hi, ids, Gr, col_off, flags, scratch):
        d = G.shape[1]
        j = torch.nonzero((ids >= lo) & (ids < hi)).flatten()
        G.index_add_(0, ids[j] - lo, Gr[j, col_off:col_off + d])

    def scatter_rows_sorted(self, G, rows_sorted, src, Gr, col_off, flags, scratch):
        d = G.shape[1]
        keep = rows_sorted >= 0
        G.index_add_(0, rows_sorted[keep].long(), Gr[src[keep].long(), col_off:col_off + d])

    def opt_step(self, p, g, m, v, opt):
        _opt_step_cpu(p, g, m, v, opt)


def ngcf_case(rank):
    inter = syn.make_interactions(num_users=203, num_items=158, nnz=4000, seed=21, n_clusters=4, star_ratings=True)
    split = syn.split_per_user(inter, seed=42)
    L = build_laplacian(inter.user, inter.item, inter.rating, inter.num_users, inter.num_items)
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    batches = syn.to_batches(u, p, n, 257)[:3]
    N, d, layers = inter.num_users + inter.num_items, 64, 3
    g = torch.Generator().manual_seed(3)
    init = {"embedding.weight": torch.randn(N, d, generator=g) * 0.3}
    for l in range(layers):
        init[f"W1.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
        init[f"W2.{l}.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) / 8
    for optname, lr, wd in (("sgd", 0.05, 0.0), ("adam", 1e-2, 1e-4)):
        cfg = SimpleNamespace(embed_size=d, num_orders=layers, optimizer=optname, lr=lr, weight_decay=wd, seed=1)
        tr = ShardedNGCFTrainer(cfg, inter.num_items, inter.num_users, L, init=init, device="cpu", kernels=CpuNgcfShardKernels())
        loss = tr.train(batches)
        E0 = tr.gather_embedding()
        port = NGCFPort(init["embedding.weight"], [init[f"W1.{l}.weight"] for l in range(layers)],
                        [init[f"W2.{l}.weight"] for l in range(layers)], inter.num_users, L, optname, lr, wd)
        ref_loss, _ = port.train(batches)
        rel = float((E0 - port.emb.detach()).norm() / port.emb.detach().norm())
        assert rel < 2e-6, (optname, "E", rel)
        for l in range(layers):
            for got, ref in ((tr.W1[l], port.W1[l].detach()), (tr.W2[l], port.W2[l].detach())):
                rel = float((got - ref).norm() / ref.norm())
                assert rel < 5e-6, (optname, "W", l, rel)
        assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss), (loss, ref_loss)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    inter = syn.make_interactions(num_users=301, num_items=157, nnz=5000, seed=11, n_clusters=4)
    split = syn.split_per_user(inter, seed=42)
    U0, V0 = (torch.from_numpy(np.ascontiguousarray(a)) for a in syn.planted_embeddings(inter, d=32, seed=5))
    u, p, n = syn.sample_triples(split, inter.num_items, seed=9)
    batches = syn.to_batches(u, p, n, 509)[:4]          # 509: not a multiple of world -> ragged slices
    for optname, lr, wd in (("sgd", 0.05, 0.0), ("adam", 1e-2, 1e-4), ("adamw", 1e-2, 1e-2)):
        cfg = SimpleNamespace(embed_size=32, optimizer=optname, lr=lr, weight_decay=wd, seed=1)
        port = MFPort(U0, V0, optimizer=optname, lr=lr, weight_decay=wd)
        ref_loss, _ = port.train(batches)
        for exchange, adam_mode in (("all_to_all", "sparse"), ("all_to_all", "dense"), ("all_reduce", "dense")):
            tr = ShardedMFTrainer(cfg, inter.num_items, inter.num_users, init=(U0, V0), device="cpu", kernels=CpuShardKernels(),
                                  exchange=exchange, adam_mode=adam_mode)
            assert (tr.u1 - tr.u0) in (150, 151) and (tr.i1 - tr.i0) in (78, 79)
            loss = tr.train(batches)
            U, V = tr.gather_tables()
            for got, ref in ((U, port.user.weight.detach()), (V, port.item.weight.detach())):
                assert got.shape == ref.shape
                rel = float((got - ref).norm() / ref.norm())
                assert rel < 2e-6, (optname, exchange, adam_mode, rel)
            assert abs(loss - ref_loss) < 1e-5 * abs(ref_loss), (loss, ref_loss)
    # out-of-range id -> IndexError on every rank (nn.Embedding behaviour)
    bad = dict(batches[0])
    bad["pos_item"] = bad["pos_item"].clone()
    bad["pos_item"][3] = inter.num_items
    try:
        tr.train([bad])
        raise AssertionError("expected IndexError")
    except IndexError:
        pass
    ngcf_case(rank)
    dist.barrier()
    if rank == 0:
        print("DIST_SHARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
