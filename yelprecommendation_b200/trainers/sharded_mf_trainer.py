"""Row-sharded BPR-MF training across GPUs (BASELINE config 5: 10 M users x 2 M items, d = 128) — SURVEY.md §8(e).

One process per GPU (torch.distributed, NCCL over NVLink). Rank r owns a contiguous block of user rows and of item
rows together with their optimizer state; nothing else is replicated. Every rank sees the same batch of triples and
computes the slice [r * S, (r + 1) * S) of it, S = ceil(B / world). One step (exchange = 'all_to_all', the default):

    1. plan                   owner of every (triple, role) slot from the ids; who sends how many rows to whom (one 2 x world
                              int copy to the host per step)
    2. yr_shard_gather_local  every owner packs, in slot order, the rows each requester's slice needs
    3. exchange rows          grouped NCCL send/recv: rank r receives exactly the 3 S rows of its slice (3 S d 4 bytes in,
                              the same out) — 1 / world of what an all-gather of the batch's rows moves
    4. yr_bpr_rows_grad       loss + gradient rows of the slice
    5. exchange gradients     the same pattern backwards: every gradient row goes to the owner of its table row
    6. yr_shard_accumulate_sorted   owners group the received rows by table row (stable sort of row ids: index plumbing) and sum
                              every segment left to right — no floating-point atomics, bit-identical from run to run
    7. step                   plain SGD: the listed rows; Adam / AdamW: sparse-traffic catch-up (yr_shard_step_sparse_adam,
                              bit-identical to the dense sweep) or, adam_mode = 'dense', the sweep over every local row

exchange = 'all_reduce' keeps round 1's choreography (owner gather -> SUM all-reduce of [B x 3 x d] -> all-gather of the
gradient rows -> atomics accumulate) as the comparison. The arithmetic per row is the single-GPU fused trainer's (reference
trainers/mf_trainer.py:104-114), so the gathered tables equal MFTrainer's after the same triples up to the fp32 order in which
duplicate rows are summed. With world_size == 1 the exchanges are no-ops (the 1-GPU parity reference: the reference itself
cannot run this configuration).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist

from .. import _cabi, ops
from ..parallel import shard_range
from .base_trainer import FusedOptimizer

I32, I64, F32, F64 = torch.int32, torch.int64, torch.float32, torch.float64


class CabiShardKernels:
    """The four device-side pieces of a step, through the C-ABI (include/yelprec_b200.h, "Row-sharded BPR-MF").
    tests/_dist_shard_worker.py substitutes a CPU restatement to exercise the collective choreography under gloo."""

    def __init__(self, device):
        self.device = device
        self.lib = _cabi.load()

    def _st(self):
        return _cabi.stream_ptr(self.device)

    @staticmethod
    def _state(s, d) -> _cabi.YrShardState:
        p = _cabi.dptr
        return _cabi.YrShardState(p(s["T"], F32), p(s["m"]), p(s["v"]), p(s["g"]), p(s["flags"]), p(s["rows_list"]),
                                  p(s["counters"]), s["lo"], s["hi"], d)

    def gather_rows(self, s, total, ids, R, which, err):
        B, _, d = R.shape
        _cabi.check(self.lib.yr_shard_gather_rows(_cabi.dptr(s["T"], F32), s["lo"], s["hi"], total, d, _cabi.dptr(ids, I64),
                                                  int(ids.numel()), R.data_ptr() + which * d * 4, 3 * d,
                                                  _cabi.dptr(err), self._st()), "yr_shard_gather_rows")

    def rows_grad(self, R, B, b0, b1, Gs, loss_acc):
        d = R.shape[2]
        # the kernel indexes G by the absolute triple number: shift the base so that triple b0 lands at Gs[0]
        _cabi.check(self.lib.yr_bpr_rows_grad(_cabi.dptr(R, F32), d, B, b0, b1, _cabi.dptr(Gs, F32) - b0 * 3 * d * 4,
                                              _cabi.dptr(loss_acc, F64), self._st()), "yr_bpr_rows_grad")

    def accumulate(self, s, opt, ids, G, which):
        d = G.shape[2]
        st = self._state(s, d)
        _cabi.check(self.lib.yr_shard_accumulate(C.byref(st), C.byref(opt), _cabi.dptr(ids, I64), int(ids.numel()),
                                                 G.data_ptr() + which * d * 4, 3 * d, self._st()), "yr_shard_accumulate")

    def step(self, s, opt, max_rows, d):
        st = self._state(s, d)
        _cabi.check(self.lib.yr_shard_step(C.byref(st), C.byref(opt), max_rows, self._st()), "yr_shard_step")

    # ---- all-to-all choreography ------------------------------------------------------------------------------
    def gather_local(self, su, sv, sel, row, out):
        d = out.shape[1]
        _cabi.check(self.lib.yr_shard_gather_local(_cabi.dptr(su["T"], F32), _cabi.dptr(sv["T"], F32), d, _cabi.dptr(sel, I32),
                                                   _cabi.dptr(row, I32), int(sel.numel()), _cabi.dptr(out, F32), d, self._st()),
                    "yr_shard_gather_local")

    def rows_grad_slice(self, R, B, b0, b1, G, loss_acc):
        """R, G: [S x 3 x d] hold triples b0 .. b1-1 of the global batch of B (the mean is over B)"""
        d = R.shape[2]
        off = b0 * 3 * d * 4
        _cabi.check(self.lib.yr_bpr_rows_grad(_cabi.dptr(R, F32) - off, d, B, b0, b1, _cabi.dptr(G, F32) - off,
                                              _cabi.dptr(loss_acc, F64), self._st()), "yr_bpr_rows_grad")

    def accumulate_sorted(self, s, opt, rows_sorted, src, G, list_rows):
        d = G.shape[1]
        st = self._state(s, d)
        _cabi.check(self.lib.yr_shard_accumulate_sorted(C.byref(st), C.byref(opt), _cabi.dptr(rows_sorted, I32), _cabi.dptr(src, I32),
                                                        int(rows_sorted.numel()), _cabi.dptr(G, F32), d, 1 if list_rows else 0,
                                                        self._st()), "yr_shard_accumulate_sorted")

    def adam_scalars(self, opt, n_steps, scal):
        _cabi.check(self.lib.yr_adam_scalars(C.byref(opt), n_steps, _cabi.dptr(scal, F32), self._st()), "yr_adam_scalars")

    def catch_up(self, s, opt, scal, n_scal, rows_sorted, d):
        st = self._state(s, d)
        _cabi.check(self.lib.yr_shard_catch_up(C.byref(st), C.byref(opt), _cabi.dptr(scal, F32), n_scal, _cabi.dptr(s["last"], I32),
                                               _cabi.dptr(rows_sorted, I32), int(rows_sorted.numel()), self._st()),
                    "yr_shard_catch_up")

    def step_sparse_adam(self, s, opt, scal, n_scal, max_rows, d, flush=False):
        st = self._state(s, d)
        _cabi.check(self.lib.yr_shard_step_sparse_adam(C.byref(st), C.byref(opt), _cabi.dptr(scal, F32), n_scal,
                                                       _cabi.dptr(s["last"], I32), max_rows, 1 if flush else 0, self._st()),
                    "yr_shard_step_sparse_adam")


class ShardedMFTrainer:
    def __init__(self, cfg, num_items: int, num_users: int, init=None, group=None, device=None, kernels=None,
                 exchange: str = None, adam_mode: str = None):
        """`init`: optional (U [num_users x d], V [num_items x d]) full tables to slice (tests); otherwise each shard is
        drawn with the xavier-uniform bound of the FULL table (models/mf.py:15-18) from a per-rank generator.
        `device`/`kernels` default to the current CUDA device and the C-ABI kernels; there is no CPU product path."""
        self.cfg, self.num_items, self.num_users = cfg, num_items, num_users
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.kernels = kernels if kernels is not None else CabiShardKernels(self.device)
        self.d = d = int(cfg.embed_size)
        self.optimizer = FusedOptimizer(cfg.optimizer, cfg.lr, cfg.weight_decay)
        self.exchange = exchange or getattr(cfg, "shard_exchange", "all_to_all")
        self.adam_mode = adam_mode or getattr(cfg, "shard_adam_mode", "sparse")
        if self.exchange not in ("all_to_all", "all_reduce") or self.adam_mode not in ("sparse", "dense"):
            raise ValueError("exchange in ('all_to_all', 'all_reduce'), adam_mode in ('sparse', 'dense')")
        self.u0, self.u1 = shard_range(num_users, self.rank, self.world)
        self.i0, self.i1 = shard_range(num_items, self.rank, self.world)
        # block starts of every rank (owner of an id = the block it falls in)
        self._ustart = torch.tensor([shard_range(num_users, r, self.world)[0] for r in range(self.world)] + [num_users])
        self._istart = torch.tensor([shard_range(num_items, r, self.world)[0] for r in range(self.world)] + [num_items])
        gen = torch.Generator(device=self.device).manual_seed(int(getattr(cfg, "seed", 42)) * 1000 + self.rank)

        def shard(full, rows_total, lo, hi):
            if full is not None:
                return full[lo:hi].detach().to(self.device, F32).contiguous().clone()
            bound = math.sqrt(6.0 / (rows_total + d))
            return torch.rand(hi - lo, d, generator=gen, device=self.device).mul_(2 * bound).sub_(bound)

        self.U = shard(init[0] if init else None, num_users, self.u0, self.u1)
        self.V = shard(init[1] if init else None, num_items, self.i0, self.i1)
        z = lambda *s, dt=F32: torch.zeros(*s, device=self.device, dtype=dt)
        self._sh = {}
        for name, T, lo, hi in (("U", self.U, self.u0, self.u1), ("V", self.V, self.i0, self.i1)):
            rows = max(hi - lo, 1)
            self._sh[name] = dict(T=T, lo=lo, hi=hi, g=z(rows, d), flags=z(rows, dt=I32), counters=z(16, dt=I32),
                                  m=z(rows, d) if self.optimizer.needs_moments else None,
                                  v=z(rows, d) if self.optimizer.needs_moments else None, rows_list=None,
                                  last=z(rows, dt=I32) if self.optimizer.needs_moments else None)
        self._ustart, self._istart = self._ustart.to(self.device), self._istart.to(self.device)
        self._scal, self._n_scal = None, 0
        self._peers = [r for r in range(self.world) if r != self.rank]
        self._grank = (lambda r: dist.get_global_rank(self.group, r)) if (self.group is not None and self.world > 1) else (lambda r: r)
        self.err = z(1, dt=I32)
        self._cap = 0
        self._a2a, self._a2a_cap = None, 0
        self._slot_cache = None
        self.last_step_losses = None

    # ------------------------------------------------------------------------------------------
    def _buffers(self, B: int):
        if B > self._cap:
            per = (B + self.world - 1) // self.world
            d = self.d
            self._R = torch.empty(B, 3, d, device=self.device, dtype=F32)
            self._Gs = torch.zeros(per, 3, d, device=self.device, dtype=F32)
            self._G = torch.empty(per * self.world, 3, d, device=self.device, dtype=F32)
            self._sh["U"]["rows_list"] = torch.zeros(B, device=self.device, dtype=I32)
            self._sh["V"]["rows_list"] = torch.zeros(2 * B, device=self.device, dtype=I32)
            self._cap, self._per = B, per
        return self._R[:B], self._Gs, self._G, self._per

    # ------------------------------------------------------------------------------------------
    def _scalars(self, upto: int):
        """device table of the per-step Adam scalars, grown geometrically"""
        if upto >= self._n_scal:
            n = max(1024, 2 * (upto + 1))
            self._scal = torch.empty(2 * n, device=self.device, dtype=F32)
            self.kernels.adam_scalars(self.optimizer.opt_struct(1), n, self._scal)
            self._n_scal = n
        return self._scal, self._n_scal

    def _sparse_adam(self) -> bool:
        return self.optimizer.needs_moments and self.adam_mode == "sparse"

    def flush(self) -> None:
        """Sparse Adam only: bring every local row up to the current step (before the tables are read)."""
        if not self._sparse_adam() or self.optimizer.step_count == 0:
            return
        scal, n_scal = self._scalars(self.optimizer.step_count)
        opt = self.optimizer.opt_struct(self.optimizer.step_count)
        for name in ("U", "V"):
            self.kernels.step_sparse_adam(self._sh[name], opt, scal, n_scal, 0, self.d, flush=True)

    def _p2p(self, send, send_counts, recv, recv_counts):
        """all-to-all-v of rows as one grouped send/recv (segments ordered by rank); NCCL and gloo both run it"""
        me, so, ro, ops_ = self.rank, 0, 0, []
        for r in range(self.world):
            ns, nr = int(send_counts[r]), int(recv_counts[r])
            if r == me:
                if ns:
                    recv[ro:ro + nr].copy_(send[so:so + ns])
            else:
                if ns:
                    ops_.append(dist.P2POp(dist.isend, send[so:so + ns], self._grank(r), self.group))
                if nr:
                    ops_.append(dist.P2POp(dist.irecv, recv[ro:ro + nr], self._grank(r), self.group))
            so, ro = so + ns, ro + nr
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()

    def _train_step_a2a(self, uid, pos, neg, loss_acc) -> None:
        k, d, N, me, dev = self.kernels, self.d, self.world, self.rank, self.device
        B = int(uid.numel())
        S = (B + N - 1) // N
        b0, b1 = min(me * S, B), min(me * S + S, B)
        bad = ((uid < 0) | (uid >= self.num_users) | (pos < 0) | (pos >= self.num_items) | (neg < 0) | (neg >= self.num_items)).any()
        self.err |= bad.to(I32)
        uid, pos, neg = uid.clamp(0, self.num_users - 1), pos.clamp(0, self.num_items - 1), neg.clamp(0, self.num_items - 1)
        # ---- 1. plan: slot = 3 * triple + role; owner and owner-local row of every slot. One host round trip per step: the
        # world x world count matrix (who sends how many rows to whom) plus this rank's per-table counts.
        ou = torch.bucketize(uid, self._ustart[1:], right=True)
        op = torch.bucketize(pos, self._istart[1:], right=True)
        on = torch.bucketize(neg, self._istart[1:], right=True)
        owner = torch.stack((ou, op, on), 1).view(-1)
        lrow = torch.stack((uid - self._ustart[ou], pos - self._istart[op], neg - self._istart[on]), 1).view(-1)
        if self._slot_cache is None or self._slot_cache[0] != (B, S):
            slot = torch.arange(3 * B, device=dev)
            self._slot_cache = ((B, S), (slot // 3) // S, (slot % 3 != 0).to(I64))      # requester and table (0 user, 1 item) of a slot
        requester, table = self._slot_cache[1], self._slot_cache[2]
        if N == 1:                                   # one rank owns and computes everything: no plan, no host round trip
            send_counts = recv_counts = [3 * B]
            n_send = n_slice = 3 * B
            n_u, n_v = B, 2 * B
            my_sel, my_row, order = table.to(I32), lrow.to(I32), None
        else:
            is_mine = owner == me
            cm = torch.bincount(requester * N + owner, minlength=N * N)                 # [requester, owner] counts
            n_item_mine = (is_mine & (table == 1)).sum()
            host = torch.cat((cm, n_item_mine.view(1))).cpu()
            cm_h = host[: N * N].view(N, N)
            send_counts, recv_counts = cm_h[:, me].tolist(), cm_h[me, :].tolist()
            n_send, n_slice = int(sum(send_counts)), 3 * (b1 - b0)
            n_v = int(host[-1])
            n_u = n_send - n_v
            # slots I own, in slot (= requester, slot) order: a stable sort that moves them to the front
            mine = torch.argsort(~is_mine, stable=True)[:n_send]
            my_sel = table[mine].to(I32)
            my_row = lrow[mine].to(I32)
            order = torch.argsort(owner[3 * b0: 3 * b1], stable=True)                    # my slice's slots grouped by owner
        if self._a2a_cap < max(n_send, n_slice, 1):
            cap = max(n_send, n_slice, 1) * 5 // 4
            self._a2a = [torch.empty(cap, d, device=dev, dtype=F32) for _ in range(4)]
            self._a2a_cap = cap
        packed, rbuf, R, G = (t[:n] for t, n in zip(self._a2a, (n_send, n_slice, n_slice, n_slice)))
        su, sv = self._sh["U"], self._sh["V"]
        opt = self.optimizer.opt_struct(self.optimizer.step_count + 1)
        dense = (opt.kind != _cabi.YR_OPT_SGD) or (opt.weight_decay != 0.0)
        sparse_adam = self._sparse_adam()
        listed = (not dense) or sparse_adam
        # the slots I own, grouped by (table, table row) with ONE stable sort (equal rows keep slot order) — used twice: sparse
        # Adam brings the rows up to date before they are read, and the ordered accumulate sums every segment left to right
        key_sorted, perm = torch.sort(my_sel.to(I64) * (1 << 32) + my_row.to(I64), stable=True)
        rows_all = (key_sorted & 0xFFFFFFFF).to(I32)
        src_all = perm.to(I32)
        groups = []
        for s_, lo_, n_ in ((su, 0, n_u), (sv, n_u, n_v)):
            groups.append((s_, n_, rows_all[lo_: lo_ + n_].contiguous(), src_all[lo_: lo_ + n_].contiguous()))
            if sparse_adam and n_:
                scal, n_scal = self._scalars(opt.step)
                k.catch_up(s_, opt, scal, n_scal, groups[-1][2], d)
        # ---- 2./3. owners pack the rows, requesters receive them grouped by owner, then put them in slot order
        k.gather_local(su, sv, my_sel, my_row, packed)
        if N > 1:
            self._p2p(packed, send_counts, rbuf, recv_counts)
            R.index_copy_(0, order, rbuf)
        else:
            R = packed                                               # already in slot order
        # ---- 4. loss + gradient rows of my slice
        if b1 > b0:
            k.rows_grad_slice(R.view(-1, 3, d), B, b0, b1, G.view(-1, 3, d), loss_acc)
        # ---- 5. gradient rows back to the owners (received in requester, slot order = the order of `mine`)
        if N > 1:
            gsend = G.index_select(0, order)
            grecv = self._a2a[0][:n_send]                                    # the packed rows are no longer needed
            self._p2p(gsend, recv_counts, grecv, send_counts)
        else:
            grecv = G
        # ---- 6./7. ordered accumulate per table, one optimizer step per table
        for s_, n_, rows_sorted, src in groups:
            n_t = n_ if listed else 0
            if n_:
                if s_["rows_list"] is None or s_["rows_list"].numel() < n_:
                    s_["rows_list"] = torch.zeros(n_ * 5 // 4 + 16, device=dev, dtype=I32)
                k.accumulate_sorted(s_, opt, rows_sorted, src, grecv, listed)
            if s_["rows_list"] is None:
                s_["rows_list"] = torch.zeros(16, device=dev, dtype=I32)
            if sparse_adam:
                scal, n_scal = self._scalars(opt.step)
                k.step_sparse_adam(s_, opt, scal, n_scal, max(n_t, 1), d)
            else:
                k.step(s_, opt, max(n_t, 1), d)
        self.optimizer.step_count += 1

    def train_step(self, uid: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, loss_acc: torch.Tensor) -> None:
        """uid/pos/neg: int64 tensors on self.device, identical on every rank. loss_acc: double[1], += this rank's
        partial sum of -logsigmoid terms (reduced over ranks by train())."""
        if self.exchange == "all_to_all":
            return self._train_step_a2a(uid, pos, neg, loss_acc)
        k = self.kernels
        B, d = int(uid.numel()), self.d
        R, Gs, G, per = self._buffers(B)
        su, sv = self._sh["U"], self._sh["V"]
        k.gather_rows(su, self.num_users, uid, R, 0, self.err)
        k.gather_rows(sv, self.num_items, pos, R, 1, self.err)
        k.gather_rows(sv, self.num_items, neg, R, 2, self.err)
        if self.world > 1:
            dist.all_reduce(R, op=dist.ReduceOp.SUM, group=self.group)
        b0 = min(self.rank * per, B)
        b1 = min(b0 + per, B)
        k.rows_grad(R, B, b0, b1, Gs, loss_acc)
        if self.world > 1:
            dist.all_gather_into_tensor(G, Gs, group=self.group)   # G[r * per + j] = triple r * per + j
        else:
            G = Gs
        opt = self.optimizer.opt_struct(self.optimizer.step_count + 1)
        k.accumulate(su, opt, uid, G, 0)
        k.accumulate(sv, opt, pos, G, 1)
        k.accumulate(sv, opt, neg, G, 2)
        k.step(su, opt, B, d)
        k.step(sv, opt, 2 * B, d)
        self.optimizer.step_count += 1

    def train(self, batches) -> float:
        """Same contract as MFTrainer.train: returns the SUM of batch-mean losses (quirk Q1)."""
        accs, sizes = [], []
        for data in batches:
            u, p, n = (data[k].to(self.device, I64, non_blocking=True).contiguous() for k in ("user_id", "pos_item", "neg_item"))
            acc = torch.zeros(1, device=self.device, dtype=F64)
            self.train_step(u, p, n, acc)
            accs.append(acc)
            sizes.append(int(u.numel()))
        if not accs:
            return 0
        tot = torch.cat(accs)
        if self.world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        means = (tot / torch.tensor(sizes, device=self.device, dtype=F64)).to(F32)
        self.last_step_losses = means
        ops._raise_if_err(self.err, "ShardedMFTrainer.train")
        return float(means.to(F64).sum().item())

    def traffic_model(self, B: int) -> dict:
        """Algorithmic bytes of one step (SURVEY.md 8(d) per-triple figures at this width) and what crosses NVLink per GPU."""
        d, N = self.d, self.world
        row = d * 4
        if not self.optimizer.needs_moments:
            per_triple = 3 * 8 + 3 * row * 2                                   # ids + 3 rows read + 3 rows written
            what = "plain SGD: 3 rows read + written per triple"
        elif self.adam_mode == "sparse":
            per_triple = 3 * 8 + 3 * 4 + 3 * 3 * row * 2                       # ids, last-step words, (p, m, v) read + written
            what = "sparse-traffic Adam: (p, m, v) of 3 rows read + written per triple"
        else:
            rows_local = (self.u1 - self.u0) + (self.i1 - self.i0)
            per_triple = 3 * 8 + 3 * row * 2 + rows_local * row * 6 / max(B / N, 1)
            what = "dense Adam sweep: p, m, v of every local row read + written per step"
        S = (B + N - 1) // N
        nvl = 0 if N == 1 else (2 * 3 * S * row * (N - 1) // N if self.exchange == "all_to_all" else 2 * 3 * B * row * (N - 1) // N)
        return {"alg_bytes_per_triple": per_triple, "alg_model": what, "exchange": self.exchange,
                "nvlink_bytes_per_step_per_gpu": int(nvl), "adam_mode": self.adam_mode if self.optimizer.needs_moments else None}

    def gather_tables(self):
        """Full (U, V) on every rank (tests / evaluation of modest sizes)."""
        self.flush()
        if self.world == 1:
            return self.U, self.V
        out = []
        for T, total in ((self.U, self.num_users), (self.V, self.num_items)):
            per = (total + self.world - 1) // self.world + 1
            pad = torch.zeros(per, self.d, device=self.device, dtype=F32)
            pad[: T.shape[0]] = T
            allp = torch.empty(self.world * per, self.d, device=self.device, dtype=F32)
            dist.all_gather_into_tensor(allp, pad, group=self.group)
            parts = []
            for r in range(self.world):
                lo, hi = shard_range(total, r, self.world)
                parts.append(allp[r * per: r * per + (hi - lo)])
            out.append(torch.cat(parts))
        return out[0], out[1]
