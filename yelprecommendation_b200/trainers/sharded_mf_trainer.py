"""Row-sharded BPR-MF training across GPUs (BASELINE config 5: 10 M users x 2 M items, d = 128) — SURVEY.md §8(e).

One process per GPU (torch.distributed, NCCL over NVLink). Rank r owns a contiguous block of user rows and of item
rows together with their optimizer state; nothing else is replicated. Every rank sees the same batch of triples.
One step:

    1. yr_shard_gather_rows   owners write the rows they hold of (u, pos, neg) into R [B x 3 x d], zeros elsewhere
    2. all_reduce(R, SUM)     = exact gather of all 3B rows on every rank (one non-zero contributor per row)
    3. yr_bpr_rows_grad       rank r computes loss + gradient rows for ITS slice of the triples
    4. all_gather(G slices)   every rank gets all 3B gradient rows
    5. yr_shard_accumulate x3 + yr_shard_step x2   owners sum the rows of their ids and step their shard once

The arithmetic per row is the single-GPU fused trainer's (trainers/mf_trainer.py of the reference, :104-114), so the
gathered tables equal MFTrainer's after the same triples up to the fp32 order in which duplicate rows are summed.
With world_size == 1 the collectives are no-ops and the class runs as is (used as the 1-GPU parity reference, since the
reference itself cannot run this configuration).
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


class ShardedMFTrainer:
    def __init__(self, cfg, num_items: int, num_users: int, init=None, group=None, device=None, kernels=None):
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
        self.u0, self.u1 = shard_range(num_users, self.rank, self.world)
        self.i0, self.i1 = shard_range(num_items, self.rank, self.world)
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
                                  v=z(rows, d) if self.optimizer.needs_moments else None, rows_list=None)
        self.err = z(1, dt=I32)
        self._cap = 0
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

    def train_step(self, uid: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, loss_acc: torch.Tensor) -> None:
        """uid/pos/neg: int64 tensors on self.device, identical on every rank. loss_acc: double[1], += this rank's
        partial sum of -logsigmoid terms (reduced over ranks by train())."""
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

    def gather_tables(self):
        """Full (U, V) on every rank (tests / evaluation of modest sizes)."""
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
