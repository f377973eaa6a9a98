"""Row-sharded NGCF training across GPUs (BASELINE config 5): 1-D row partition of the Laplacian,
of the node-embedding table and of its optimizer state; the 2·num_orders d x d weights are replicated.

Rank r owns node rows [r*per, min((r+1)*per, N)), per = ceil(N / world) (so the all-gathered operand [world*per x d]
is indexed by GLOBAL node id), and holds rows [r0, r1) of L and of L^T as CSR blocks with global column ids.
One step (reference semantics: models/ngcf.py:30-72, trainers/ngcf_trainer.py:102-115):

    forward, per layer l :  X = all_gather(E_l)            NCCL, N*d*4 bytes
                            LE_l = L[r0:r1, :] X           yr_spmm_csr on the local row block
                            E_{l+1} = dense(E_l, LE_l)     yr_ngcf_dense_fwd on local rows
    tail                 :  owner gather of rows u / U+pos / U+neg of every layer -> all_reduce (exact gather),
                            yr_bpr_rows_grad on the concatenated rows (every rank, all B triples: the loss and the row
                            gradients are replicated, no collective), yr_shard_accumulate scatters owned rows into G_l
    backward, per layer  :  yr_ngcf_dense_bwd on local rows -> T, G_l += ..., dW partial
                            X = all_gather(T);  G_l += L^T[r0:r1, :] X
                            all_reduce(dW1, dW2)
    update               :  yr_dense_opt_step on the local rows of E_0 and (identically on every rank) on the weights

The result equals the single-GPU NGCFTrainer up to fp32 summation order of dW (per-CTA partials, then ranks) and of
duplicate tail rows. No CPU product path: `device`/`kernels` exist so that tests/_dist_shard_worker.py can drive the
choreography under gloo with a CPU restatement of the kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .. import _cabi, ops
from ..data.graph import CSRMatrix, coo_to_csr
from .base_trainer import FusedOptimizer

I32, I64, F32, F64 = torch.int32, torch.int64, torch.float32, torch.float64
_SLOPE = 0.01


class CabiNgcfShardKernels:
    """Device-side pieces through the C-ABI (include/yelprec_b200.h)."""

    def __init__(self, device):
        self.device = device
        self.lib = _cabi.load()
        self._ws = None
        self.dense_mode = _cabi.YR_DENSE_TC_FWD

    def _st(self):
        return _cabi.stream_ptr(self.device)

    def make_csr(self, rowptr, col, val):
        return CSRMatrix(rowptr, col, val, self.device)

    def spmm(self, A, X, out, accumulate):
        ops.spmm_csr(A, X, out=out, accumulate=accumulate)

    def dense_fwd(self, E, LE, W1, W2, out):
        d = E.shape[1]
        _cabi.check(self.lib.yr_ngcf_dense_fwd(d, E.shape[0], _cabi.dptr(E, F32), _cabi.dptr(LE, F32), _cabi.dptr(W1, F32),
                                               _cabi.dptr(W2, F32), _SLOPE, _cabi.dptr(out, F32), self.dense_mode, self._st()),
                    "yr_ngcf_dense_fwd")

    def dense_bwd(self, E, LE, En, Gn, W1, W2, G, T, dW1, dW2):
        d = E.shape[1]
        nbytes = self.lib.yr_ngcf_layer_bwd_ws_bytes(d)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        p = _cabi.dptr
        _cabi.check(self.lib.yr_ngcf_dense_bwd(d, E.shape[0], p(E, F32), p(LE, F32), p(En, F32), p(Gn, F32), p(W1, F32),
                                               p(W2, F32), _SLOPE, p(G, F32), p(T, F32), p(dW1, F32), p(dW2, F32),
                                               p(self._ws), nbytes, self.dense_mode, self._st()), "yr_ngcf_dense_bwd")

    def gather_rows(self, T, lo, hi, total, ids, R, col_off, err):
        """R[:, col_off : col_off + d] = T[ids - lo] for owned ids, zeros elsewhere. R: [n x ld] fp32."""
        d = T.shape[1]
        _cabi.check(self.lib.yr_shard_gather_rows(_cabi.dptr(T, F32), lo, hi, total, d, _cabi.dptr(ids, I64), int(ids.numel()),
                                                  R.data_ptr() + col_off * 4, R.shape[1], _cabi.dptr(err), self._st()),
                    "yr_shard_gather_rows")

    def rows_grad(self, R, B, width, Gr, loss_acc):
        _cabi.check(self.lib.yr_bpr_rows_grad(_cabi.dptr(R, F32), width, B, 0, B, _cabi.dptr(Gr, F32), _cabi.dptr(loss_acc, F64),
                                              self._st()), "yr_bpr_rows_grad")

    def scatter_rows(self, G, lo, hi, ids, Gr, col_off, flags, scratch):
        """G[ids - lo] += Gr[:, col_off : col_off + d] for owned ids (duplicates summed)."""
        d = G.shape[1]
        p = _cabi.dptr
        st = _cabi.YrShardState(p(G, F32), None, None, p(G, F32), p(flags), p(scratch), p(scratch), lo, hi, d)
        opt = _cabi.make_opt("adam", 0.0, 0.0, 1)          # dense flavour: only marks flags, no row list
        _cabi.check(self.lib.yr_shard_accumulate(C.byref(st), C.byref(opt), p(ids, I64), int(ids.numel()),
                                                 Gr.data_ptr() + col_off * 4, Gr.shape[1], self._st()), "yr_shard_accumulate")

    def opt_step(self, p, g, m, v, opt):
        ops.dense_opt_step(p, g, m, v, opt)


class ShardedNGCFTrainer:
    def __init__(self, cfg, num_items: int, num_users: int, laplacian_matrix: torch.Tensor, init=None, group=None,
                 device=None, kernels=None):
        """`laplacian_matrix`: the reference's sparse COO [N x N] (every rank passes the same; each keeps its row block).
        `init`: optional dict with 'embedding.weight' [N x d], 'W1.l.weight', 'W2.l.weight' (tests); otherwise
        torch.manual_seed(cfg.seed)-driven N(0,1) / kaiming-uniform like the reference (models/ngcf.py:9-23)."""
        self.cfg, self.nI, self.nU = cfg, int(num_items), int(num_users)
        self.N = self.nU + self.nI
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.k = kernels if kernels is not None else CabiNgcfShardKernels(self.device)
        self.d, self.n_layers = int(cfg.embed_size), int(cfg.num_orders)
        self.width = (self.n_layers + 1) * self.d
        if self.width not in (32, 64, 128, 256):
            raise NotImplementedError(f"concatenated width {self.width} not in (32, 64, 128, 256)")
        self.optimizer = FusedOptimizer(cfg.optimizer, cfg.lr, cfg.weight_decay)
        self.per = (self.N + self.world - 1) // self.world
        self.r0 = min(self.rank * self.per, self.N)
        self.r1 = min(self.r0 + self.per, self.N)
        self.n_loc = self.r1 - self.r0
        if self.n_loc <= 0:
            raise ValueError("more ranks than node rows")
        # ---- row blocks of L and of L^T (global column ids)
        Lc = laplacian_matrix.detach().cpu().coalesce()
        idx, val = Lc.indices().numpy(), Lc.values().numpy().astype(np.float32)
        self.A = self._row_block(idx[0], idx[1], val)
        self.AT = self._row_block(idx[1], idx[0], val)
        # ---- parameters
        dev, d, n_loc = self.device, self.d, self.n_loc
        if init is not None:
            E0 = init["embedding.weight"][self.r0:self.r1]
            W1 = [init[f"W1.{l}.weight"] for l in range(self.n_layers)]
            W2 = [init[f"W2.{l}.weight"] for l in range(self.n_layers)]
        else:
            g = torch.Generator().manual_seed(int(getattr(cfg, "seed", 42)))
            E0 = torch.randn(self.N, d, generator=g)[self.r0:self.r1]
            bound = 1.0 / np.sqrt(d)                                  # nn.Linear default: kaiming_uniform(a=sqrt(5))
            W1 = [(torch.rand(d, d, generator=g) * 2 - 1) * bound for _ in range(self.n_layers)]
            W2 = [(torch.rand(d, d, generator=g) * 2 - 1) * bound for _ in range(self.n_layers)]
        f = lambda t: t.detach().to(dev, F32).contiguous().clone()
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        self.E: List[torch.Tensor] = [f(E0)] + [z(n_loc, d) for _ in range(self.n_layers)]
        self.LE = [z(n_loc, d) for _ in range(self.n_layers)]
        self.G = [z(n_loc, d) for _ in range(self.n_layers + 1)]
        self.T = z(self.per, d)                                       # padded: it is an all_gather input
        self.X = z(self.world * self.per, d)                          # all-gathered operand, global row ids
        self.Epad = z(self.per, d)
        self.W1, self.W2 = [f(w) for w in W1], [f(w) for w in W2]
        self.dW = z(2 * self.n_layers, d, d)
        mom = self.optimizer.needs_moments
        self.mE, self.vE = (z(n_loc, d), z(n_loc, d)) if mom else (None, None)
        self.mW = z(2 * self.n_layers, d, d) if mom else None
        self.vW = z(2 * self.n_layers, d, d) if mom else None
        self.flags = z(max(n_loc, 1), dt=I32)
        self.scratch = z(16, dt=I32)
        self.err = z(1, dt=I32)
        self._cap = 0
        self.last_step_losses = None

    def _row_block(self, rows, cols, vals):
        sel = (rows >= self.r0) & (rows < self.r1)
        rp, ci, va = coo_to_csr(rows[sel] - self.r0, cols[sel], vals[sel], self.n_loc)
        return self.k.make_csr(rp, ci, va)

    # ------------------------------------------------------------------------------------------
    def _all_gather(self, local: torch.Tensor) -> torch.Tensor:
        """local: [n_loc x d] (or the padded [per x d] T) -> self.X [world*per x d] indexed by global node id."""
        if self.world == 1:
            if local.shape[0] == self.per:
                return local
            self.Epad[: self.n_loc] = local
            return self.Epad
        src = local
        if local.shape[0] != self.per:
            self.Epad[: self.n_loc] = local
            src = self.Epad
        dist.all_gather_into_tensor(self.X, src, group=self.group)
        return self.X

    def propagate(self):
        k = self.k
        for l in range(self.n_layers):
            X = self._all_gather(self.E[l])
            k.spmm(self.A, X, self.LE[l], False)
            k.dense_fwd(self.E[l], self.LE[l], self.W1[l], self.W2[l], self.E[l + 1])
        return self.E

    def train_step(self, uid, pos, neg, loss_acc) -> None:
        """uid/pos/neg int64 on self.device, identical on every rank; loss_acc double[1] += sum of -logsigmoid terms
        (replicated: every rank accumulates the same value)."""
        k, d, B, W = self.k, self.d, int(uid.numel()), self.width
        if B > self._cap:
            self._R = torch.empty(3 * B, W, device=self.device, dtype=F32)
            self._Gr = torch.empty(3 * B, W, device=self.device, dtype=F32)
            self._cap = B
        R, Gr = self._R[: 3 * B], self._Gr[: 3 * B]
        self.propagate()
        # ---- tail: R viewed as [B x 3 x W]; row (b, which) = b*3 + which -> ids interleaved the same way
        ids = torch.stack((uid, pos + self.nU, neg + self.nU), dim=1).reshape(-1).contiguous()
        bad = ((uid < 0) | (uid >= self.nU) | (pos < 0) | (pos >= self.nI) | (neg < 0) | (neg >= self.nI)).any()
        self.err |= bad.to(I32)
        ids = ids.clamp(0, self.N - 1)
        for l in range(self.n_layers + 1):
            k.gather_rows(self.E[l], self.r0, self.r1, self.N, ids, R, l * d, self.err)
        if self.world > 1:
            dist.all_reduce(R, op=dist.ReduceOp.SUM, group=self.group)
        k.rows_grad(R, B, W, Gr, loss_acc)
        for g in self.G:
            g.zero_()
        for l in range(self.n_layers + 1):
            k.scatter_rows(self.G[l], self.r0, self.r1, ids, Gr, l * d, self.flags, self.scratch)
        # ---- backward through the layers
        for l in reversed(range(self.n_layers)):
            Tl = self.T[: self.n_loc]
            k.dense_bwd(self.E[l], self.LE[l], self.E[l + 1], self.G[l + 1], self.W1[l], self.W2[l], self.G[l], Tl,
                        self.dW[l], self.dW[self.n_layers + l])
            X = self._all_gather(self.T)
            k.spmm(self.AT, X, self.G[l], True)
        if self.world > 1:
            dist.all_reduce(self.dW, op=dist.ReduceOp.SUM, group=self.group)
        # ---- optimizer: parameter order embedding, W1.*, W2.* (nn.Module.parameters()); all tensors share the step count
        opt = self.optimizer.opt_struct(self.optimizer.step_count + 1)
        k.opt_step(self.E[0], self.G[0], self.mE, self.vE, opt)
        for i in range(2 * self.n_layers):
            Wt = self.W1[i] if i < self.n_layers else self.W2[i - self.n_layers]
            k.opt_step(Wt, self.dW[i], self.mW[i] if self.mW is not None else None,
                       self.vW[i] if self.vW is not None else None, opt)
        self.optimizer.step_count += 1

    def train(self, batches) -> float:
        """Same contract as NGCFTrainer.train: the SUM of batch-mean losses (quirk Q1)."""
        accs, sizes = [], []
        for data in batches:
            u, p, n = (data[key].to(self.device, I64, non_blocking=True).contiguous() for key in ("user_id", "pos_item", "neg_item"))
            acc = torch.zeros(1, device=self.device, dtype=F64)
            self.train_step(u, p, n, acc)
            accs.append(acc)
            sizes.append(int(u.numel()))
        if not accs:
            return 0
        means = (torch.cat(accs) / torch.tensor(sizes, device=self.device, dtype=F64)).to(F32)
        self.last_step_losses = means
        ops._raise_if_err(self.err, "ShardedNGCFTrainer.train")
        return float(means.to(F64).sum().item())

    def gather_embedding(self) -> torch.Tensor:
        """Full E_0 [N x d] on every rank (tests)."""
        return self._all_gather(self.E[0])[: self.N].clone()
