"""Row-sharded NGCF training across GPUs (BASELINE config 5): 1-D row partition of the Laplacian,
of the node-embedding table and of its optimizer state; the 2·num_orders d x d weights are replicated.

Partition (data/scaled.py::ShardLayout): rank k owns a contiguous block of USERS and a contiguous block of ITEMS — `per`
rows in all — so every rank holds about nnz / world non-zeros (a block partition of the node ids alone would give the
item-owning ranks several times the non-zeros of the others). The all-gathered operand X [world * per x d] is indexed by
position = rank * per + local row; users and items map monotonically, so every CSR row keeps the column order it has
on one GPU and the SpMM results are bit-identical at any world size.
One step (reference semantics: models/ngcf.py:30-72, trainers/ngcf_trainer.py:102-115), local rows cut into P row panels:

    forward, per layer l :  for panel p:  LE_l[p] = L[p, :] X_l          yr_spmm_csr on the panel's rows
                                          E_{l+1}[p] = dense(E_l, LE_l)  yr_ngcf_dense_fwd
                                          exchange(E_{l+1}[p])           grouped NCCL send/recv into every peer's X_{l+1},
                                                                          issued asynchronously: it runs underneath panel p+1
    tail                 :  owner gather of rows u / pos / neg of every layer -> all_reduce (exact gather),
                            yr_bpr_rows_grad on the concatenated rows (every rank, all B triples: the loss and the row
                            gradients are replicated, no further collective), yr_shard_accumulate_sorted sums owned rows into G_l in batch order (no atomics)
    backward, per layer  :  for panel p:  yr_ngcf_dense_bwd -> T[p], G_l[p] += ..., dW partial;  exchange(T[p]) (async)
                            G_l += L^T[block, :] X_T;   all_reduce(dW1, dW2) once per step
    update               :  yr_dense_opt_step on the local rows of E_0 and (identically on every rank) on the weights

Where the whole operand is exchanged right before it is consumed — layer 0 (it follows the optimizer step) and every
backward layer (T) — the SpMM is cut by COLUMN panel instead: the entries whose column is a row of panel p on some rank; the
SpMM over column panel p needs only exchange round p and runs underneath round p+1. Only one round (1 / P of N * d * 4
bytes) per layer stays exposed. It needs panels that each hold users AND items (YR_SHARD_INTERLEAVE=1, data/scaled.py); measured
on 8 x B200 the pair gains nothing over row panels alone (133.7 vs 131.0 ms per step), so both are off by default. The graph arrives either as the reference's torch sparse COO Laplacian (row blocks
cut on the device, any weights) or as a data.scaled.ScaledGraph (config 5: generated and normalised on the device).
No CPU product path: `device` / `kernels` exist so that tests/_dist_shard_worker.py can drive the choreography under gloo
with a CPU restatement of the kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from .. import _cabi, ops
from ..data.graph import CSRMatrix
from ..data.scaled import ScaledGraph, ShardLayout, shard_laplacian, shard_laplacian_from_coo
from .base_trainer import FusedOptimizer

I32, I64, F32, F64 = torch.int32, torch.int64, torch.float32, torch.float64
_SLOPE = 0.01


class CabiNgcfShardKernels:
    """Device-side pieces through the C-ABI (include/yelprec_b200.h)."""

    def __init__(self, device):
        self.device = device
        self.lib = _cabi.load()
        self._ws = None
        self.dense_mode = _cabi.YR_DENSE_TC
        self.reserve_sms = 0

    def _st(self):
        return _cabi.stream_ptr(self.device)

    def make_csr(self, rowptr, col, val):
        """rowptr: absolute offsets into col / val (a row panel shares the block's storage)"""
        m = CSRMatrix(rowptr, col, val, self.device)
        m.reserve_sms = self.reserve_sms          # yr_csr.reserve_sms: SMs left empty for the exchange's NCCL kernels
        return m

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

    def scatter_rows_sorted(self, G, rows_sorted, src, Gr, col_off, flags, scratch):
        """Ordered form of scatter_rows on a G that has just been cleared: G[r] = sum, in batch order, of the gradient rows
        Gr[src[j], col_off : col_off + d] with rows_sorted[j] == r (rows_sorted: stable-sorted local rows, < 0 = not owned).
        No floating-point atomics: bit-identical run to run."""
        d = G.shape[1]
        p = _cabi.dptr
        st = _cabi.YrShardState(p(G, F32), None, None, p(G, F32), p(flags), p(scratch), p(scratch), 0, G.shape[0], d)
        opt = _cabi.make_opt("adam", 0.0, 0.0, 1)
        _cabi.check(self.lib.yr_shard_accumulate_sorted(C.byref(st), C.byref(opt), p(rows_sorted, I32), p(src, I32),
                                                        int(rows_sorted.numel()), Gr.data_ptr() + col_off * 4, Gr.shape[1], 0,
                                                        self._st()), "yr_shard_accumulate_sorted")

    def opt_step(self, p, g, m, v, opt):
        ops.dense_opt_step(p, g, m, v, opt)


def _default_exchange(world: int) -> str:
    return "symm" if 1 < world <= 4 else "p2p"


class ShardedNGCFTrainer:
    def __init__(self, cfg, num_items: int, num_users: int, laplacian_matrix, init=None, group=None,
                 device=None, kernels=None, n_panels: int = None, solo: bool = False):
        """`laplacian_matrix`: the reference's sparse COO [N x N] (every rank passes the same; each keeps its row block) or
        a data.scaled.ScaledGraph. `init`: optional dict with 'embedding.weight' [N x d] (reference node order),
        'W1.l.weight', 'W2.l.weight' (tests); otherwise N(0,1) / kaiming-uniform like the reference (models/ngcf.py:9-23),
        generated on the device per rank (a function of cfg.seed and the rank)."""
        self.cfg, self.nI, self.nU = cfg, int(num_items), int(num_users)
        self.N = self.nU + self.nI
        self.group = group
        multi = dist.is_initialized() and not solo          # solo: this process trains the whole model alone (parity runs)
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # YR_SHARD_HIPRIO=1: the NCCL panel rounds go through a process group of their own with high-priority streams. Off by
        # default — it changes nothing (measured on 2 GPUs: 285.7 vs 287.1 ms per step): NCCL's 640-thread CTAs need an EMPTY SM
        # and do not start while a grid of small CTAs keeps refilling every SM, whatever their priority (notes/README.md).
        self.xgroup = group
        if (multi and self.world > 1 and kernels is None and dist.get_backend(group) == "nccl"
                and os.environ.get("YR_SHARD_EXCHANGE", _default_exchange(self.world)) == "p2p"
                and int(os.environ.get("YR_SHARD_HIPRIO", "0")) != 0):
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            self.xgroup = dist.new_group(ranks=dist.get_process_group_ranks(group if group is not None else dist.group.WORLD),
                                         backend="nccl", pg_options=opts)
        self.k = kernels if kernels is not None else CabiNgcfShardKernels(self.device)
        if kernels is None:                     # yr_dense_mode of the d x d transforms (default: tensor cores for both passes)
            self.k.dense_mode = int(getattr(cfg, "ngcf_dense_mode", _cabi.YR_DENSE_TC))
            if self.k.dense_mode not in (0, 1, 2):
                raise ValueError(f"ngcf_dense_mode {self.k.dense_mode} not in (0, 1, 2)")
            # YR_SHARD_RESERVE_SMS=R (world > 1): the SpMM and the dense transforms run on SM count - R persistent one-per-SM
            # CTAs and leave R SMs empty for the NCCL kernels of the panel exchange (yr_csr.reserve_sms, YR_DENSE_RESERVE);
            # NCCL is told to use at most R CTAs unless NCCL_MAX_CTAS is already set.
            R = int(os.environ.get("YR_SHARD_RESERVE_SMS", "0")) if (multi and self.world > 1) else 0
            if R > 0:
                os.environ.setdefault("NCCL_MAX_CTAS", str(R))
                self.k.reserve_sms = R
                self.k.dense_mode |= (R & 0xff) << 8
        self.d, self.n_layers = int(cfg.embed_size), int(cfg.num_orders)
        self.width = (self.n_layers + 1) * self.d
        if self.width not in (32, 64, 128, 256, 512, 1024):
            raise NotImplementedError(f"concatenated width {self.width} not in (32, 64, 128, 256, 512, 1024)")
        self.optimizer = FusedOptimizer(cfg.optimizer, cfg.lr, cfg.weight_decay)
        P = int(n_panels if n_panels is not None else getattr(cfg, "shard_panels", int(os.environ.get("YR_SHARD_PANELS", "8")) if self.world > 1 else 1))
        # YR_SHARD_INTERLEAVE=1: every row panel holds a slice of the users and a slice of the items (what the column-panel
        # schedule needs to be balanced). Measured on 8 x B200 (profiles/README.md): the mixed panels cost the SpMM 15 % (the
        # user-vector gathers evict the hot item vectors from L2) and the column panels win back about as much — 133.7 ms per
        # step against 131.0 ms for the plain [users ; items] order with row panels only, which therefore stays the default.
        interleave = int(os.environ.get("YR_SHARD_INTERLEAVE", "0")) != 0 and P > 1
        self.layout = ShardLayout(self.nU, self.nI, self.world, max(1, P) if interleave else 1)
        self.per = self.layout.per
        self.lo, self.hi = self.rank * self.per, (self.rank + 1) * self.per          # positions of the local rows
        self.total = self.world * self.per
        # ---- row block of L (cut into panels) and of L^T, columns = positions in the gathered operand
        dev = self.device
        if isinstance(laplacian_matrix, ScaledGraph):
            rp, ci, va = shard_laplacian(laplacian_matrix, self.layout, self.rank)
            rpT, ciT, vaT = rp, ci, va                               # binary ratings: L is bit-wise symmetric
        else:
            rp, ci, va = shard_laplacian_from_coo(laplacian_matrix, self.layout, self.rank, dev)
            rpT, ciT, vaT = shard_laplacian_from_coo(laplacian_matrix, self.layout, self.rank, dev, transpose=True)
        self.nnz_local = int(ci.numel())
        if interleave:
            step = self.layout.pp                                    # a panel = a slice of the users + a slice of the items
        else:
            P = max(1, min(P, self.per))
            step = ((self.per + P - 1) // P + 127) // 128 * 128      # whole 128-row tiles of the dense kernels
        rp_h = rp.cpu()
        self.panels = [(a, min(a + step, self.per), self.k.make_csr(rp_h[a: min(a + step, self.per) + 1], ci, va))
                       for a in range(0, self.per, step)]
        self.AT = self.k.make_csr(rpT.cpu(), ciT, vaT) if (rpT is not rp or len(self.panels) > 1) else self.panels[0][2]
        # COLUMN panels (one per row panel of the gathered operand): the SpMM over column panel p only needs the rows
        # every rank sent in exchange round p, so it can run underneath round p+1 — used where the whole operand is
        # exchanged right before it is consumed (layer 0 after the optimizer step; every backward layer)
        # how a panel reaches the peers: 'symm' = copy-engine pushes through symmetric memory, 'p2p' = grouped NCCL send / recv,
        # 'allgather' / 'none' = experiments (one NCCL all-gather per layer / compute only). Measured at config 5
        # (profiles/README.md): 2 GPUs 269 ms (symm) vs 285 ms (p2p) per step — the pushes run underneath the SpMM, NCCL's SM
        # kernels queue behind it; 8 GPUs 123.4 vs 118.6 ms — seven concurrent copy streams per GPU reach 418 GB/s where NCCL
        # reaches 553 GB/s, which costs more than the overlap gains; 4 GPUs 165.5 vs 172.6 ms. Default: symm up to 4 ranks, p2p beyond.
        self._xmode = os.environ.get("YR_SHARD_EXCHANGE", _default_exchange(self.world) if kernels is None else "p2p")
        self.use_col_panels = (self.world > 1 and len(self.panels) > 1 and self._xmode not in ("allgather", "symm")
                               and int(os.environ.get("YR_SHARD_COLPANELS", "1" if interleave else "0")) != 0)
        self.colA = self.colAT = None
        if self.use_col_panels:
            self.colAT = self._column_panels(rpT, ciT, vaT, step)
            self.colA = self.colAT if rpT is rp else self._column_panels(rp, ci, va, step)
        # ---- parameters
        d, per = self.d, self.per
        f = lambda t: t.detach().to(dev, F32).contiguous().clone()
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        if init is not None:
            E0 = self.layout.local_rows(self.rank, init["embedding.weight"].to(F32)).to(dev)
            W1 = [init[f"W1.{l}.weight"] for l in range(self.n_layers)]
            W2 = [init[f"W2.{l}.weight"] for l in range(self.n_layers)]
        else:
            seed = int(getattr(cfg, "seed", 42))
            g = torch.Generator(device=dev).manual_seed(seed * 1000003 + self.rank)
            E0 = torch.randn(per, d, generator=g, device=dev, dtype=F32)
            gw = torch.Generator().manual_seed(seed)                  # the weights are replicated: same stream on every rank
            bound = 1.0 / np.sqrt(d)                                  # nn.Linear default: kaiming_uniform(a=sqrt(5))
            W1 = [(torch.rand(d, d, generator=gw) * 2 - 1) * bound for _ in range(self.n_layers)]
            W2 = [(torch.rand(d, d, generator=gw) * 2 - 1) * bound for _ in range(self.n_layers)]
        self.E: List[torch.Tensor] = [f(E0)] + [z(per, d) for _ in range(self.n_layers)]
        self.LE = [z(per, d) for _ in range(self.n_layers)]
        self.G = [z(per, d) for _ in range(self.n_layers + 1)]
        self.T = z(per, d)
        # gathered operands: two buffers, so that the exchange of layer l+1 can land while layer l's SpMM still reads
        self.X = [z(self.total, d) for _ in range(2)] if self.world > 1 else None
        self._pending = [[], []]
        self._symm = None
        if self.world > 1 and self._xmode == "symm":
            ok = torch.ones(1, device=dev, dtype=I32)
            try:
                self._init_symm(d)
            except Exception as ex:                    # no symmetric-memory support on this system: NCCL send / recv instead
                ok.zero_()
                self._symm_error = repr(ex)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # every rank takes the same path
            if int(ok.item()) == 0:
                self._symm, self._xmode = None, "p2p"
                self.X = [z(self.total, d) for _ in range(2)]
        self.W1, self.W2 = [f(w) for w in W1], [f(w) for w in W2]
        self.dW = z(2 * self.n_layers, d, d)
        self.dWp = z(len(self.panels), 2, d, d)
        mom = self.optimizer.needs_moments
        self.mE, self.vE = (z(per, d), z(per, d)) if mom else (None, None)
        self.mW = z(2 * self.n_layers, d, d) if mom else None
        self.vW = z(2 * self.n_layers, d, d) if mom else None
        self.flags = z(max(per, 1), dt=I32)
        self.scratch = z(16, dt=I32)
        self.err = z(1, dt=I32)
        self._cap = 0
        self.last_step_losses = None
        self._peers = [r for r in range(self.world) if r != self.rank]
        self._grank = (lambda r: dist.get_global_rank(self.group, r)) if (self.group is not None and self.world > 1) else (lambda r: r)

    def _column_panels(self, rp, ci, va, step):
        """The row block cut by COLUMN panel: panel p holds the entries whose column is a row of panel p on some rank
        ((col mod per) in [a_p, b_p)), all local rows, columns ascending inside a row. A stable sort by panel keeps the
        (row, column) order inside every panel; the flat cumulative count is every panel's rowptr (absolute offsets into the
        re-ordered col / val, which the panels share)."""
        per, P = self.per, len(self.panels)
        pan = (ci.to(I64) % per) // step
        order = torch.argsort(pan, stable=True)
        rows = torch.repeat_interleave(torch.arange(per, device=ci.device), (rp[1:] - rp[:-1]).to(I64))
        key = pan[order] * per + rows[order]
        del pan, rows
        ci_p, va_p = ci[order].contiguous(), va[order].contiguous()
        del order
        ptr = torch.zeros(P * per + 1, dtype=I64, device=ci.device)
        ptr[1:] = torch.cumsum(torch.bincount(key, minlength=P * per), 0)
        del key
        ptr_h = ptr.to(I32).cpu()
        return [self.k.make_csr(ptr_h[p * per: (p + 1) * per + 1], ci_p, va_p) for p in range(P)]

    # ------------------------------------------------------------------------------------------
    def _init_symm(self, d: int) -> None:
        """YR_SHARD_EXCHANGE=symm: the gathered operands live in symmetric memory (torch.distributed._symmetric_memory: every
        rank maps every peer's buffer) and a panel is PUSHED into the peers' buffers with plain device-to-device copies, one
        stream per peer — the copy engines move the data over NVLink, no SM is involved, so the exchange really runs underneath
        the SpMM (NCCL's send / recv kernels need SMs of their own: measured, they queued behind the compute kernels and a
        step cost compute + exchange). A layer's exchange ends with one put_signal per peer behind the copies; the consumer
        waits for the peers' signals on its compute stream."""
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        self._symm = []
        for b in range(2):
            t = symm_mem.empty((self.total, d), dtype=F32, device=self.device)
            t.zero_()
            h = symm_mem.rendezvous(t, group=grp)
            views = {r: h.get_buffer(r, (self.total, d), F32, 0) for r in self._peers_list()}
            self.X[b] = t
            self._symm.append((h, views))
        self._pstreams = {r: torch.cuda.Stream(device=self.device) for r in self._peers_list()}
        self._xchan = [[], []]                    # signal channels of the posted, not yet awaited, exchanges per buffer
        self._xseq = 0
        self._copied = []                         # events: my outgoing copies of the current exchange

    def _peers_list(self):
        return [(self.rank + i) % self.world for i in range(1, self.world)]

    def _post_symm(self, src: torch.Tensor, a: int, b: int, buf: int) -> None:
        h, views = self._symm[buf]
        cur = torch.cuda.current_stream(self.device)
        self.X[buf][self.lo + a: self.lo + b].copy_(src[a:b])
        ready = torch.cuda.Event()
        ready.record(cur)
        first, last = a == 0, b >= self.per
        if first:
            self._xchan[buf].append(self._xseq % 8)
            self._xseq += 1
        ch = self._xchan[buf][-1]
        for r, st in self._pstreams.items():
            st.wait_event(ready)
            with torch.cuda.stream(st):
                views[r][self.lo + a: self.lo + b].copy_(src[a:b], non_blocking=True)
                if last:
                    h.put_signal(r, ch)
                    ev = torch.cuda.Event()
                    ev.record(st)
                    self._copied.append(ev)

    def _wait_symm(self, buf: int) -> None:
        h, _ = self._symm[buf]
        cur = torch.cuda.current_stream(self.device)
        while self._xchan[buf]:
            ch = self._xchan[buf].pop(0)
            for r in self._pstreams:
                h.wait_signal(r, ch)              # peer r's rows of this exchange are in my buffer
        for ev in self._copied:                   # ... and my rows have left (the source may be overwritten)
            cur.wait_event(ev)
        self._copied = []

    def _post_exchange(self, src: torch.Tensor, a: int, b: int, buf: int) -> None:
        """rows [a, b) of the local [per x d] matrix -> the same rows of this rank's slot in EVERY rank's X[buf]:
        one grouped send/recv per panel, asynchronous (waited for by _wait before X[buf] is read)."""
        if self.world == 1:
            return
        X, per = self.X[buf], self.per
        if self._symm is not None:
            self._post_symm(src, a, b, buf)
            return
        if self._xmode == "allgather":
            # experiment (YR_SHARD_EXCHANGE=allgather): ONE NCCL all-gather of the whole matrix, posted with the last panel
            # (the rank-major layout of X is exactly all_gather_into_tensor's output) — no panel pipelining
            if b >= self.per:
                self._pending[buf].append([dist.all_gather_into_tensor(X, src, group=self.group, async_op=True)])
            return
        X[self.lo + a: self.lo + b].copy_(src[a:b])
        if self._xmode == "none":              # experiment (YR_SHARD_EXCHANGE=none): compute only, results are garbage
            self._pending[buf].append([])
            return
        ops_ = []
        for r in self._peers:
            ops_.append(dist.P2POp(dist.isend, src[a:b], self._grank(r), self.xgroup))
            ops_.append(dist.P2POp(dist.irecv, X[r * per + a: r * per + b], self._grank(r), self.xgroup))
        self._pending[buf].append(dist.batch_isend_irecv(ops_))      # one entry per posted round, in posting order

    def _wait(self, buf: int) -> None:
        if self._symm is not None:
            self._wait_symm(buf)
            return
        for works in self._pending[buf]:
            for w in works:
                w.wait()
        self._pending[buf] = []

    def _wait_round(self, buf: int) -> None:
        """wait for the OLDEST posted exchange round of this buffer"""
        if self._pending[buf]:
            for w in self._pending[buf].pop(0):
                w.wait()

    def _exchange_all(self, src: torch.Tensor, buf: int) -> None:
        for a, b, _ in self.panels:
            self._post_exchange(src, a, b, buf)

    def propagate(self):
        k, L = self.k, self.n_layers
        self._exchange_all(self.E[0], 0)
        for l in range(L):
            buf = l & 1
            X = self.X[buf] if self.world > 1 else self.E[l]
            if l == 0 and self.use_col_panels:
                # the operand of layer 0 has only just been posted (it follows the optimizer step): consume it column panel
                # by column panel as the rounds arrive, then transform and send the next layer's panels
                for p in range(len(self.panels)):
                    self._wait_round(buf)
                    k.spmm(self.colA[p], X, self.LE[l], p > 0)
                for a, b, _ in self.panels:
                    k.dense_fwd(self.E[l][a:b], self.LE[l][a:b], self.W1[l], self.W2[l], self.E[l + 1][a:b])
                    if l + 1 < L:
                        self._post_exchange(self.E[l + 1], a, b, buf ^ 1)
                continue
            self._wait(buf)
            for a, b, Ap in self.panels:
                k.spmm(Ap, X, self.LE[l][a:b], False)
                k.dense_fwd(self.E[l][a:b], self.LE[l][a:b], self.W1[l], self.W2[l], self.E[l + 1][a:b])
                if l + 1 < L:
                    self._post_exchange(self.E[l + 1], a, b, buf ^ 1)
        return self.E

    def train_step(self, uid, pos, neg, loss_acc) -> None:
        """uid/pos/neg int64 on self.device, identical on every rank; loss_acc double[1] += sum of -logsigmoid terms
        (replicated: every rank accumulates the same value)."""
        k, d, B, W, L = self.k, self.d, int(uid.numel()), self.width, self.n_layers
        if B > self._cap:
            self._R = torch.empty(3 * B, W, device=self.device, dtype=F32)
            self._Gr = torch.empty(3 * B, W, device=self.device, dtype=F32)
            self._cap = B
        R, Gr = self._R[: 3 * B], self._Gr[: 3 * B]
        self.propagate()
        # ---- tail: R viewed as [B x 3 x W]; row (b, which) = b*3 + which -> positions interleaved the same way
        bad = ((uid < 0) | (uid >= self.nU) | (pos < 0) | (pos >= self.nI) | (neg < 0) | (neg >= self.nI)).any()
        self.err |= bad.to(I32)
        lay = self.layout
        ids = torch.stack((lay.user_pos(uid.clamp(0, self.nU - 1)), lay.item_pos(pos.clamp(0, self.nI - 1)),
                           lay.item_pos(neg.clamp(0, self.nI - 1))), dim=1).reshape(-1).contiguous()
        for l in range(L + 1):
            k.gather_rows(self.E[l], self.lo, self.hi, self.total, ids, R, l * d, self.err)
        if self.world > 1:
            dist.all_reduce(R, op=dist.ReduceOp.SUM, group=self.group)
        k.rows_grad(R, B, W, Gr, loss_acc)
        for g in self.G:
            g.zero_()
        # owners sum the gradient rows of their table rows in batch order (stable sort of the local row ids: index plumbing,
        # shared by all layers) — no floating-point atomics, so a step is bit-identical from run to run
        own = (ids >= self.lo) & (ids < self.hi)
        rows_sorted, src = torch.sort(torch.where(own, ids - self.lo, torch.full_like(ids, -1)).to(I32), stable=True)
        src = src.to(I32)
        for l in range(L + 1):
            k.scatter_rows_sorted(self.G[l], rows_sorted, src, Gr, l * d, self.flags, self.scratch)
        # ---- backward through the layers
        for l in reversed(range(L)):
            buf = l & 1
            for pi, (a, b, _) in enumerate(self.panels):
                k.dense_bwd(self.E[l][a:b], self.LE[l][a:b], self.E[l + 1][a:b], self.G[l + 1][a:b], self.W1[l], self.W2[l],
                            self.G[l][a:b], self.T[a:b], self.dWp[pi, 0], self.dWp[pi, 1])
                self._post_exchange(self.T, a, b, buf)
                if self.use_col_panels and pi >= 1:              # round pi-1 has had a panel's worth of compute to arrive
                    self._wait_round(buf)
                    k.spmm(self.colAT[pi - 1], self.X[buf], self.G[l], True)
            torch.sum(self.dWp[:, 0], dim=0, out=self.dW[l])
            torch.sum(self.dWp[:, 1], dim=0, out=self.dW[L + l])
            if self.use_col_panels:
                self._wait_round(buf)
                k.spmm(self.colAT[len(self.panels) - 1], self.X[buf], self.G[l], True)
            else:
                self._wait(buf)
                k.spmm(self.AT, self.X[buf] if self.world > 1 else self.T, self.G[l], True)
        if self.world > 1:
            dist.all_reduce(self.dW, op=dist.ReduceOp.SUM, group=self.group)
        # ---- optimizer: parameter order embedding, W1.*, W2.* (nn.Module.parameters()); all tensors share the step count
        opt = self.optimizer.opt_struct(self.optimizer.step_count + 1)
        k.opt_step(self.E[0], self.G[0], self.mE, self.vE, opt)
        for i in range(2 * L):
            Wt = self.W1[i] if i < L else self.W2[i - L]
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
        """Full E_0 [N x d] in the reference's node order on every rank (tests)."""
        if self.world == 1:
            return self.layout.to_node_order(self.E[0])
        full = torch.empty(self.total, self.d, device=self.device, dtype=F32)
        dist.all_gather_into_tensor(full, self.E[0].contiguous(), group=self.group)
        return self.layout.to_node_order(full)
