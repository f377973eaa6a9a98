"""BASELINE config 5 inputs, built on the device: the scaled synthetic interaction graph (10 M users x 2 M items, ~500 M
interactions) and each rank's row block of its symmetric-normalised Laplacian — no host COO, no dense array.

* `make_scaled_graph`   — user-side CSR (sorted, duplicate-free item lists) from the Philox generator kernel
                          (yr_synth_user_rows); log-normal user activity, Zipf item popularity, calibrated to `target_nnz`.
                          A function of (seed, sizes) only: every rank generates the same graph.
* `ShardLayout`         — who owns which node: rank k owns a contiguous block of users AND a contiguous block of items (a block
                          partition of the node ids alone would give the item-owning ranks ~5x the non-zeros of the others);
                          the all-gathered operand X is indexed by `position = rank * per + local row`.
* `shard_laplacian`     — rows [users of rank k ; items of rank k] of L = (D^-1/2 A) D^-1/2 (data/datasets/ngcf_data_pipeline.py:19-44,
                          binary ratings) as CSR with X positions as column ids. Users and items are mapped monotonically, so
                          every row keeps the column order — hence the fp32 summation order — it has on one GPU.
* `shard_laplacian_from_coo` — the same row blocks from the reference's torch sparse COO Laplacian (any weights), on the device.

torch ops used here are index plumbing (cumsum / sort / bincount on the device); the arithmetic of the Laplacian values is
yr_laplacian_binary_values.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from .. import _cabi
from .graph import CSRMatrix

I32, I64, F32 = torch.int32, torch.int64, torch.float32


@dataclass
class ScaledGraph:
    num_users: int
    num_items: int
    user_ptr: torch.Tensor      # int32 [num_users + 1]
    user_items: torch.Tensor    # int32 [nnz], ascending and unique inside a user
    seed: int

    @property
    def nnz(self) -> int:
        return int(self.user_items.numel())


def _synth(lib, seed, nU, nI, alpha, mu, sigma, min_deg, max_deg, rowptr, cnt, items, dev):
    p = _cabi.dptr
    _cabi.check(lib.yr_synth_user_rows(int(seed) & (2 ** 64 - 1), int(nU), int(nI), float(alpha), float(mu), float(sigma),
                                       int(min_deg), int(max_deg), p(rowptr), p(cnt), p(items), _cabi.stream_ptr(dev)),
                "yr_synth_user_rows")


def make_scaled_graph(num_users: int, num_items: int, target_nnz: int, seed: int = 5, device="cuda", zipf_alpha: float = 0.8,
                      sigma: float = 1.0, min_deg: int = 10, max_deg: int = 1024) -> ScaledGraph:
    lib = _cabi.load()
    dev = torch.device(device)
    cnt = torch.empty(num_users, dtype=I32, device=dev)
    mu = math.log(max(target_nnz / num_users, 1.0)) - 0.5 * sigma * sigma
    for it in range(4):                     # calibrate the activity scale to the requested number of interactions
        _synth(lib, seed, num_users, num_items, zipf_alpha, mu, sigma, min_deg, max_deg, None, cnt, None, dev)
        total = int(cnt.sum(dtype=I64).item())
        if abs(total - target_nnz) <= 0.005 * target_nnz or it == 3:
            break                           # `cnt` and `mu` belong together from here on (the fill pass repeats these draws)
        mu += math.log(target_nnz / max(total, 1))
    if total >= 2 ** 31:
        raise ValueError("more than 2^31 interactions")
    ptr = torch.zeros(num_users + 1, dtype=I64, device=dev)
    ptr[1:] = torch.cumsum(cnt, 0, dtype=I64)
    ptr = ptr.to(I32)
    items = torch.empty(total, dtype=I32, device=dev)
    _synth(lib, seed, num_users, num_items, zipf_alpha, mu, sigma, min_deg, max_deg, ptr, None, items, dev)
    return ScaledGraph(int(num_users), int(num_items), ptr, items, int(seed))


@dataclass
class ShardLayout:
    """Who owns which node and where its row sits. Rank k owns users [k * perU, (k+1) * perU) and items [k * perI, (k+1) * perI).
    Inside a rank the rows are cut into `panels` row panels of `pp = perUp + perIp` rows, each holding a slice of the rank's
    users FOLLOWED BY a slice of its items: every panel then carries about 1 / panels of the rank's non-zeros and of BOTH kinds
    of columns (a [all users ; all items] order would put every item — half of the non-zeros, and all the columns the user
    rows reference — into the last panel, and the panel pipeline would degenerate). panels = 1 is [users ; items]."""
    num_users: int
    num_items: int
    world: int
    panels: int = 1

    def __post_init__(self):
        P = max(1, int(self.panels))
        self.perU = (self.num_users + self.world - 1) // self.world      # users / items a rank owns (ids are blocked by these)
        self.perI = (self.num_items + self.world - 1) // self.world
        self.perUp = (self.perU + P - 1) // P                           # of them per panel
        self.perIp = (self.perI + P - 1) // P
        self.panels = P
        self.pp = self.perUp + self.perIp                                # rows of a panel ...
        if P > 1:
            self.pp += (-self.pp) % 128                                  # ... padded to whole 128-row tiles of the dense kernels
        self.per = P * self.pp                                           # rows of every rank's block (the rest is padding)

    def user_block(self, rank: int) -> Tuple[int, int]:
        a = min(rank * self.perU, self.num_users)
        return a, min(a + self.perU, self.num_users)

    def item_block(self, rank: int) -> Tuple[int, int]:
        a = min(rank * self.perI, self.num_items)
        return a, min(a + self.perI, self.num_items)

    def panel_rows(self, p: int) -> Tuple[int, int]:
        return p * self.pp, (p + 1) * self.pp

    def user_local(self, lu: torch.Tensor) -> torch.Tensor:
        """index inside the rank's user block -> local row"""
        return (lu // self.perUp) * self.pp + (lu % self.perUp)

    def item_local(self, li: torch.Tensor) -> torch.Tensor:
        return (li // self.perIp) * self.pp + self.perUp + (li % self.perIp)

    def user_pos(self, u: torch.Tensor) -> torch.Tensor:
        """position of user ids in the gathered operand [world * per rows] (monotonic in u)"""
        if self.world == 1 and self.panels == 1:
            return u
        return (u // self.perU) * self.per + self.user_local(u % self.perU)

    def item_pos(self, i: torch.Tensor) -> torch.Tensor:
        """position of item ids (monotonic in i)"""
        if self.world == 1 and self.panels == 1:
            return i + self.num_users
        return (i // self.perI) * self.per + self.item_local(i % self.perI)

    def node_pos(self, n: torch.Tensor) -> torch.Tensor:
        """reference node ids ([users ; items], models/ngcf.py:33-35) -> positions"""
        is_u = n < self.num_users
        return torch.where(is_u, self.user_pos(torch.where(is_u, n, torch.zeros_like(n))),
                           self.item_pos(torch.where(is_u, torch.zeros_like(n), n - self.num_users)))

    def local_rows(self, rank: int, table: torch.Tensor) -> torch.Tensor:
        """rows of a [num_users + num_items, d] table (reference node order) owned by `rank`, padded to `per` rows"""
        (u0, u1), (i0, i1) = self.user_block(rank), self.item_block(rank)
        out = torch.zeros(self.per, table.shape[1], dtype=table.dtype, device=table.device)
        dev = table.device
        out[self.user_local(torch.arange(u1 - u0, device=dev))] = table[u0:u1]
        out[self.item_local(torch.arange(i1 - i0, device=dev))] = table[self.num_users + i0: self.num_users + i1]
        return out

    def to_node_order(self, gathered: torch.Tensor) -> torch.Tensor:
        """[world * per, d] gathered blocks -> [num_users + num_items, d] in the reference's node order"""
        dev = gathered.device
        return torch.cat([gathered[self.user_pos(torch.arange(self.num_users, device=dev))],
                          gathered[self.item_pos(torch.arange(self.num_items, device=dev))]])


def _values(lib, rowptr, row_node, col, deg_pos, dev) -> torch.Tensor:
    val = torch.empty(max(int(col.numel()), 1), dtype=F32, device=dev)
    p = _cabi.dptr
    _cabi.check(lib.yr_laplacian_binary_values(p(rowptr), int(rowptr.numel() - 1), p(row_node), p(col), p(deg_pos), p(val),
                                               _cabi.stream_ptr(dev)), "yr_laplacian_binary_values")
    return val[: col.numel()]


def shard_laplacian(g: ScaledGraph, layout: ShardLayout, rank: int):
    """-> (rowptr int32 [per + 1], col int32 (X positions), val fp32) of this rank's row block, all on the device."""
    lib = _cabi.load()
    dev = g.user_ptr.device
    nU, nI = g.num_users, g.num_items
    (u0, u1), (i0, i1) = layout.user_block(rank), layout.item_block(rank)
    ptr64 = g.user_ptr.to(I64)
    # ---- degrees of every node, indexed by position
    deg_user = (ptr64[1:] - ptr64[:-1]).to(I32)
    deg_item = torch.bincount(g.user_items, minlength=nI).to(I32)
    deg_pos = torch.zeros(layout.world * layout.per, dtype=I32, device=dev)
    deg_pos[layout.user_pos(torch.arange(nU, device=dev))] = deg_user
    deg_pos[layout.item_pos(torch.arange(nI, device=dev))] = deg_item
    # ---- user rows of the block: a slice of the user-side CSR, item ids -> positions
    e0, e1 = int(ptr64[u0].item()), int(ptr64[u1].item())
    ucol = layout.item_pos(g.user_items[e0:e1].to(I64)).to(I32)
    ucnt = deg_user[u0:u1].to(I64)
    # ---- item rows of the block: the entries whose item falls in [i0, i1), grouped by item; users ascending inside an
    # item because the entries are enumerated user-major and the sort is stable
    sel = ((g.user_items >= i0) & (g.user_items < i1)).nonzero(as_tuple=False).view(-1)
    it_local = (g.user_items[sel] - i0).to(I32)
    owner_user = (torch.searchsorted(ptr64, sel, right=True) - 1)
    del sel
    it_sorted, perm = torch.sort(it_local, stable=True)
    icol = layout.user_pos(owner_user[perm]).to(I32)
    del owner_user, perm, it_local
    icnt = torch.bincount(it_sorted, minlength=i1 - i0).to(I64)
    del it_sorted
    # ---- rowptr in LOCAL row order (panel by panel: a slice of the users, then a slice of the items, then padding)
    lrow_u = layout.user_local(torch.arange(u1 - u0, device=dev))
    lrow_i = layout.item_local(torch.arange(i1 - i0, device=dev))
    cnt = torch.zeros(layout.per, dtype=I64, device=dev)
    cnt[lrow_u] = ucnt
    cnt[lrow_i] = icnt
    rowptr = torch.zeros(layout.per + 1, dtype=I64, device=dev)
    rowptr[1:] = torch.cumsum(cnt, 0)
    if int(rowptr[-1].item()) >= 2 ** 31:
        raise ValueError("row block with more than 2^31 non-zeros")
    if layout.panels == 1:
        col = torch.cat([ucol, icol])                    # [users ; items] is already the local order
    else:
        # every entry moves to (start of its row in the local order) + (its rank inside the row)
        col = torch.empty(int(rowptr[-1].item()), dtype=I32, device=dev)
        for cols_, cnts_, lrow_ in ((ucol, ucnt, lrow_u), (icol, icnt, lrow_i)):
            start_src = torch.cumsum(cnts_, 0) - cnts_                                   # first entry of every source row
            shift = rowptr[lrow_] - start_src                                            # dest - src, constant inside a row
            col[torch.arange(cols_.numel(), device=dev) + torch.repeat_interleave(shift, cnts_)] = cols_
        del ucol, icol
    rowptr = rowptr.to(I32)
    row_node = torch.arange(rank * layout.per, (rank + 1) * layout.per, dtype=I32, device=dev)
    val = _values(lib, rowptr, row_node, col, deg_pos, dev)
    return rowptr, col, val


def shard_laplacian_from_coo(L: torch.Tensor, layout: ShardLayout, rank: int, device, transpose: bool = False):
    """Row block of L (or of L^T) from the reference's sparse COO [N x N] Laplacian, on the device: node ids -> positions,
    rows of this rank kept, sorted by (row, column)."""
    dev = torch.device(device)
    Lc = L.detach().to(dev).coalesce()
    idx, val = Lc.indices(), Lc.values().to(F32)
    r, c = (idx[1], idx[0]) if transpose else (idx[0], idx[1])
    rp, cp = layout.node_pos(r), layout.node_pos(c)
    lo, hi = rank * layout.per, (rank + 1) * layout.per
    keep = (rp >= lo) & (rp < hi)
    rp, cp, val = rp[keep] - lo, cp[keep], val[keep]
    order = torch.argsort(rp * (layout.world * layout.per) + cp)
    rp, cp, val = rp[order], cp[order], val[order]
    rowptr = torch.zeros(layout.per + 1, dtype=I64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(rp, minlength=layout.per), 0)
    return rowptr.to(I32), cp.to(I32), val.contiguous()


def csr_row_panels(rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n_panels: int, device):
    """The row block as `n_panels` CSRMatrix views over consecutive row ranges (shared col / val storage; each panel has
    its own SpMM plan): the unit of the compute / all-gather pipeline. Returns [(row0, row1, CSRMatrix)]."""
    n = int(rowptr.numel() - 1)
    n_panels = max(1, min(int(n_panels), n))
    step = (n + n_panels - 1) // n_panels
    step = (step + 127) // 128 * 128                      # whole 128-row tiles of the dense kernels
    rp_h = rowptr.cpu()
    out = []
    for a in range(0, n, step):
        b = min(a + step, n)
        out.append((a, b, CSRMatrix(rp_h[a: b + 1], col, val, device)))
    return out
