"""Host-side builders of the hot path's sparse inputs.

* `build_laplacian` — sparse restatement of data/datasets/ngcf_data_pipeline.py:19-44 (the reference
  materialises two dense N x N float32 arrays, 19.4 GB each at Yelp shape, and hard-codes `.to('cuda')`):
      A[u, U+i] = A[U+i, u] = mean rating;  deg = A.sum(axis=0);  L = (D^-1/2 A) D^-1/2  in fp32,
  returned as a coalesced torch sparse COO tensor exactly like `NGCFDataPipeline.laplacian_matrix`.
* `coo_to_csr` / `LaplacianCSR` — COO -> CSR (int32) of L and of L^T for the SpMM kernels.
* `build_eval_csr` — `eval_data` (index user_id, columns pos_items / mask_items, mf_data_pipeline.py:49-50)
  -> CSR arrays of yr_eval_topk_metrics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def build_laplacian(user: np.ndarray, item: np.ndarray, rating: np.ndarray, num_users: int,
                    num_items: int) -> torch.Tensor:
    user = np.asarray(user, dtype=np.int64)
    item = np.asarray(item, dtype=np.int64)
    rating = np.asarray(rating, dtype=np.float64)
    n = num_users + num_items
    # pivot_table(values='rating') aggregates duplicates with the mean (ngcf_data_pipeline.py:23)
    key = user * num_items + item
    uniq, inv = np.unique(key, return_inverse=True)
    mean = (np.bincount(inv, weights=rating) / np.bincount(inv)).astype(np.float32)
    u, i = uniq // num_items, uniq % num_items
    nz = mean != 0            # fillna(0)/to_sparse() drop explicit zeros
    u, i, mean = u[nz], i[nz], mean[nz]
    rows = np.concatenate([u, num_users + i])
    cols = np.concatenate([num_users + i, u])
    vals = np.concatenate([mean, mean])
    # deg = column sums, accumulated in float32 like ndarray.sum(axis=0) on a float32 matrix
    deg = np.zeros(n, dtype=np.float32)
    order = np.argsort(rows, kind="stable")          # row-by-row accumulation order
    np.add.at(deg, cols[order], vals[order])
    with np.errstate(divide="ignore"):
        dinv = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)
    lv = (dinv[rows] * vals).astype(np.float32) * dinv[cols]     # (D^-1/2 A) D^-1/2, fp32, this association
    L = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, cols])), torch.from_numpy(lv.astype(np.float32)),
                                size=(n, n))
    return L.coalesce()


def coo_to_csr(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, n: int):
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    assert rowptr[-1] < 2 ** 31
    return rowptr.astype(np.int32), cols.astype(np.int32), vals.astype(np.float32)


class CSRMatrix:
    """Device CSR (int32 / fp32) + the SpMM load-balancing plan of include/yelprec_b200.h (yr_csr).
    On CPU (tests of the host logic) the plan is still built, only the struct() call needs CUDA."""

    def __init__(self, rowptr, col, val, device):
        """rowptr/col/val: numpy arrays, or torch tensors (col/val may already live on `device`: only rowptr is
        read on the host, for the plan)."""
        import ctypes as C
        from .. import _cabi
        lib = _cabi.load()
        if isinstance(rowptr, torch.Tensor):
            rowptr = rowptr.detach().cpu().numpy()
        self.n_rows = int(rowptr.shape[0] - 1)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        nc, ns, npart = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _cabi.check(lib.yr_spmm_plan_size_h(rowptr.ctypes.data, self.n_rows, C.byref(nc), C.byref(ns), C.byref(npart)),
                    "yr_spmm_plan_size_h")
        self.n_chunks, self.n_split_rows, self.n_partials = nc.value, ns.value, npart.value
        cdesc = np.zeros((max(nc.value, 1), 4), np.int32)
        srow, sptr = np.empty(max(ns.value, 1), np.int32), np.zeros(ns.value + 1, np.int32)
        _cabi.check(lib.yr_spmm_plan_fill_h(rowptr.ctypes.data, self.n_rows, cdesc.ctypes.data, srow.ctypes.data,
                                            sptr.ctypes.data), "yr_spmm_plan_fill_h")
        nbig = C.c_int32(0)
        _cabi.check(lib.yr_spmm_plan_big_h(rowptr.ctypes.data, self.n_rows, C.byref(nbig), None), "yr_spmm_plan_big_h")
        big = np.zeros(max(nbig.value, 1), np.int32)
        if nbig.value:
            _cabi.check(lib.yr_spmm_plan_big_h(rowptr.ctypes.data, self.n_rows, C.byref(nbig), big.ctypes.data), "yr_spmm_plan_big_h")
        self.n_big_rows = nbig.value
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.device = torch.device(device)
        tt = lambda a, dt_np, dt_t: (a.to(device=device, dtype=dt_t).contiguous() if isinstance(a, torch.Tensor)
                                     else t(a.astype(dt_np)))
        self.rowptr, self.col, self.val = t(rowptr), tt(col, np.int32, torch.int32), tt(val, np.float32, torch.float32)
        self.chunk_desc = t(cdesc)
        self.split_row, self.split_ptr = t(srow), t(sptr)
        self.big_split_idx = t(big)
        self.split_count = torch.zeros(max(ns.value, 1), dtype=torch.int32, device=device)
        self._partials = {}

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def struct(self, d: int):
        """yr_csr for embedding width d (allocates the partials scratch for that width once)."""
        from .. import _cabi
        part = self._partials.get(d)
        if part is None:
            part = torch.empty(max(self.n_partials, 1) * d, device=self.device, dtype=torch.float32)
            self._partials[d] = part
        p = _cabi.dptr
        return _cabi.YrCsr(self.n_rows, self.nnz, p(self.rowptr), p(self.col), p(self.val), self.n_chunks,
                           p(self.chunk_desc), self.n_split_rows,
                           p(self.split_row), p(self.split_ptr), p(part), p(self.split_count),
                           self.n_big_rows, p(self.big_split_idx) if self.n_big_rows else None,
                           int(getattr(self, "reserve_sms", 0)))


@dataclass
class LaplacianCSR:
    n: int
    fwd: CSRMatrix          # L
    bwd: CSRMatrix          # L^T (the same object when L is bit-wise symmetric, e.g. binary ratings)
    symmetric: bool

    @property
    def nnz(self) -> int:
        return self.fwd.nnz


def laplacian_to_csr(L: torch.Tensor, device) -> LaplacianCSR:
    """COO (as the reference hands it over) -> device CSR of L and L^T. Duplicate entries are summed."""
    Lc = L.detach().cpu().coalesce()
    idx = Lc.indices().numpy()
    val = Lc.values().numpy().astype(np.float32)
    n = int(Lc.shape[0])
    rp, ci, va = coo_to_csr(idx[0], idx[1], val, n)
    rpt, cit, vat = coo_to_csr(idx[1], idx[0], val, n)
    sym = bool(np.array_equal(rp, rpt) and np.array_equal(ci, cit) and np.array_equal(va, vat))
    fwd = CSRMatrix(rp, ci, va, device)
    bwd = fwd if sym else CSRMatrix(rpt, cit, vat, device)
    return LaplacianCSR(n, fwd, bwd, sym)


def build_laplacian_csr_device(user: torch.Tensor, item: torch.Tensor, rating: torch.Tensor, num_users: int,
                               num_items: int) -> LaplacianCSR:
    """ngcf_data_pipeline.py:19-44 on the device (yr_laplacian_build): interactions (int64 ids, fp32 ratings, any
    order, duplicates averaged) -> LaplacianCSR ready for the SpMM kernels. Bit-identical to
    laplacian_to_csr(build_laplacian(...)). L is symmetric by construction, so L^T shares the arrays."""
    from .. import _cabi
    lib = _cabi.load()
    dev = user.device
    user, item = user.to(torch.int64).contiguous(), item.to(torch.int64).contiguous()
    rating = rating.to(torch.float32).contiguous()
    nnz, n = int(user.numel()), int(num_users + num_items)
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    col = torch.empty(2 * nnz, dtype=torch.int32, device=dev)
    val = torch.empty(2 * nnz, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = lib.yr_laplacian_ws_bytes(nnz, int(num_users), int(num_items))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    p = _cabi.dptr
    _cabi.check(lib.yr_laplacian_build(p(user), p(item), p(rating), nnz, int(num_users), int(num_items), p(rowptr), p(col),
                                       p(val), p(ws), nbytes, p(err), _cabi.stream_ptr(dev)), "yr_laplacian_build")
    rp = rowptr.cpu()
    if int(err.item()):
        raise IndexError("build_laplacian_csr_device: index out of range in self")
    total = int(rp[-1])
    m = CSRMatrix(rp, col[:total].clone(), val[:total].clone(), dev)
    return LaplacianCSR(n, m, m, True)


@dataclass
class EvalCSR:
    eval_uid: np.ndarray     # int64 [n_eval]
    mask_ptr: np.ndarray     # int32 [n_eval+1]
    mask_idx: np.ndarray     # int32, ascending and unique per row
    act_ptr: np.ndarray      # int32 [n_eval+1]
    act_idx: np.ndarray      # int32, ORIGINAL order (metric.py:73-75 depends on it, quirk Q7)
    act_nuniq: np.ndarray    # int32 [n_eval] = |set(pos_items)|

    @property
    def n_eval(self) -> int:
        return int(self.eval_uid.shape[0])


def build_eval_csr(eval_uid: Sequence[int], pos_items: Sequence[Sequence[int]],
                   mask_items: Sequence[Sequence[int]], num_items: Optional[int] = None) -> EvalCSR:
    n = len(eval_uid)
    mlen = np.zeros(n + 1, dtype=np.int64)
    alen = np.zeros(n + 1, dtype=np.int64)
    masks, acts, nun = [], [], np.zeros(n, dtype=np.int32)
    for e in range(n):
        m = np.unique(np.asarray(mask_items[e], dtype=np.int64)) if len(mask_items[e]) else np.zeros(0, np.int64)
        if num_items is not None and m.size:
            m = np.where(m < 0, m + num_items, m)        # pred[mask] accepts negative (wrap-around) indices
            if m.min() < 0 or m.max() >= num_items:
                raise IndexError(f"mask item out of range for {num_items} items")   # numpy would raise too
            m = np.unique(m)
        a = np.asarray(pos_items[e], dtype=np.int64).reshape(-1)
        masks.append(m)
        acts.append(a)
        mlen[e + 1], alen[e + 1] = m.size, a.size
        nun[e] = np.unique(a).size
    cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0, np.int64)).astype(np.int32)
    return EvalCSR(np.asarray(eval_uid, dtype=np.int64), np.cumsum(mlen).astype(np.int32), cat(masks),
                   np.cumsum(alen).astype(np.int32), cat(acts), nun)


def eval_csr_from_frame(eval_data, num_items: Optional[int] = None) -> EvalCSR:
    """`eval_data`: DataFrame, index = user id, list columns `pos_items`, `mask_items`."""
    return build_eval_csr(eval_data.index.to_numpy(), list(eval_data["pos_items"]), list(eval_data["mask_items"]),
                          num_items)


def inv_log2_table(k: int) -> np.ndarray:
    return np.array([1.0 / math.log2(i + 1) for i in range(1, k + 1)], dtype=np.float64)
