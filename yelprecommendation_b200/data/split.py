"""Per-user train / valid / test split and the evaluation CSRs on the device.

`split_per_user_device` is MFDataPipeline.split (reference data/datasets/mf_data_pipeline.py:18-52, pairwise branch) through
yr_split_per_user: MT19937 + numpy's legacy shuffle reproduced on the GPU, bit-identical to data/synthetic.py::split_per_user
(which is pinned to the reference's own split). `eval_csr_device` builds `valid_eval_data` / `test_eval_data`
(:49-50: pos_items of the evaluated part, mask_items = train (+ valid) items) as the device CSR of the fused evaluation kernel
without a host round trip; the torch ops in it are index plumbing (sort / unique / cumsum on the device).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import _cabi, ops

I32, I64 = torch.int32, torch.int64


@dataclass
class DeviceSplit:
    """CSR over users of the item lists, in the order the reference's DataFrames hold them (device tensors)."""
    train_ptr: torch.Tensor
    train_items: torch.Tensor
    valid_ptr: torch.Tensor
    valid_items: torch.Tensor
    test_ptr: torch.Tensor
    test_items: torch.Tensor


def _scan(cnt: torch.Tensor) -> torch.Tensor:
    ptr = torch.zeros(cnt.numel() + 1, dtype=I64, device=cnt.device)
    ptr[1:] = torch.cumsum(cnt, 0, dtype=I64)
    return ptr.to(I32)


def split_per_user_device(user_ptr: torch.Tensor, items: torch.Tensor, seed: int = 42) -> DeviceSplit:
    """user_ptr int32 [U + 1], items int64 [nnz] (each user's interactions in DataFrame order), both on the device."""
    lib = _cabi.load()
    dev = user_ptr.device
    user_ptr, items = user_ptr.to(I32).contiguous(), items.to(I64).contiguous()
    U = int(user_ptr.numel() - 1)
    p = _cabi.dptr
    st = _cabi.stream_ptr(dev)
    n_tr, n_va, n_te = (torch.empty(max(U, 1), dtype=I32, device=dev) for _ in range(3))
    _cabi.check(lib.yr_split_sizes(p(user_ptr), U, p(n_tr), p(n_va), p(n_te), st), "yr_split_sizes")
    tr_ptr, va_ptr, te_ptr = _scan(n_tr[:U]), _scan(n_va[:U]), _scan(n_te[:U])
    max_len = int((user_ptr[1:] - user_ptr[:-1]).max().item()) if U else 1
    sizes = torch.stack([tr_ptr[-1], va_ptr[-1], te_ptr[-1]]).tolist()
    tr, va, te = (torch.empty(max(int(n), 1), dtype=I64, device=dev) for n in sizes)
    nbytes = lib.yr_split_ws_bytes(max(max_len, 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=I32, device=dev)
    _cabi.check(lib.yr_split_per_user(p(user_ptr), p(items), U, max(max_len, 1), int(seed) & 0xFFFFFFFF, p(tr_ptr), p(va_ptr),
                                      p(te_ptr), p(tr), p(va), p(te), p(ws), nbytes, p(err), st), "yr_split_per_user")
    if int(err.item()):
        raise RuntimeError("yr_split_per_user: permutation stream exhausted")
    return DeviceSplit(tr_ptr, tr[: int(sizes[0])], va_ptr, va[: int(sizes[1])], te_ptr, te[: int(sizes[2])])


def _rows_of(ptr: torch.Tensor) -> torch.Tensor:
    n = ptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n, device=ptr.device), (ptr[1:] - ptr[:-1]).to(I64))


def eval_csr_device(split: DeviceSplit, mode: str, num_items: int, K: int) -> ops.DeviceEvalCSR:
    """valid: pos = valid items, mask = train items; test: pos = test items, mask = train + valid items
    (mf_data_pipeline.py:49-50). Users without positives in the evaluated part are dropped (groupby does)."""
    dev = split.train_ptr.device
    if mode == "valid":
        pos_ptr, pos_items = split.valid_ptr, split.valid_items
        m_user, m_item = _rows_of(split.train_ptr), split.train_items
    else:
        pos_ptr, pos_items = split.test_ptr, split.test_items
        m_user = torch.cat([_rows_of(split.train_ptr), _rows_of(split.valid_ptr)])
        m_item = torch.cat([split.train_items, split.valid_items])
    U = int(pos_ptr.numel() - 1)
    pos_cnt = (pos_ptr[1:] - pos_ptr[:-1]).to(I64)
    keep = pos_cnt > 0
    eval_uid = keep.nonzero(as_tuple=False).view(-1)
    new_row = torch.cumsum(keep.to(I64), 0) - 1                       # user -> evaluation row
    # mask_items: ascending and unique per row
    mk = torch.unique(m_user[keep[m_user]] * num_items + m_item[keep[m_user]])
    m_rows = new_row[mk // num_items]
    mask_ptr = torch.zeros(eval_uid.numel() + 1, dtype=I64, device=dev)
    mask_ptr[1:] = torch.cumsum(torch.bincount(m_rows, minlength=eval_uid.numel()), 0)
    # pos_items: original order (metric.py:73-75 depends on it, quirk Q7); |set(pos_items)| per row
    act_ptr = torch.zeros(eval_uid.numel() + 1, dtype=I64, device=dev)
    act_ptr[1:] = torch.cumsum(pos_cnt[keep], 0)
    p_user = _rows_of(pos_ptr)
    pk = torch.unique(p_user * num_items + pos_items)
    nuniq = torch.bincount(new_row[pk // num_items], minlength=eval_uid.numel())
    return ops.DeviceEvalCSR.from_device(eval_uid, mask_ptr.to(I32), (mk % num_items).to(I32), act_ptr.to(I32),
                                         pos_items.to(I32), nuniq.to(I32), K)
