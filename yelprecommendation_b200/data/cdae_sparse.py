"""Index-list form of the reference's CDAE batches (data/datasets/cdae_dataset.py:38-62 yields dense [num_items] float32 masks
per user: 152 KB each at Yelp shape). `sparse_cdae_batch` converts a collated dense batch once, on the host, into the few
hundred bytes per user that `CDAETrainer.train` ships to the device through yr_cdae_step_idx; a dataset that already holds
item lists can build the same dict directly.

    {'user_id': int64 [B],
     'input_ptr': int32 [B + 1], 'input_idx': int32 [nnz]       active inputs (input_mask != 0), ascending per row
     'loss_ptr':  int32 [B + 1], 'loss_idx':  int32 [m], 'loss_val': float32 [m]
                                                                  positions where target + negative_mask != 0 (loss.py:12-16),
                                                                  ascending per row, value = the target (1 / 0)}
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch


def _csr(mask: np.ndarray):
    r, c = np.nonzero(mask)
    ptr = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.add.at(ptr, r + 1, 1)
    return np.cumsum(ptr).astype(np.int32), c.astype(np.int32), r


def sparse_cdae_batch(data: Dict[str, torch.Tensor], target_extra: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """`data`: a dense batch ('user_id', 'input_mask', 'negative_mask', optionally 'valid_mask'); `target_extra='valid_mask'`
    makes the loss target input_mask + valid_mask as CDAETrainer.validate does (trainers/cdae_trainer.py:67)."""
    x = data["input_mask"].cpu().numpy()
    tgt = x if target_extra is None else x + data[target_extra].cpu().numpy()
    neg = data["negative_mask"].cpu().numpy()
    in_ptr, in_idx, _ = _csr(x != 0)
    sel = (tgt + neg) != 0
    ls_ptr, ls_idx, ls_row = _csr(sel)
    ls_val = tgt[ls_row, ls_idx].astype(np.float32)
    t = torch.from_numpy
    return {"user_id": data["user_id"].to(torch.int64), "input_ptr": t(in_ptr), "input_idx": t(in_idx),
            "loss_ptr": t(ls_ptr), "loss_idx": t(ls_idx), "loss_val": t(ls_val)}
