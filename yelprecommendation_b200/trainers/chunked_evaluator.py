"""Model-agnostic chunked full-catalog evaluator on the fused top-K + metrics kernels — SURVEY.md 8(f)4.

The reference's DCN trainer (trainers/dcn_trainer.py:145-203) evaluates a model that is NOT a dot product: for every evaluation
row it scores the catalog in chunks of cfg.batch_size items through `self.model(user, items, categories, statecity)`, masks the
train items with 0 (its outputs are sigmoids, :191), takes argpartition / argsort top-K on the host and feeds metric.py.
Here the same loop keeps everything on the device: `score_fn(user_ids [R], item_ids [C]) -> [R x C]` is called chunk by chunk
for a block of R evaluation rows at a time, the [R x num_items] score block never leaves HBM, yr_topk_masked_rows applies the
mask value and picks the K best by (score desc, item id asc), and yr_topk_metrics accumulates the reference's four metrics
(metric.py:7-109, quirks Q6-Q8). The DCN network itself is out of scope; any model plugs in through `score_fn`.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from .. import _cabi, ops
from ..data.graph import EvalCSR, eval_csr_from_frame

I32, I64, F32 = torch.int32, torch.int64, torch.float32


class ChunkedTopKEvaluator:
    def __init__(self, num_items: int, top_n: int, device, chunk_size: int, mask_value: float = 0.0, rows_per_block: int = 64,
                 valid_rows: int = 1000):
        """chunk_size = cfg.batch_size (items per model call, dcn_trainer.py:149); mask_value 0 as in :191;
        valid_rows: mode == 'valid' evaluates the first 1,000 rows only (:152-153)."""
        self.num_items, self.top_n = int(num_items), int(top_n)
        self.device = torch.device(device)
        self.chunk, self.mask_value, self.rows, self.valid_rows = int(chunk_size), float(mask_value), int(rows_per_block), int(valid_rows)
        self.last_topk = None

    def evaluate(self, score_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], eval_data, mode: str = "valid") -> Tuple[float, float, float, float]:
        """eval_data: the reference's DataFrame (index user_id, list columns pos_items / mask_items) or a data.graph.EvalCSR."""
        lib = _cabi.load()
        dev, nI, K = self.device, self.num_items, self.top_n
        if mode == "valid" and not isinstance(eval_data, EvalCSR):
            eval_data = eval_data[: self.valid_rows]
        csr = eval_data if isinstance(eval_data, EvalCSR) else eval_csr_from_frame(eval_data, nI)
        ecsr = ops.DeviceEvalCSR(csr, dev, K)
        n = ecsr.n_eval
        if n == 0:
            raise ZeroDivisionError("division by zero")
        items = torch.arange(nI, device=dev, dtype=I64)
        topk = torch.empty(n, K, device=dev, dtype=I64)
        pred = torch.empty(self.rows, nI, device=dev, dtype=F32)
        ws = torch.empty(self.rows * nI, device=dev, dtype=torch.uint8)
        p = _cabi.dptr
        for r0 in range(0, n, self.rows):
            r1 = min(r0 + self.rows, n)
            uid = ecsr.eval_uid[r0:r1]
            for c0 in range(0, nI, self.chunk):
                c1 = min(c0 + self.chunk, nI)
                out = score_fn(uid, items[c0:c1])
                if not out.is_cuda:
                    raise _cabi.YelprecError("score_fn must return a CUDA tensor (no CPU fallback)")
                pred[: r1 - r0, c0:c1] = out.reshape(r1 - r0, c1 - c0).to(F32)
            mp = (ecsr.mask_ptr[r0:r1 + 1]).contiguous()        # absolute offsets into mask_idx
            _cabi.check(lib.yr_topk_masked_rows(p(pred), nI, r1 - r0, nI, p(mp), p(ecsr.mask_idx), self.mask_value, K,
                                                topk.data_ptr() + r0 * K * 8, p(ws), ws.numel(), _cabi.stream_ptr(dev)),
                        "yr_topk_masked_rows")
        self.last_topk = topk
        _, sums = ops.topk_metrics(topk, ecsr)
        return ops.metrics_from_sums(sums.cpu(), n)
