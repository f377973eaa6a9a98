"""Drop-in for the reference's metric.py:7-109: same four functions, same (quirky) definitions, computed by
the topk_metrics kernel (yr_topk_metrics) on cuda — `actual` keeps its ORIGINAL order because MAP depends on it."""
from typing import Sequence

import numpy as np
import torch

from . import ops
from .data.graph import build_eval_csr


def _run(actual: Sequence, predicted: Sequence, k: int):
    n = len(actual)
    if n == 0:
        raise ZeroDivisionError("division by zero")
    dev = torch.device("cuda", torch.cuda.current_device())
    pred = np.full((n, k), -1, dtype=np.int64)
    for r, p in enumerate(predicted):
        p = np.asarray(p, dtype=np.int64).reshape(-1)[:k]
        if p.size < k:
            raise IndexError("index out of range")          # metric.py:74 indexes user_predicted[i-1] up to k
        pred[r] = p
    csr = build_eval_csr(np.zeros(n, np.int64), [list(a) for a in actual], [[] for _ in range(n)])
    ecsr = ops.DeviceEvalCSR(csr, dev, k)
    _, sums = ops.topk_metrics(torch.from_numpy(pred).to(dev), ecsr)
    return [float(x) for x in sums.cpu()], n


def precision_at_k(actual, predicted, k: int = 20) -> float:
    s, n = _run(actual, predicted, k)
    return s[0] / n


def recall_at_k(actual, predicted, k: int = 20) -> float:
    s, _ = _run(actual, predicted, k)
    return s[1] / s[4]


def map_at_k(actual, predicted, k: int = 20) -> float:
    s, _ = _run(actual, predicted, k)
    return s[2] / s[5]


def ndcg_at_k(actual, predicted, k: int = 20) -> float:
    s, _ = _run(actual, predicted, k)
    return s[3] / s[4]
