"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front-end of oracle/yr_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Everything takes and returns numpy arrays on the host.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


class OrcOpt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("step", C.c_int32), ("lr", C.c_double), ("weight_decay", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double)]


_KINDS = {"sgd": 0, "adam": 1, "adamw": 2}


def build() -> str:
    src = os.path.join(_HERE, "yr_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_bpr_loss.restype = C.c_float
        _lib.orc_mf_train_step.restype = C.c_float
        _lib.orc_ngcf_tail.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def make_opt(kind, lr, weight_decay=0.0, step=1, betas=(0.9, 0.999), eps=1e-8) -> OrcOpt:
    return OrcOpt(_KINDS[kind.lower()], int(step), float(lr), float(weight_decay), float(betas[0]),
                  float(betas[1]), float(eps))


def inv_log2_table(k: int) -> np.ndarray:
    """1/log2(i+1) for i = 1..k, with Python's math.log2 (metric.py:107)."""
    return np.array([1.0 / math.log2(i + 1) for i in range(1, k + 1)], dtype=np.float64)


def mf_score(U, V, uid, iid):
    U, V, uid, iid = _f32(U), _f32(V), _i64(uid), _i64(iid)
    out = np.empty(len(uid), dtype=np.float32)
    lib().orc_mf_score(_p(U), _p(V), C.c_int(U.shape[1]), _p(uid), _p(iid), C.c_int64(len(uid)), _p(out))
    return out


def bpr_loss(pos, neg) -> float:
    pos, neg = _f32(pos), _f32(neg)
    return float(lib().orc_bpr_loss(_p(pos), _p(neg), C.c_int64(len(pos))))


def dense_opt_step(p, g, m, v, opt: OrcOpt):
    """In place on float32 arrays p, m, v."""
    assert p.dtype == np.float32 and p.flags.c_contiguous
    g = _f32(g)
    lib().orc_dense_opt_step(_p(p), _p(g), _p(m), _p(v), C.c_int64(p.size), C.byref(opt))


class MFTrainerOracle:
    """MFTrainer.train over pre-collated batches (trainers/mf_trainer.py:100-116)."""

    def __init__(self, U, V, optimizer="sgd", lr=1e-4, weight_decay=0.0):
        self.U, self.V = _f32(U).copy(), _f32(V).copy()
        self.kind, self.lr, self.wd = optimizer, lr, weight_decay
        self.mU, self.vU = np.zeros_like(self.U), np.zeros_like(self.U)
        self.mV, self.vV = np.zeros_like(self.V), np.zeros_like(self.V)
        self.gU, self.gV = np.zeros_like(self.U), np.zeros_like(self.V)
        self.step = 0

    def train(self, batches):
        total, per_step = 0.0, []
        for b in batches:
            uid, pos, neg = _i64(b["user_id"]), _i64(b["pos_item"]), _i64(b["neg_item"])
            self.step += 1
            opt = make_opt(self.kind, self.lr, self.wd, self.step)
            l = lib().orc_mf_train_step(_p(self.U), _p(self.V), C.c_int64(self.U.shape[0]),
                                        C.c_int64(self.V.shape[0]), C.c_int(self.U.shape[1]), _p(self.mU),
                                        _p(self.vU), _p(self.mV), _p(self.vV), _p(self.gU), _p(self.gV),
                                        _p(uid), _p(pos), _p(neg), C.c_int64(len(uid)), C.byref(opt))
            per_step.append(float(l))
            total += float(l)
        return total, per_step


def spmm_csr(rowptr, col, val, X, Y=None):
    rowptr, col, val, X = _i32(rowptr), _i32(col), _f32(val), _f32(X)
    acc = Y is not None
    out = _f32(Y).copy() if acc else np.empty_like(X)
    lib().orc_spmm_csr(_p(rowptr), _p(col), _p(val), C.c_int64(len(rowptr) - 1), C.c_int(X.shape[1]), _p(X),
                       _p(out), C.c_int(1 if acc else 0))
    return out


def ngcf_layer_fwd(csr, E, W1, W2, slope=0.01):
    rowptr, col, val = (_i32(csr[0]), _i32(csr[1]), _f32(csr[2]))
    E, W1, W2 = _f32(E), _f32(W1), _f32(W2)
    En, LE = np.empty_like(E), np.empty_like(E)
    lib().orc_ngcf_layer_fwd(_p(rowptr), _p(col), _p(val), C.c_int64(E.shape[0]), C.c_int(E.shape[1]), _p(E),
                             _p(W1), _p(W2), C.c_float(slope), _p(En), _p(LE))
    return En, LE


def ngcf_layer_bwd(csrT, E, LE, En, Gn, W1, W2, G, slope=0.01):
    rowptr, col, val = (_i32(csrT[0]), _i32(csrT[1]), _f32(csrT[2]))
    E, LE, En, Gn, W1, W2 = map(_f32, (E, LE, En, Gn, W1, W2))
    G = _f32(G).copy()
    T = np.empty_like(E)
    d = E.shape[1]
    dW1, dW2 = np.empty((d, d), np.float32), np.empty((d, d), np.float32)
    lib().orc_ngcf_layer_bwd(_p(rowptr), _p(col), _p(val), C.c_int64(E.shape[0]), C.c_int(d), _p(E), _p(LE),
                             _p(En), _p(Gn), _p(W1), _p(W2), C.c_float(slope), _p(G), _p(T), _p(dW1), _p(dW2))
    return G, dW1, dW2


def ngcf_tail(E_layers, nU, uid, pos, neg, want_grad=True):
    E_layers = [_f32(e) for e in E_layers]
    uid, pos, neg = _i64(uid), _i64(pos), _i64(neg)
    L = len(E_layers) - 1
    d = E_layers[0].shape[1]
    G_layers = [np.zeros_like(e) for e in E_layers] if want_grad else None
    EP = (C.c_void_p * (L + 1))(*[e.ctypes.data for e in E_layers])
    GP = (C.c_void_p * (L + 1))(*[g.ctypes.data for g in G_layers]) if want_grad else None
    po, no = np.empty(len(uid), np.float32), np.empty(len(uid), np.float32)
    loss = lib().orc_ngcf_tail(EP, GP, C.c_int(L), C.c_int64(nU), C.c_int(d), _p(uid), _p(pos), _p(neg),
                               C.c_int64(len(uid)), _p(po), _p(no))
    return float(loss), po, no, G_layers


def eval_topk_metrics(Uemb, Vemb, eval_uid, mask_ptr, mask_idx, act_ptr, act_idx, K=10):
    Uemb, Vemb = _f32(Uemb), _f32(Vemb)
    eval_uid = _i64(eval_uid)
    mask_ptr, mask_idx, act_ptr, act_idx = map(_i32, (mask_ptr, mask_idx, act_ptr, act_idx))
    n = len(eval_uid)
    topk = np.empty((n, K), np.int64)
    tsc = np.empty((n, K), np.float32)
    um = np.zeros((n, 4), np.float64)
    sums = np.zeros(6, np.float64)
    il = inv_log2_table(K)
    lib().orc_eval_topk_metrics(_p(Uemb), _p(Vemb), C.c_int64(Vemb.shape[0]), C.c_int(Uemb.shape[1]),
                                _p(eval_uid), C.c_int64(n), _p(mask_ptr), _p(mask_idx), _p(act_ptr), _p(act_idx),
                                _p(il), C.c_int(K), _p(topk), _p(tsc), _p(um), _p(sums))
    return topk, tsc, um, sums


def metrics_from_sums(sums, n_eval):
    """(precision, recall, map, ndcg) exactly as metric.py divides them (Q6)."""
    p = sums[0] / n_eval if n_eval else float("nan")
    r = sums[1] / sums[4] if sums[4] else float("nan")
    m = sums[2] / sums[5] if sums[5] else float("nan")
    n = sums[3] / sums[4] if sums[4] else float("nan")
    return p, r, m, n


def philox4x32_10(ctr, key) -> np.ndarray:
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(np.ascontiguousarray(ctr, dtype=np.uint32)), _p(np.ascontiguousarray(key, dtype=np.uint32)), _p(out))
    return out


def sample_negatives(uid, pos_ptr, pos_idx, num_items, seed, offset=0, max_blocks=64):
    """(neg int64[n], n_failed) — Philox4x32-10 rejection sampler, see oracle/yr_oracle.c."""
    uid = _i64(uid)
    neg = np.empty(uid.shape[0], np.int64)
    f = lib().orc_sample_negatives
    f.restype = C.c_int64
    failed = f(_p(uid), C.c_int64(uid.shape[0]), _p(_i32(pos_ptr)), _p(_i32(pos_idx)), C.c_int64(num_items),
               C.c_uint64(seed), C.c_uint64(offset), C.c_int(max_blocks), _p(neg))
    return neg, int(failed)
