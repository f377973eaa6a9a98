"""Drop-in for the reference's models/ngcf.py:7-72 on the sm_100a kernels.

Same constructor `(cfg, num_users, num_items)`, same parameters / state_dict keys (`embedding.weight`,
`W1.{l}.weight`, `W2.{l}.weight`), same quirk Q9 (`_init_weights` exists but is never called: embeddings stay
N(0,1), Linear weights keep their kaiming-uniform default). `laplacian_matrix` arrives as the torch sparse COO
tensor the reference's pipeline builds; it is converted to device CSR (L and L^T) once and cached by identity.

One layer = yr_ngcf_layer_fwd: CSR SpMM (LE = L E) + fused [LE+E | E*LE] x [W1^T; W2^T] + LeakyReLU — the
N x N identity of models/ngcf.py:61 is never built ((L+I)E = LE + E).
"""
import torch
import torch.nn as nn

from .. import ops
from ..data.graph import LaplacianCSR, laplacian_to_csr
from .base_model import BaseModel

_LEAKY_SLOPE = 0.01   # nn.functional.leaky_relu default, models/ngcf.py:72


class NGCF(BaseModel):
    def __init__(self, cfg, num_users, num_items):
        super().__init__()
        if cfg.embed_size not in (32, 64, 128):
            from .. import _cabi
            raise _cabi.YelprecError(f"NGCF embed_size {cfg.embed_size}: the propagation kernels take 32 (FP32 pipe), 64 or 128 (tensor cores)")
        self.cfg = cfg
        # where the d x d transforms run (yr_dense_mode, include/yelprec_b200.h): per model, nothing process-wide
        self.dense_mode = int(getattr(cfg, "ngcf_dense_mode", 2))
        if self.dense_mode not in (0, 1, 2):
            raise ValueError(f"ngcf_dense_mode {self.dense_mode} not in (0, 1, 2)")
        self.num_users = num_users
        self.num_items = num_items
        self.embedding = nn.Embedding(num_users + num_items, cfg.embed_size, dtype=torch.float32)
        self.W1 = nn.ModuleList([nn.Linear(cfg.embed_size, cfg.embed_size, bias=False) for _ in range(cfg.num_orders)])
        self.W2 = nn.ModuleList([nn.Linear(cfg.embed_size, cfg.embed_size, bias=False) for _ in range(cfg.num_orders)])
        self._csr_cache = (None, None)

    def _init_weights(self):   # present, never invoked — exactly like the reference (Q9)
        for child in self.children():
            if isinstance(child, nn.Embedding):
                nn.init.xavier_uniform_(child.weight)

    def csr(self, laplacian_matrix) -> LaplacianCSR:
        if isinstance(laplacian_matrix, LaplacianCSR):
            return laplacian_matrix
        key, val = self._csr_cache
        if key is not laplacian_matrix:
            val = laplacian_to_csr(laplacian_matrix, self.embedding.weight.device)
            self._csr_cache = (laplacian_matrix, val)
        return val

    def embedding_propagation(self, last_embed: torch.Tensor, w1, w2, laplacian_matrix):
        return ops.ngcf_layer(last_embed, w1.weight, w2.weight, self.csr(laplacian_matrix), _LEAKY_SLOPE,
                              self.dense_mode)

    def _layers(self, laplacian_matrix):
        outs = [self.embedding.weight]
        for w1, w2 in zip(self.W1, self.W2):
            outs.append(self.embedding_propagation(outs[-1], w1, w2, laplacian_matrix))
        return outs

    def bpr_forward(self, user_id, pos_item_ids, neg_item_ids, laplacian_matrix):
        dev = self.embedding.weight.device
        u, p, n = (ops._ids(t, dev) for t in (user_id, pos_item_ids, neg_item_ids))
        cat = torch.cat(self._layers(laplacian_matrix), dim=1)
        users, items = cat[: self.num_users], cat[self.num_users:]
        return ops.mf_score(users, items, u, p), ops.mf_score(users, items, u, n)

    def forward(self, user_id, item_id, laplacian_matrix):
        dev = self.embedding.weight.device
        cat = torch.cat(self._layers(laplacian_matrix), dim=1)
        users, items = cat[: self.num_users], cat[self.num_users:]
        return ops.mf_score(users, items, ops._ids(user_id, dev), ops._ids(item_id, dev))
