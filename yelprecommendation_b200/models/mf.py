"""Drop-in for the reference's models/mf.py:7-23 on the sm_100a kernels.

Same constructor `(cfg, num_users, num_items)`, same parameters / state_dict keys
(`user_embedding.weight`, `item_embedding.weight`, xavier-uniform), same `forward(user_id, item_id)`.
forward runs yr_mf_score (one fp32 fma chain per pair) and is differentiable (yr_mf_score_bwd), so the
reference's own trainer loop — loss.backward() + torch.optim — works unchanged on it; the fused trainer in
trainers/mf_trainer.py bypasses autograd entirely.
"""
import torch
import torch.nn as nn

from .. import ops
from .base_model import BaseModel


class MatrixFactorization(BaseModel):
    def __init__(self, cfg, num_users, num_items):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, cfg.embed_size, dtype=torch.float32)
        self.item_embedding = nn.Embedding(num_items, cfg.embed_size, dtype=torch.float32)
        self._init_weights()

    def _init_weights(self):
        for child in self.children():
            if isinstance(child, nn.Embedding):
                nn.init.xavier_uniform_(child.weight)

    def forward(self, user_id, item_id):
        return ops.mf_score(self.user_embedding.weight, self.item_embedding.weight, user_id, item_id)
