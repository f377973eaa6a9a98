"""Device-resident training-triple loader — SURVEY.md §8(f)2.

Stands where the reference puts `DataLoader(MFDataset(train_data, num_items), batch_size, shuffle=True)`
(train.py:70-77, data/datasets/mf_dataset.py): every epoch each training interaction (user, pos_item) gets ONE fresh
negative, uniform over the items outside the user's `pos_items`, and the triples are visited in a new random order.
The reference does this with a per-sample pandas `.iloc` + Python rejection loop (its real end-to-end bottleneck);
here the interactions live in HBM, negatives come from `yr_sample_negatives` (counter-based Philox, a function of
(seed, epoch, interaction index) only) and the epoch never leaves the device.

Iterating yields the reference's batch dicts {'user_id','pos_item','neg_item'} (int64, on the device, last short batch
kept); trainers that see `.epoch_triples()` skip the per-batch path and run the whole epoch in one launch.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class DeviceTripleLoader:
    def __init__(self, user: np.ndarray, pos: np.ndarray, rej_ptr: np.ndarray, rej_items: np.ndarray, num_users: int,
                 num_items: int, batch_size: int = 2048, device="cuda", seed: int = 42, shuffle: bool = True):
        """user/pos: the training interactions (one row of the reference's train_data each). rej_ptr/rej_items: CSR of
        every user's `pos_items` column (train positives for the train frame, train+valid for the valid frame —
        mf_data_pipeline.py:41-48); sorted per user here."""
        self.device = torch.device(device)
        self.num_users, self.num_items, self.batch_size = int(num_users), int(num_items), int(batch_size)
        self.seed, self.shuffle, self.epoch = int(seed), bool(shuffle), 0
        rej_ptr = np.asarray(rej_ptr, dtype=np.int64)
        rej_items = np.asarray(rej_items, dtype=np.int64)
        rows = np.repeat(np.arange(len(rej_ptr) - 1), np.diff(rej_ptr))
        order = np.lexsort((rej_items, rows))
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(self.device)
        self.rej_ptr, self.rej_idx = t(rej_ptr, np.int32), t(rej_items[order], np.int32)
        self.user, self.pos = t(np.asarray(user), np.int64), t(np.asarray(pos), np.int64)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)

    @classmethod
    def from_split(cls, split, num_items: int, which="train", **kw):
        """`split`: data.synthetic.Split. train frame rejects train positives; valid frame train+valid (reference)."""
        ptr, items = getattr(split, f"{which}_ptr"), getattr(split, f"{which}_items")
        U = len(ptr) - 1
        user = np.repeat(np.arange(U, dtype=np.int64), np.diff(ptr))
        if which == "train":
            rp, ri = split.train_ptr, split.train_items
        else:
            cnt = np.diff(split.train_ptr) + np.diff(split.valid_ptr)
            rp = np.concatenate([[0], np.cumsum(cnt)])
            # per user: train items then valid items
            tu = np.repeat(np.arange(U), np.diff(split.train_ptr))
            vu = np.repeat(np.arange(U), np.diff(split.valid_ptr))
            users = np.concatenate([tu, vu])
            items_all = np.concatenate([split.train_items, split.valid_items])
            o = np.argsort(users, kind="stable")
            ri = items_all[o]
        return cls(user, items.astype(np.int64), rp, ri, U, num_items, **kw)

    def __len__(self) -> int:
        return (self.user.numel() + self.batch_size - 1) // self.batch_size

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def epoch_triples(self):
        """(user, pos, neg) int64 device tensors of this epoch, already in visiting order; advances the epoch."""
        n = self.user.numel()
        # stream offset: epoch e owns counters [e * 2^40, e * 2^40 + n)
        neg = ops.sample_negatives(self.user, self.rej_ptr, self.rej_idx, self.num_users, self.num_items, self.seed,
                                   offset=self.epoch << 40, err=self.err)
        if self.shuffle:
            g = torch.Generator(device=self.device).manual_seed(self.seed * 1_000_003 + self.epoch)
            perm = torch.randperm(n, device=self.device, generator=g)
            out = (self.user[perm], self.pos[perm], neg[perm])
        else:
            out = (self.user, self.pos, neg)
        self.epoch += 1
        code = int(self.err.item())
        if code:
            self.err.zero_()
            raise (IndexError("user id out of range") if code == 1 else
                   RuntimeError("negative sampling found no item outside a user's positives"))
        return out

    def __iter__(self):
        u, p, n = self.epoch_triples()
        B = self.batch_size
        for s in range(0, u.numel(), B):
            yield {"user_id": u[s:s + B], "pos_item": p[s:s + B], "neg_item": n[s:s + B]}
