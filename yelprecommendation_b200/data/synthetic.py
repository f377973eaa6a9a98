"""Seeded synthetic Yelp2018-shape inputs (host side, numpy) — SURVEY.md §8(d).

The real Yelp dump is not available (no network), so every config of BASELINE.json runs on a generated
graph with the published shape: 31,668 users x 38,048 items, exactly 1,561,406 unique interactions,
degree-skewed (log-normal user activity >= 10, Zipf item popularity, every item >= 1 interaction so the
Laplacian has no zero degree), with planted user clusters so that evaluation metrics are non-trivial.

Also restated here, because they define the INPUT FORMATS of the hot path:
  * the per-user 60/20/20 split of data/datasets/mf_data_pipeline.py:18-52 (sklearn train_test_split
    semantics: RandomState(seed).permutation(n), test = first ceil(.2 n), then valid = first ceil(.25 n'));
  * the one-negative-per-positive rejection sampling of data/datasets/mf_dataset.py:18-32, done ONCE
    ("pre-sampled triples", BASELINE.json north_star) and cut into DataLoader-style batches.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

YELP2018 = dict(num_users=31_668, num_items=38_048, nnz=1_561_406)


@dataclass
class Interactions:
    num_users: int
    num_items: int
    user: np.ndarray      # int64 [nnz], sorted by (user, item)
    item: np.ndarray      # int64 [nnz]
    rating: np.ndarray    # float32 [nnz]
    user_cluster: np.ndarray  # int32 [num_users]  (planted structure, for "trained-ish" embeddings)
    cluster_item_logit: np.ndarray  # float32 [n_clusters, num_items]


def make_interactions(num_users=YELP2018["num_users"], num_items=YELP2018["num_items"], nnz=YELP2018["nnz"],
                      seed=2018, n_clusters=16, min_user_deg=10, star_ratings=False) -> Interactions:
    rng = np.random.default_rng(seed)
    min_user_deg = min(min_user_deg, max(1, num_items // 4))
    assert nnz >= num_users * min_user_deg and nnz >= num_items, "nnz too small for the degree floors"
    assert nnz <= num_users * num_items // 2
    # --- user degrees: log-normal, floor, exact total ---
    raw = rng.lognormal(mean=0.0, sigma=1.0, size=num_users)
    extra = nnz - num_users * min_user_deg
    deg = min_user_deg + np.floor(raw / raw.sum() * extra).astype(np.int64)
    deg = np.minimum(deg, max(min_user_deg, num_items // 3))
    # --- planted clusters: each cluster re-ranks a Zipf(0.8) popularity curve ---
    zipf = 1.0 / np.power(np.arange(1, num_items + 1, dtype=np.float64), 0.8)
    global_rank = rng.permutation(num_items)
    user_cluster = rng.integers(0, n_clusters, size=num_users).astype(np.int32)
    logits = np.empty((n_clusters, num_items), dtype=np.float32)
    cdfs = []
    for c in range(n_clusters):
        rank_c = np.where(rng.random(num_items) < 0.5, global_rank, rng.permutation(num_items))
        p = zipf[np.argsort(np.argsort(rank_c))]
        p = p / p.sum()
        logits[c] = np.log(p).astype(np.float32)
        cdfs.append(np.cumsum(p))
    # --- draw with oversampling, dedupe, trim to the per-user quota ---
    keys = []
    for c in range(n_clusters):
        users_c = np.nonzero(user_cluster == c)[0]
        if users_c.size == 0:
            continue
        quota = (deg[users_c] * 1.5 + 4).astype(np.int64)
        u_rep = np.repeat(users_c, quota)
        it = np.searchsorted(cdfs[c], rng.random(u_rep.size), side="right").clip(0, num_items - 1)
        keys.append(u_rep * num_items + it)
    keys = np.unique(np.concatenate(keys))
    keys = keys[rng.permutation(keys.size)]
    order = np.argsort(keys // num_items, kind="stable")
    keys = keys[order]
    u_sorted = keys // num_items
    start = np.searchsorted(u_sorted, np.arange(num_users), side="left")
    rank_in_user = np.arange(keys.size) - start[u_sorted]
    keys = keys[rank_in_user < deg[u_sorted]]
    # --- every item at least once ---
    present = np.zeros(num_items, dtype=bool)
    present[keys % num_items] = True
    missing = np.nonzero(~present)[0]
    if missing.size:
        keys = np.concatenate([keys, rng.integers(0, num_users, size=missing.size) * num_items + missing])
    keys = np.unique(keys)
    # --- exact nnz: top up with uniform pairs, or trim pairs that keep both floors ---
    while keys.size < nnz:
        need = nnz - keys.size
        cand = rng.integers(0, num_users, size=need * 2 + 16) * num_items + rng.integers(0, num_items, size=need * 2 + 16)
        cand = np.setdiff1d(np.unique(cand), keys)
        keys = np.union1d(keys, cand[rng.permutation(cand.size)[:need]])
    while keys.size > nnz:
        u, i = keys // num_items, keys % num_items
        du, di = np.bincount(u, minlength=num_users), np.bincount(i, minlength=num_items)
        ok = np.nonzero((du[u] > min_user_deg) & (di[i] > 1))[0]
        drop = ok[rng.permutation(ok.size)[: max(1, (keys.size - nnz) // 4)]]
        # never drop two pairs of the same user/item in one pass (keeps the floors valid)
        _, fu = np.unique(u[drop], return_index=True)
        drop = drop[fu]
        _, fi = np.unique(i[drop], return_index=True)
        drop = drop[fi][: keys.size - nnz]
        keys = np.delete(keys, drop)
    user, item = keys // num_items, keys % num_items
    if star_ratings:
        rating = rng.integers(1, 6, size=nnz).astype(np.float32)
    else:
        rating = np.ones(nnz, dtype=np.float32)
    return Interactions(num_users, num_items, user.astype(np.int64), item.astype(np.int64), rating,
                        user_cluster, logits)


# ----------------------------------------------------------------------------------------------------
# per-user 60/20/20 split (data/datasets/mf_data_pipeline.py:18-52)
# ----------------------------------------------------------------------------------------------------
@dataclass
class Split:
    """CSR over users of the item lists, in the order the reference's DataFrames hold them."""
    train_ptr: np.ndarray
    train_items: np.ndarray
    valid_ptr: np.ndarray
    valid_items: np.ndarray
    test_ptr: np.ndarray
    test_items: np.ndarray

    def lists(self, which: str, users=None) -> List[List[int]]:
        ptr, items = getattr(self, f"{which}_ptr"), getattr(self, f"{which}_items")
        users = range(len(ptr) - 1) if users is None else users
        return [items[ptr[u]:ptr[u + 1]].tolist() for u in users]


def split_sizes(n: int) -> Tuple[int, int, int]:
    n_test = math.ceil(0.2 * n)
    n_rest = n - n_test
    n_valid = math.ceil(0.25 * n_rest)
    return n_rest - n_valid, n_valid, n_test


def split_per_user(inter: Interactions, seed=42) -> Split:
    """Vectorised restatement: sklearn's ShuffleSplit uses RandomState(seed).permutation(n), so the
    permutation depends only on the user's degree n — users are processed in groups of equal degree."""
    U = inter.num_users
    deg = np.bincount(inter.user, minlength=U)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    sizes = np.array([split_sizes(int(n)) if n > 0 else (0, 0, 0) for n in range(int(deg.max()) + 1)])
    tr_n, va_n, te_n = sizes[deg, 0], sizes[deg, 1], sizes[deg, 2]
    tr_ptr, va_ptr, te_ptr = (np.concatenate([[0], np.cumsum(x)]) for x in (tr_n, va_n, te_n))
    tr = np.empty(tr_ptr[-1], np.int64)
    va = np.empty(va_ptr[-1], np.int64)
    te = np.empty(te_ptr[-1], np.int64)
    for n in np.unique(deg):
        n = int(n)
        if n == 0:
            continue
        users = np.nonzero(deg == n)[0]
        n_train, n_valid, n_test = split_sizes(n)
        perm1 = np.random.RandomState(seed).permutation(n)
        rest_idx, test_idx = perm1[n_test:], perm1[:n_test]
        perm2 = np.random.RandomState(seed).permutation(n - n_test)
        valid_idx, train_idx = rest_idx[perm2[:n_valid]], rest_idx[perm2[n_valid:]]
        block = inter.item[ptr[users][:, None] + np.arange(n)[None, :]]   # [n_users_with_deg, n]
        tr[(tr_ptr[users][:, None] + np.arange(n_train)[None, :]).ravel()] = block[:, train_idx].ravel()
        va[(va_ptr[users][:, None] + np.arange(n_valid)[None, :]).ravel()] = block[:, valid_idx].ravel()
        te[(te_ptr[users][:, None] + np.arange(n_test)[None, :]).ravel()] = block[:, test_idx].ravel()
    return Split(tr_ptr, tr, va_ptr, va, te_ptr, te)


def eval_lists(split: Split, mode: str, users: Optional[np.ndarray] = None):
    """(eval_uid, pos_items, mask_items) of valid_eval_data / test_eval_data (mf_data_pipeline.py:49-50):
    valid -> pos = valid items, mask = train items; test -> pos = test items, mask = train + valid items."""
    U = len(split.train_ptr) - 1
    users = np.arange(U) if users is None else np.asarray(users)
    if mode == "valid":
        pos = split.lists("valid", users)
        mask = split.lists("train", users)
    else:
        pos = split.lists("test", users)
        tr, va = split.lists("train", users), split.lists("valid", users)
        mask = [a + b for a, b in zip(tr, va)]
    keep = [k for k, p in enumerate(pos) if len(p) > 0]   # groupby drops users without eval positives
    return users[keep].astype(np.int64), [pos[k] for k in keep], [mask[k] for k in keep]


# ----------------------------------------------------------------------------------------------------
# pre-sampled BPR triples (data/datasets/mf_dataset.py:18-32, done once)
# ----------------------------------------------------------------------------------------------------
def sample_triples(split: Split, num_items: int, seed=42, which="train", reject="train"):
    rng = np.random.default_rng(seed)
    ptr, items = getattr(split, f"{which}_ptr"), getattr(split, f"{which}_items")
    U = len(ptr) - 1
    user = np.repeat(np.arange(U, dtype=np.int64), np.diff(ptr))
    pos = items.astype(np.int64)
    # rejection set = the user's `pos_items` column of the sampled frame (train: train positives;
    # valid frame: train+valid positives, mf_data_pipeline.py:47-48)
    rp, ri = split.train_ptr, split.train_items
    rej_keys = np.repeat(np.arange(U, dtype=np.int64), np.diff(rp)) * num_items + ri
    if reject == "train+valid":
        vk = np.repeat(np.arange(U, dtype=np.int64), np.diff(split.valid_ptr)) * num_items + split.valid_items
        rej_keys = np.concatenate([rej_keys, vk])
    rej_keys = np.unique(rej_keys)
    neg = rng.integers(0, num_items, size=user.size)
    bad = np.nonzero(np.isin(user * num_items + neg, rej_keys, assume_unique=False))[0]
    while bad.size:
        neg[bad] = rng.integers(0, num_items, size=bad.size)
        bad = bad[np.isin(user[bad] * num_items + neg[bad], rej_keys)]
    perm = rng.permutation(user.size)
    return user[perm], pos[perm], neg[perm].astype(np.int64)


def to_batches(user, pos, neg, batch_size=2048):
    """Pre-collated DataLoader-style batches (last short batch kept, train.py:76)."""
    import torch
    out = []
    for s in range(0, len(user), batch_size):
        sl = slice(s, s + batch_size)
        out.append({"user_id": torch.from_numpy(np.ascontiguousarray(user[sl])),
                    "pos_item": torch.from_numpy(np.ascontiguousarray(pos[sl])),
                    "neg_item": torch.from_numpy(np.ascontiguousarray(neg[sl]))})
    return out


def planted_embeddings(inter: Interactions, d=64, seed=7, noise=0.05):
    """'Trained-ish' MF tables: user = cluster code + noise, item = projected cluster log-affinities, so that
    top-10 lists recover planted preferences and the eval metrics are non-zero."""
    rng = np.random.default_rng(seed)
    C = inter.cluster_item_logit.shape[0]
    code = rng.standard_normal((C, d)).astype(np.float32) / np.float32(np.sqrt(d))
    U = code[inter.user_cluster] + noise * rng.standard_normal((inter.num_users, d)).astype(np.float32)
    logit = inter.cluster_item_logit - inter.cluster_item_logit.mean(axis=0, keepdims=True)
    V = (np.linalg.pinv(code.astype(np.float64)) @ logit.astype(np.float64)).T.astype(np.float32)
    V += noise * rng.standard_normal(V.shape).astype(np.float32)
    return np.ascontiguousarray(U, dtype=np.float32), np.ascontiguousarray(V, dtype=np.float32)
