"""Synthetic rating data in the shapes the reference's datasets produce.

There is no network in the build/bench environment, so ML1M / Douban / Amazon are
replaced by seeded synthetic matrices with the same dimensions and the same
on-disk layout (``(data_csr, target_csr)`` per split, ``item_attr``,
``user_profile``; reference ``src/datasets/movielens.py:231-247,325-337,362-371``).
The shapes are the ones SURVEY.md §8d fixes (Douban / Amazon dimensions are
ASSUMED there because the reference never records them).
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass

import numpy as np
from scipy.sparse import csr_matrix

# name -> (M users, N items, nnz, #genre columns, #profile columns or 0)
SHAPES = {
    "ML100K": (943, 1682, 100_000, 18, 30),
    "ML1M": (6040, 3706, 1_000_209, 18, 30),
    "Douban": (2800, 36000, 1_300_000, 3, 8),
    "Amazon": (5000, 40000, 600_000, 4, 0),
    # tiny shapes used by the parity tests / golden fixtures (named like the real ones so
    # that the reference's per-dataset tables in process_control() apply)
    "tiny-ML100K": (120, 90, 3_000, 18, 30),
    "tiny-Douban": (150, 110, 4_000, 3, 8),
    "tiny-Amazon": (130, 100, 2_500, 4, 0),
}

_RATING_P = np.array([0.056, 0.108, 0.261, 0.349, 0.226])  # ML1M-like marginals of 1..5


@dataclass
class RatingData:
    """One dataset in the reference's in-memory form (user-major CSR, fp32)."""

    name: str
    train: csr_matrix  # M x N, explicit ratings in {1..5}
    test: csr_matrix  # M x N
    item_attr: np.ndarray  # N x G fp32 multi-hot / one-hot
    user_profile: np.ndarray | None  # M x P fp32 one-hot blocks, or None (Amazon)

    @property
    def shape(self):
        return self.train.shape

    def split(self, target_mode: str):
        """(train_data, train_target), (test_data, test_target) like make_explicit/implicit_data
        (reference src/datasets/movielens.py:352-392): test 'data' is the train matrix."""
        if target_mode == "explicit":
            tr, te = self.train, self.test
        elif target_mode == "implicit":
            tr, te = _binarize(self.train), _binarize(self.test)
        else:
            raise ValueError("Not valid target mode")
        return (tr, tr), (tr, te)


def _binarize(m: csr_matrix) -> csr_matrix:
    out = m.copy()
    out.data = (m.data >= 3.5).astype(np.float32)  # explicit zeros are kept, as in the reference
    return out


def _degrees(rng, count, total, minimum, cap):
    """Heavy-tailed positive integer degrees, each in [minimum, cap], summing to ``total``."""
    if count * minimum > total:
        raise ValueError("nnz too small for the per-row minimum")
    w = rng.lognormal(mean=0.0, sigma=1.0, size=count)
    d = minimum + np.floor(w / w.sum() * (total - count * minimum)).astype(np.int64)
    d = np.minimum(d, cap)
    # distribute the remainder one by one over rows that still have room
    rest = int(total - d.sum())
    while rest > 0:
        room = np.flatnonzero(d < cap)
        take = rng.choice(room, size=min(rest, len(room)), replace=False)
        d[take] += 1
        rest = int(total - d.sum())
    return d


def make_rating_data(name: str = "ML1M", seed: int = 0, shape=None, min_per_user: int | None = None) -> RatingData:
    """Seeded synthetic dataset. Distinct (user,item) pairs, every user >= min_per_user ratings,
    Zipf-like item popularity, ML1M-like rating marginals, one 90/10 permutation split."""
    M, N, nnz, G, P = SHAPES[name] if shape is None else shape
    rng = np.random.default_rng(seed)
    if min_per_user is None:
        min_per_user = 20 if nnz >= 20 * M else max(1, nnz // (2 * M))
    deg = _degrees(rng, M, nnz, min_per_user, N)
    pop = 1.0 / np.arange(1, N + 1) ** 0.8
    pop = pop[rng.permutation(N)]
    logw = np.log(pop)
    rows = np.repeat(np.arange(M, dtype=np.int64), deg)
    cols = np.empty(nnz, dtype=np.int64)
    off = 0
    chunk = max(1, (1 << 24) // N)
    for u0 in range(0, M, chunk):
        u1 = min(M, u0 + chunk)
        key = logw[None, :] + rng.gumbel(size=(u1 - u0, N))  # Gumbel top-k = weighted sampling w/o replacement
        order = np.argsort(-key, axis=1)
        for u in range(u0, u1):
            d = deg[u]
            cols[off:off + d] = np.sort(order[u - u0, :d])
            off += d
    rating = rng.choice(np.arange(1, 6), size=nnz, p=_RATING_P).astype(np.float32)
    idx = rng.permutation(nnz)
    n_train = int(nnz * 0.9)
    tr, te = idx[:n_train], idx[n_train:]
    train = csr_matrix((rating[tr], (rows[tr], cols[tr])), shape=(M, N))
    test = csr_matrix((rating[te], (rows[te], cols[te])), shape=(M, N))
    # genres: 1..3 hot for ML*, exactly one-hot otherwise
    item_attr = np.zeros((N, G), dtype=np.float32)
    first = rng.integers(0, G, size=N)
    item_attr[np.arange(N), first] = 1
    if name.endswith("ML1M") or name.endswith("ML100K"):
        extra = rng.random((N, G)) < (0.6 / G)
        item_attr[extra] = 1
    user_profile = None
    if P:
        blocks = _profile_blocks(P)
        user_profile = np.zeros((M, P), dtype=np.float32)
        o = 0
        for b in blocks:
            user_profile[np.arange(M), o + rng.integers(0, b, size=M)] = 1
            o += b
    return RatingData(name, train, test, item_attr, user_profile)


def make_scaled_data(M, N, nnz, n_genre=8, seed=0, name="scaled"):
    """Large synthetic shapes (config 5 family: up to 1M x 500K): Zipf-like item draws by inverse CDF, duplicates
    dropped, so generation is O(nnz log N) instead of the O(M*N) Gumbel top-k used for the small shapes."""
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, N + 1) ** 0.8
    cdf = np.cumsum(pop[rng.permutation(N)])
    cdf /= cdf[-1]
    per = max(1, int(nnz * 1.08 / M))
    rows = np.repeat(np.arange(M, dtype=np.int64), per)
    cols = np.searchsorted(cdf, rng.random(rows.size)).astype(np.int64)
    key = np.unique(rows * N + cols)
    if key.size > nnz:
        key = np.sort(rng.choice(key, size=nnz, replace=False))
    rows, cols = key // N, key % N
    rating = rng.choice(np.arange(1, 6), size=key.size, p=_RATING_P).astype(np.float32)
    idx = rng.permutation(key.size)
    n_train = int(key.size * 0.9)
    tr, te = idx[:n_train], idx[n_train:]
    train = csr_matrix((rating[tr], (rows[tr], cols[tr])), shape=(M, N))
    test = csr_matrix((rating[te], (rows[te], cols[te])), shape=(M, N))
    item_attr = np.zeros((N, n_genre), dtype=np.float32)
    item_attr[np.arange(N), rng.integers(0, n_genre, size=N)] = 1
    return RatingData(name, train, test, item_attr, None)


def make_scaled_data_device(M, N, nnz, n_genre=8, seed=0, name="scaled", device="cuda"):
    """Same family as make_scaled_data, generated with torch on `device` (the host version spends ~2 s per million
    ratings in numpy sorts; at 1e8 ratings that is minutes of a GPU call): Zipf-like item draws by inverse CDF, the
    (row, column) keys deduplicated by one device sort, a Bernoulli(0.9) train / test split (the reference's split is an
    exact 90 % permutation, src/datasets/movielens.py:362-364; the scaled runs only need the shape), CSR arrays built
    from the sorted keys and handed to scipy without another sort."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pop = 1.0 / torch.arange(1, N + 1, device=device, dtype=torch.float64) ** 0.8
    cdf = torch.cumsum(pop[torch.randperm(N, device=device, generator=g)], 0)
    cdf = (cdf / cdf[-1]).float()
    per = max(1, int(nnz * 1.08 / M))
    rows = torch.arange(M, device=device, dtype=torch.int64).repeat_interleave(per)
    cols = torch.searchsorted(cdf, torch.rand(rows.numel(), device=device, generator=g)).clamp_(max=N - 1)
    key = torch.unique(rows * N + cols)  # sorted: row-major, ascending columns
    del rows, cols
    if key.numel() > nnz:
        keep = torch.rand(key.numel(), device=device, generator=g) < nnz / key.numel()
        key = key[keep]
    p = torch.tensor(_RATING_P, device=device, dtype=torch.float32).cumsum(0)
    rating = (torch.bucketize(torch.rand(key.numel(), device=device, generator=g), p).clamp_(max=4) + 1).float()
    is_train = torch.rand(key.numel(), device=device, generator=g) < 0.9
    out = []
    for m in (is_train, ~is_train):
        k = key[m]
        r = torch.div(k, N, rounding_mode="floor")
        indptr = torch.zeros(M + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(torch.bincount(r, minlength=M), 0)
        out.append(csr_matrix((rating[m].cpu().numpy(), (k - r * N).to(torch.int32).cpu().numpy(),
                               indptr.cpu().numpy()), shape=(M, N)))
        del k, r
    rng = np.random.default_rng(seed)
    item_attr = np.zeros((N, n_genre), dtype=np.float32)
    item_attr[np.arange(N), rng.integers(0, n_genre, size=N)] = 1
    return RatingData(name, out[0], out[1], item_attr, None)


def _profile_blocks(P):
    if P == 30:
        return [7, 2, 21]  # age, gender, occupation (reference src/datasets/movielens.py:409-415)
    return [P]


def reference_data_name(name: str) -> str:
    return name.split("-")[-1] if name.startswith("tiny-") else name


def write_reference_layout(data: RatingData, root: str) -> str:
    """Write ``<root>/<NAME>/processed/...`` pickles exactly as the reference's datasets load them
    (reference src/datasets/movielens.py:231-247; Amazon has no user_profile, src/datasets/amazon.py:82-83)."""
    base = os.path.join(root, reference_data_name(data.name), "processed")
    for mode in ("explicit", "implicit"):
        os.makedirs(os.path.join(base, mode), exist_ok=True)
        train_set, test_set = data.split(mode)
        with open(os.path.join(base, mode, "train.pt"), "wb") as f:
            pickle.dump(train_set, f)
        with open(os.path.join(base, mode, "test.pt"), "wb") as f:
            pickle.dump(test_set, f)
    with open(os.path.join(base, "item_attr.pt"), "wb") as f:
        pickle.dump(data.item_attr, f)
    if data.user_profile is not None:
        with open(os.path.join(base, "user_profile.pt"), "wb") as f:
            pickle.dump(data.user_profile, f)
    return base
