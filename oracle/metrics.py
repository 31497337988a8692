"""Oracle evaluation metrics with the reference's batching (TEST INFRASTRUCTURE, see oracle/__init__.py).

reference src/train_recsys_assist.py:175-217 evaluates the global output in row blocks of ``batch_size`` and
src/logger.py:35-55 keeps the n-weighted mean of the per-block values; src/metrics/metrics.py:8-11 (RMSE) and
:63-84 (NDCG@10 on a densified block, unobserved = -inf for the ranking and 0 for the relevance)."""
import numpy as np
import torch

from .models import loss_fn


def ndcg_block(pred, rel, row, col, topk=10):
    ur, ri = np.unique(row, return_inverse=True)
    uc, ci = np.unique(col, return_inverse=True)
    score = torch.full((len(ur), len(uc)), -float("inf"))
    gain = torch.zeros(len(ur), len(uc))
    score[ri, ci] = torch.as_tensor(pred)
    gain[ri, ci] = torch.as_tensor(rel)
    k = min(topk, score.shape[1])
    disc = 1.0 / torch.log2(torch.arange(1, k + 1, dtype=torch.float32) + 1)
    top = score.topk(k, dim=-1).indices
    dcg = (gain.gather(1, top) * disc).sum(-1)
    idcg = (gain.topk(k, dim=-1).values * disc).sum(-1)
    q = dcg / idcg
    q = torch.nan_to_num(q, nan=0.0, posinf=0.0, neginf=0.0)
    return float(q.mean())


def evaluate(F, y_csr, data_mode, target_mode, batch_size):
    n = y_csr.shape[0]
    acc = {}
    cnt = 0
    for s in range(0, n, batch_size):
        lo, hi = y_csr.indptr[s], y_csr.indptr[min(n, s + batch_size)]
        m = hi - lo
        if m == 0:
            continue
        p = torch.as_tensor(F[lo:hi], dtype=torch.float32)
        t = torch.as_tensor(y_csr.data[lo:hi], dtype=torch.float32)
        vals = {"Loss": float(loss_fn(p, t, target_mode))}
        if target_mode == "explicit":
            vals["RMSE"] = float(((p - t) ** 2).mean().sqrt())
        else:
            rows = np.repeat(np.arange(s, min(n, s + batch_size)), np.diff(y_csr.indptr[s:min(n, s + batch_size) + 1]))
            vals["NDCG"] = ndcg_block(p, t, rows, y_csr.indices[lo:hi])
        for k, v in vals.items():
            acc[k] = (acc.get(k, 0.0) * cnt + v * m) / (cnt + m)
        cnt += m
    return {"test/" + k: v for k, v in acc.items()}
