"""Oracle for one organization's local training / prediction (torch CPU fp32).

TEST INFRASTRUCTURE (see oracle/__init__.py). Follows reference src/organization.py:140-217 and the
optimizer contract of src/utils.py:248-259 (Adam with L2 weight decay, global-norm clip to 1).
"""
import math

import numpy as np
import torch

from . import models


def make_batch(data, target, rows, data_mode):
    """What the reference's dataset ``__getitem__`` + FlatInput/PairInput + collate produce for the row ids
    ``rows`` (reference src/datasets/movielens.py:249-286, src/data.py:140-151, src/utils.py:302-306):
    COO triples of each row in CSR storage order, the aligned id repeated per entry, rows concatenated."""
    other = "item" if data_mode == "user" else "user"
    rows = np.asarray(rows, dtype=np.int64)
    out = {}
    for pre, m in (("", data), ("target_", target)):
        starts, ends = m.indptr[rows], m.indptr[rows + 1]
        cnt = ends - starts
        pos = np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)]) if len(rows) else np.zeros(0, np.int64)
        pos = pos.astype(np.int64)
        out[pre + data_mode] = torch.from_numpy(np.repeat(rows, cnt))
        out[pre + other] = torch.from_numpy(m.indices[pos].astype(np.int64))
        out[pre + "rating"] = torch.from_numpy(m.data[pos].astype(np.float32))
    return out


def clip_coef(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_(params, 1): coef = min(1, max_norm / (||g||_2 + 1e-6)) over ALL grads
    (reference call site src/organization.py:161)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    return min(1.0, max_norm / (total + 1e-6)), total


class Adam:
    """torch.optim.Adam(lr, betas, weight_decay) semantics (reference src/utils.py:253-254): L2 decay is added
    to the (already clipped) gradient, bias-corrected moments, eps outside the sqrt. Dense: every parameter
    that received a gradient is updated every step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4):
        self.p = params
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    def step(self, grads, scale=1.0):
        for k, g in grads.items():
            if g is None:
                continue
            w = self.p[k]
            self.t[k] += 1
            t = self.t[k]
            g = g * scale + self.wd * w
            self.m[k] = self.m[k] + (g - self.m[k]) * (1 - self.b1)
            self.v[k] = self.v[k] * self.b2 + (1 - self.b2) * g * g
            step_size = self.lr / (1 - self.b1 ** t)
            denom = self.v[k].sqrt() / math.sqrt(1 - self.b2 ** t) + self.eps
            self.p[k] = w - step_size * self.m[k] / denom


def grads_of(loss, params):
    names = list(params)
    gs = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    return dict(zip(names, gs))


def leaf(params):
    return {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}


def train_step(opt, forward):
    """One optimizer step: loss.backward(); clip_grad_norm_(.,1); Adam.step() (src/organization.py:158-162)."""
    p = leaf(opt.p)
    pred, loss = forward(p)
    g = grads_of(loss, p)
    coef, _ = clip_coef([x for x in g.values() if x is not None])
    opt.p = {k: v.detach() for k, v in p.items()}
    opt.step(g, coef)
    return float(loss)


def train_org_ae(params, data, target, data_mode, target_mode, epoch_batches, masks, hp=None):
    """Organization.train (reference src/organization.py:140-178): for every epoch, for every index batch: skip
    when the batch has no DATA entries (:153-155); AE forward with local=True (MSE on the residual targets),
    backward, clip, Adam. ``epoch_batches[e]`` = list of row-id arrays (the sampler order), ``masks`` = iterator of
    dropout keep-masks in consumption order."""
    hp = hp or {}
    opt = Adam({k: v.clone() for k, v in params.items()}, **hp)
    masks = iter(masks)
    losses = []
    for batches in epoch_batches:
        for rows in batches:
            b = make_batch(data, target, rows, data_mode)
            if len(b[data_mode]) == 0:
                continue
            keep = next(masks)
            losses.append(train_step(opt, lambda p: models.ae_forward(p, b, data_mode, target_mode, True, keep,
                                                                       local=True)))
    return opt.p, losses


def predict_org_ae(params, data, target, data_mode, target_mode, batch_size):
    """Organization.predict (reference src/organization.py:180-217): sequential batches, eval forward at every
    target position, skip batches without targets; returns values in the target CSR's storage order."""
    n = min(target.shape[0], data.shape[0])  # a cold-start organization only walks the rows it holds (len = data rows)
    out = np.zeros(target.nnz, dtype=np.float32)
    with torch.no_grad():
        for s in range(0, n, batch_size):
            rows = np.arange(s, min(n, s + batch_size))
            b = make_batch(data, target, rows, data_mode)
            if len(b["target_" + data_mode]) == 0:
                continue
            pred, _ = models.ae_forward(params, b, data_mode, target_mode, False)
            out[target.indptr[rows[0]]:target.indptr[rows[-1] + 1]] = pred.numpy()
    return out


def base_round0(data, target_train, target_test, data_mode, target_mode, batch_size):
    """Organization.initialize (reference src/organization.py:29-138): one pass of models.base over the train
    loader (pair batches, no shuffle), then predictions at every train / test target position."""
    n = data.shape[0]
    model = models.Base(data.shape[1], target_mode)
    preds = {"train": np.zeros(target_train.nnz, np.float32), "test": np.zeros(target_test.nnz, np.float32)}
    for s in range(0, n, batch_size):
        rows = np.arange(s, min(n, s + batch_size))
        b = make_batch(data, target_train, rows, data_mode)
        other = "item" if data_mode == "user" else "user"
        model.fit(b[other], b["rating"], b[data_mode])
    for split, tgt in (("train", target_train), ("test", target_test)):
        preds[split] = model.predict(torch.from_numpy(tgt.indices.astype(np.int64))).numpy()
    return preds, model
