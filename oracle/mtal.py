"""Oracle for the MTAL coordinator: pseudo-residuals, privacy noise, update() (torch CPU fp32 / numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py). Follows reference src/assist.py:43-179 and src/privacy.py:6-24.
All vectors are aligned with the storage order of one canonical global CSR (rows x all columns); an
organization's view is the sub-sequence of entries whose column it owns, in that same order — which is what
scipy's ``M[:, cols_i].data`` yields in the reference (SURVEY.md §7 "Ordering contracts").
"""
import numpy as np
import torch

from . import models


def owner_views(indices, data_split, n_cols):
    """For every organization: positions (into the global CSR data array) of the entries in its columns, and the
    local column index of each (= position of the column inside data_split[i]); reference src/assist.py:94,107-108."""
    owner = np.full(n_cols, -1, dtype=np.int64)
    local = np.zeros(n_cols, dtype=np.int64)
    for i, cols in enumerate(data_split):
        cols = np.asarray(cols, dtype=np.int64)
        owner[cols] = i
        local[cols] = np.arange(len(cols))
    views = []
    for i in range(len(data_split)):
        pos = np.flatnonzero(owner[indices] == i)
        views.append((pos, local[indices[pos]]))
    return views


def residual(F, y, target_mode, clamp):
    """reference src/assist.py:45-58: r = -d/dF sum loss(F, y): explicit 2(y-F), implicit y-sigmoid(F);
    optionally clamped to [-1, 1] (Douban/Amazon except Douban-item-explicit, :51-56)."""
    F = torch.as_tensor(F, dtype=torch.float32)
    y = torch.as_tensor(y, dtype=torch.float32)
    if target_mode == "explicit":
        g = 2 * (F - y)
    elif target_mode == "implicit":
        g = torch.sigmoid(F) - y
    else:
        raise ValueError("Not valid target mode")
    if clamp:
        g = torch.clamp(g, min=-1, max=1)
    return (-g).numpy()


def needs_clamp(data_name, data_mode, target_mode):
    return data_name in ("Douban", "Amazon") and not (data_name == "Douban" and data_mode == "item"
                                                       and target_mode == "explicit")


def dp(y, alpha, rng=np.random):
    """reference src/privacy.py:6-24: clip to the [2.5 %, 97.5 %] quantiles, add Laplace((b-a)/alpha) noise drawn
    from the GLOBAL numpy generator (re-seeded to cfg['seed'] by every make_data_loader, src/data.py:76)."""
    a, b = np.quantile(y, 0.025), np.quantile(y, 0.975)
    scale = max(0, (b - a) / alpha)
    out = np.clip(y, a, b)
    return out + rng.laplace(scale=scale, size=y.shape)


def fit_assist(history, output, idx, target, n_rate, K, ar, ar_mode, aw_mode, target_mode, lr=0.1, steps=10):
    """L-BFGS fit of the assisted learning rate (one per owned column) and/or the assistance weights
    (reference src/assist.py:118-129; torch.optim.LBFGS(lr=0.1) defaults: max_iter 20, history 100, no line search;
    src/utils.py:255-256, :199-203)."""
    rate = torch.full((n_rate,), float(ar))
    weight = torch.ones(K) / K
    free = []
    if ar_mode == "optim":
        rate.requires_grad_(True)
        free.append(rate)
    if aw_mode == "optim":
        weight.requires_grad_(True)
        free.append(weight)
    if free:
        opt = torch.optim.LBFGS(free, lr=lr)
        for _ in range(steps):
            def closure():
                _, loss = models.assist_forward(rate, weight, history, output, idx, target, target_mode)
                opt.zero_grad()
                loss.backward()
                return loss

            opt.step(closure)
    return rate.detach(), weight.detach()


def update(F_prev, y, org_out, indices, data_split, n_cols, target_mode, ar, ar_mode="constant", aw_mode="constant",
           match_rate=1.0):
    """Assist.update (reference src/assist.py:81-179) without the cold-start ('cs') branch.
    F_prev, y: {split: values aligned with the global CSR}; org_out: list over organizations of {split: values};
    indices: {split: CSR column indices}. Returns F_next {split: values}, and per-owner (rate, weight)."""
    K = len(data_split)
    F_next = {k: np.array(F_prev[k], dtype=np.float32, copy=True) for k in F_prev}
    fitted = [None] * K
    views = {k: owner_views(indices[k], data_split, n_cols) for k in F_prev}

    def inputs(i, split):
        pos, idx = views[split][i]
        h = torch.from_numpy(np.asarray(F_prev[split])[pos].astype(np.float32))
        own = np.asarray(org_out[i][split])[pos]
        cols = []
        for j in range(K):
            o = own.copy()
            if match_rate < 1:
                # partial alignment: only the first int(n * match_rate) entries see the other organizations
                n_match = int(len(o) * match_rate)
                o[:n_match] = np.asarray(org_out[j][split])[pos][:n_match]
            else:
                o = np.asarray(org_out[j][split])[pos]
            cols.append(torch.from_numpy(o.astype(np.float32)))
        return pos, h, torch.stack(cols, -1), torch.from_numpy(idx)

    for i in range(K):
        n_rate = len(data_split[i])
        pos, h, O, idx = inputs(i, "train")
        t = torch.from_numpy(np.asarray(y["train"])[pos].astype(np.float32))
        fitted[i] = fit_assist(h, O, idx, t, n_rate, K, ar, ar_mode, aw_mode, target_mode)
        for split in F_prev:
            pos, h, O, idx = inputs(i, split)
            with torch.no_grad():
                pred, _ = models.assist_forward(fitted[i][0], fitted[i][1], h, O, idx)
            F_next[split][pos] = pred.numpy()
    return F_next, fitted
