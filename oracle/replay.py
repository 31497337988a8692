"""Oracle replay of a whole (shortened) MTAL experiment, RNG-identical to the reference on torch CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py). Mirrors the event order of reference
src/train_recsys_assist.py:40-93 so that every draw from torch's global CPU generator (data split, DataLoader
base seeds, sampler seeds, parameter init, dropout masks) and numpy's global generator (DP noise) happens in
the same order with the same shapes. The arithmetic itself is the oracle's (models / train / mtal).
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from . import metrics as ometrics
from . import models, mtal, train


def matrices(data, data_mode, target_mode):
    """{split: (data_csr, target_csr)} in the orientation the reference datasets use
    (rows = aligned entity; reference src/datasets/movielens.py:233-243,367-371)."""
    (trd, trt), (ted, tet) = data.split(target_mode)
    out = {"train": (trd, trt), "test": (ted, tet)}
    if data_mode == "item":
        out = {k: (a.T.tocsr(), b.T.tocsr()) for k, (a, b) in out.items()}
        for a, b in out.values():
            a.sort_indices()
            b.sort_indices()
    return out


def split_columns(mats, item_attr, mode, K, data_mode):
    """reference src/data.py:200-242. genre: one torch.multinomial draw per column, redrawn until every
    organization is non-empty in all four matrices; random-K: torch.randperm chunks, remainder to the last."""
    n_cols = mats["train"][0].shape[1]
    if "genre" in mode:
        if data_mode != "user":
            raise NotImplementedError
        attr = torch.tensor(item_attr)
        attr[attr.sum(-1) == 0] = 1
        while True:
            idx = torch.multinomial(attr, 1).view(-1).numpy()
            split = [np.where(idx == i)[0] for i in range(K)]
            ok = all(len(s) > 0 and all(m[:, s].nnz > 0 for pair in mats.values() for m in pair) for s in split)
            if ok:
                return [torch.tensor(s) for s in split]
    if "random" in mode:
        chunks = list(torch.randperm(n_cols).split(n_cols // K))
        return chunks[:K - 1] + [torch.cat(chunks[K - 1:])]
    raise ValueError("Not valid data split mode")


def loader_batches(n, batch_size, shuffle):
    """Row-id batches of one DataLoader pass. A real DataLoader over ``range(n)`` is iterated so the global
    generator is consumed exactly as in the reference (one base-seed draw per iterator, plus the RandomSampler's
    seed draw when shuffling; torch/utils/data/dataloader.py, sampler.py)."""
    return [np.asarray(b) for b in DataLoader(range(n), batch_size=batch_size, shuffle=shuffle,
                                              collate_fn=lambda x: x)]


def init_linear(n_in, n_out):
    lin = torch.nn.Linear(n_in, n_out)  # default init consumes the generator (weight, then bias)
    return lin


def init_ae_params(n_enc, n_dec, enc_hidden=(256, 128), dec_hidden=(128, 256)):
    """Parameter creation in the reference's order (src/models/ae.py:61-96, :9-28, :35-55): Encoder blocks
    (Linear default init then xavier + zero bias), Decoder blocks likewise, encoder_linear, decoder_linear
    (default init), then xavier/zero on both."""
    p = {}
    k = 0
    for i in range(len(enc_hidden) - 1):
        lin = init_linear(enc_hidden[i], enc_hidden[i + 1])
        p["encoder.blocks.{}".format(k)] = lin
        k += 2
    for key in list(p):
        torch.nn.init.xavier_uniform_(p[key].weight)
        p[key].bias.data.zero_()
    d = {}
    k = 0
    for i in range(len(dec_hidden) - 1):
        lin = init_linear(dec_hidden[i], dec_hidden[i + 1])
        d["decoder.blocks.{}".format(k)] = lin
        k += 2
    for key in d:
        torch.nn.init.xavier_uniform_(d[key].weight)
        d[key].bias.data.zero_()
    p.update(d)
    p["encoder_linear"] = init_linear(n_enc, enc_hidden[0])
    p["decoder_linear"] = init_linear(dec_hidden[-1], n_dec)
    for key in ("encoder_linear", "decoder_linear"):
        torch.nn.init.xavier_uniform_(p[key].weight)
        p[key].bias.data.zero_()
    out = {}
    for key, lin in p.items():
        out[key + ".weight"] = lin.weight.detach().clone()
        out[key + ".bias"] = lin.bias.detach().clone()
    return out


def draw_keep_mask(rows, width=128):
    """nn.Dropout(0.5) on torch CPU == x * bernoulli_(0.5) / 0.5 with the draw taken from the global generator."""
    return torch.empty(rows, width).bernoulli_(0.5)


def run_experiment(data, control, seed=0, local_epochs=20, rounds=10, batch_size=None):
    """Returns per-round global outputs F[t][split], fitted (rate, weight) per round/owner, test metrics and
    organization 0's round-1 parameters."""
    f = control.split("_")
    data_name, data_mode, target_mode, split_mode = f[0], f[1], f[2], f[5]
    ar_mode, ar = f[7].split("-")[0], float(f[7].split("-")[1])
    aw_mode = f[8]
    match_rate = float(f[9]) if len(f) > 9 else 1.0
    pl = f[10] if len(f) > 10 else "none"
    cs = float(f[11]) if len(f) > 11 else None  # cold start: 'cs' in cfg only for 12-field names (src/config.py:12-15)
    K = {"ML100K": 18, "ML1M": 18, "Douban": 3, "Amazon": 4}[data_name] if "genre" in split_mode else int(
        split_mode.split("-")[1])
    bs_table = {"user": {"ML100K": 100, "ML1M": 500, "Douban": 100, "Amazon": 500},
                "item": {"ML100K": 100, "ML1M": 500, "Douban": 1000, "Amazon": 500}}
    bs = batch_size or bs_table[data_mode][data_name]
    torch.manual_seed(seed)
    mats = matrices(data, data_mode, target_mode)
    data_split = split_columns(mats, data.item_attr, split_mode, K, data_mode)
    cols = [s.numpy() for s in data_split]
    n_rows, n_cols = mats["train"][1].shape
    org_data = [{k: mats[k][0][:, c].tocsr() for k in mats} for c in cols]
    org_tgt0 = [{k: mats[k][1][:, c].tocsr() for k in mats} for c in cols]
    y = {k: mats[k][1] for k in mats}  # canonical global CSR, values = ground truth
    start = n_rows
    if cs is not None:
        # src/train_recsys_assist.py:52-56: organization 0 keeps only the first int(n*cs) aligned rows of its TRAIN
        # split; the global train target is assembled from the organizations' targets (:98-141), so organization 0's
        # columns carry no entries below that row
        start = int(n_rows * cs)
        org_data[0]["train"] = org_data[0]["train"][:start]
        org_tgt0[0]["train"] = org_tgt0[0]["train"][:start]
        coo = y["train"].tocoo()
        own0 = np.zeros(n_cols, bool)
        own0[cols[0]] = True
        keep = ~(own0[coo.col] & (coo.row >= start))
        from scipy.sparse import csr_matrix
        y["train"] = csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=y["train"].shape)
        y["train"].sort_indices()
    indices = {k: y[k].indices for k in y}
    views = {k: mtal.owner_views(indices[k], cols, n_cols) for k in y}
    # ---- round 0: Organization.initialize per organization (src/organization.py:29-138) ----
    F = [{k: np.zeros(y[k].nnz, np.float32) for k in y}]
    for i in range(K):
        np.random.seed(seed)
        for _ in range(3):  # train loader iterated twice, test loader once: one base-seed draw each
            loader_batches(1, 1, False)
        preds, _ = train.base_round0(org_data[i]["train"], org_tgt0[i]["train"], org_tgt0[i]["test"], data_mode,
                                     target_mode, bs)
        for k in y:
            # local CSR order of the organization == global order restricted to its columns
            F[0][k][views[k][i][0]] = _to_global_order(org_tgt0[i][k], preds[k])
    def test_metrics(Ft):
        if cs is None:
            return ometrics.evaluate(Ft, y["test"], data_mode, target_mode, bs)
        # src/train_recsys_assist.py:180-182: cold-start runs are scored on organization 0's columns only
        sub = _with_data(y["test"], np.arange(y["test"].nnz, dtype=np.float64))[:, cols[0]].tocsr()
        pos = sub.data.astype(np.int64)
        return ometrics.evaluate(np.asarray(Ft)[pos], _with_values64(sub, y["test"].data[pos]), data_mode, target_mode, bs)

    metrics = {0: test_metrics(F[0]["test"])}
    fitted = [None]
    org0_sd1 = None
    clamp = mtal.needs_clamp(data_name, data_mode, target_mode)
    for t in range(1, rounds + 1):
        # ---- make_dataset (src/assist.py:43-79) ----
        res = {}
        for k in ("train", "test"):
            r = mtal.residual(F[t - 1][k], y[k].data, target_mode, clamp)
            if pl != "none":
                mode, param = pl.split("-")
                assert mode == "dp"
                r = mtal.dp(r, float(param)).astype(np.float32)
            res[k] = r
        tgt = {k: _with_data(y[k], res[k]) for k in y}
        # ---- train (src/organization.py:140-178) ----
        params = []
        for i in range(K):
            np.random.seed(seed)
            p0 = init_ae_params(org_data[i]["train"].shape[1], n_cols)
            epoch_batches, masks = [], []
            # masks must be drawn interleaved with the sampler draws, so walk the epochs now
            n_i = org_data[i]["train"].shape[0]  # organization 0 under cold start walks its truncated row range
            for _ in range(local_epochs):
                batches = loader_batches(n_i, bs, True)
                epoch_batches.append(batches)
                for rows in batches:
                    b = train.make_batch(org_data[i]["train"], tgt["train"], rows, data_mode)
                    if len(b[data_mode]) == 0:
                        continue
                    masks.append(draw_keep_mask(len(models.ae_rows(b, data_mode))))
            p, _ = train.train_org_ae(p0, org_data[i]["train"], tgt["train"], data_mode, target_mode, epoch_batches,
                                      masks)
            params.append(p)
        if t == 1:
            org0_sd1 = params[0]
        # ---- gather / predict (src/organization.py:180-217) ----
        org_out = []
        for i in range(K):
            o = {}
            for k in ("train", "test"):
                np.random.seed(seed)
                init_ae_params(org_data[i][k].shape[1], n_cols)  # predict() builds a fresh model before loading
                loader_batches(1, 1, False)
                o[k] = train.predict_org_ae(params[i], org_data[i][k], tgt[k], data_mode, target_mode, bs)
                if org_data[i][k].shape[0] < n_rows:
                    # rows the organization never saw are absent from its output (src/organization.py:186-216): NaN
                    # here, which is what the padding at src/assist.py:109-111,150-152 produces for the other owners
                    o[k][tgt[k].indptr[org_data[i][k].shape[0]]:] = np.nan
            org_out.append(o)
        # ---- update (src/assist.py:81-179) ----
        Fn, fit = mtal.update(F[t - 1], {k: y[k].data for k in y}, org_out, indices, cols, n_cols, target_mode, ar,
                              ar_mode, aw_mode, match_rate)
        F.append(Fn)
        fitted.append(fit)
        metrics[t] = test_metrics(Fn["test"])
    return {"F": F, "fitted": fitted, "metrics": metrics, "org0_sd1": org0_sd1, "data_split": cols, "y": y}


def _with_data(m, values):
    out = m.copy()
    out.data = np.asarray(values, dtype=np.float32)
    return out


def _with_values64(m, values):
    out = m.copy()
    out.data = np.asarray(values, dtype=np.float32)
    return out


def _to_global_order(local_csr, values):
    """Values given in the organization's local CSR storage order -> global-order-restricted-to-its-columns.
    Both are row-major; within a row scipy's column slice keeps the original storage order, so this is identity."""
    return values
