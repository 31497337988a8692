"""Oracle forward passes (torch CPU fp32; gradients come from autograd on these expressions).

TEST INFRASTRUCTURE (see oracle/__init__.py). Parameter dicts use the reference's state_dict names.
The formulations are deliberately the textbook dense ones (one-hot / dense matmul), not the
reference's sort/unique/index_add pipeline: same math, different summation order (fp32 tolerance).
"""
import torch


def loss_fn(output, target, target_mode, reduction="mean"):
    """reference src/models/utils.py:7-14 — implicit: BCE-with-logits, explicit: MSE."""
    if target_mode == "implicit":
        # stable log(1+exp(-|x|)) form of  -[y log s(x) + (1-y) log(1-s(x))]
        per = torch.clamp(output, min=0) - output * target + torch.log1p(torch.exp(-output.abs()))
    elif target_mode == "explicit":
        per = (output - target) ** 2
    else:
        raise ValueError("Not valid target mode")
    if reduction == "mean":
        return per.mean()
    if reduction == "sum":
        return per.sum()
    raise ValueError("Not valid reduction")


def _emb(p, w, b, idx):
    """Embedding with the bias broadcast-added to every hidden dim BEFORE the product
    (reference src/models/mf.py:36-48, nmf.py:62-88, mlp.py:52-64)."""
    return p[w][idx] + p[b][idx]


def _linear(p, name, x):
    return x @ p[name + ".weight"].t() + p[name + ".bias"]


def mf_forward(p, user, item, rating, target_mode, user_profile=None, item_attr=None):
    """reference src/models/mf.py:57-93:  pred = sum_d (Wu+bu)(Wi+bi) [+ u.(P Wp^T+bp)] [+ i.(A Wa^T+ba)] + b."""
    u = _emb(p, "user_weight.weight", "user_bias.weight", user)
    i = _emb(p, "item_weight.weight", "item_bias.weight", item)
    s = u * i
    if user_profile is not None:
        s = s + u * _linear(p, "user_profile", user_profile)
    if item_attr is not None:
        s = s + i * _linear(p, "item_attr", item_attr)
    pred = s.sum(-1) + p["bias"]
    return pred, loss_fn(pred, rating, target_mode)


def _tower(p, h):
    k = 0
    while "fc.{}.weight".format(k) in p:
        h = torch.relu(h @ p["fc.{}.weight".format(k)].t() + p["fc.{}.bias".format(k)])
        k += 2
    return h


def mlp_forward(p, user, item, rating, target_mode, user_profile=None, item_attr=None):
    """reference src/models/mlp.py:74-111: ReLU tower on [u, i (, info)] then affine(32->1)."""
    u = _emb(p, "user_weight.weight", "user_bias.weight", user)
    i = _emb(p, "item_weight.weight", "item_bias.weight", item)
    parts = [u, i]
    if user_profile is not None:
        parts.append(_linear(p, "user_profile", user_profile))
    if item_attr is not None:
        parts.append(_linear(p, "item_attr", item_attr))
    h = _tower(p, torch.cat(parts, -1))
    pred = _linear(p, "affine", h).view(-1)
    return pred, loss_fn(pred, rating, target_mode)


def nmf_forward(p, user, item, rating, target_mode, user_profile=None, item_attr=None):
    """reference src/models/nmf.py:100-147: GMF product (+ side-info products) and the MLP tower, affine(160->1)."""
    u_mlp = _emb(p, "user_weight_mlp.weight", "user_bias_mlp.weight", user)
    i_mlp = _emb(p, "item_weight_mlp.weight", "item_bias_mlp.weight", item)
    u_mf = _emb(p, "user_weight_mf.weight", "user_bias_mf.weight", user)
    i_mf = _emb(p, "item_weight_mf.weight", "item_bias_mf.weight", item)
    g = u_mf * i_mf
    parts = [u_mlp, i_mlp]
    if user_profile is not None:
        g = g + u_mf * _linear(p, "user_profile_mf", user_profile)
        parts.append(_linear(p, "user_profile_mlp", user_profile))
    if item_attr is not None:
        g = g + i_mf * _linear(p, "item_attr_mf", item_attr)
        parts.append(_linear(p, "item_attr_mlp", item_attr))
    h = _tower(p, torch.cat(parts, -1))
    pred = _linear(p, "affine", torch.cat([h, g], -1)).view(-1)
    return pred, loss_fn(pred, rating, target_mode)


def pair_forward(model_name, p, batch, target_mode, training):
    """Train mode reads user/item/rating (+user_profile/item_attr), eval reads the target_* keys
    (reference src/models/mf.py:59-78 and the same block in mlp.py / nmf.py)."""
    pre = "" if training else "target_"
    fn = {"mf": mf_forward, "mlp": mlp_forward, "nmf": nmf_forward}[model_name]
    return fn(p, batch[pre + "user"], batch[pre + "item"], batch[pre + "rating"], target_mode,
              batch.get(pre + "user_profile"), batch.get(pre + "item_attr"))


def ae_rows(batch, data_mode):
    """Row space of an AE batch: sorted unique ids of the aligned entity over data AND target entries
    (reference src/models/ae.py:101,112)."""
    return torch.unique(torch.cat([batch[data_mode], batch["target_" + data_mode]]), sorted=True)


def ae_forward(p, batch, data_mode, target_mode, training, keep_mask=None, local=False, p_drop=0.5):
    """reference src/models/ae.py:98-157 as dense algebra:
       a1 = tanh(X W1^T + b1), X[r, c] = rating  (rows without data get tanh(b1), ae.py:108-110)
       a2 = tanh(a1 W2^T + b2); c = a2 * keep / (1 - p) in training; a3 = tanh(c W3^T + b3)
       out_f = a3[r_f] . W4[c_f] + b4[c_f];  loss = MSE if local else loss_fn (ae.py:153-156).
    ``keep_mask`` is the 0/1 Bernoulli(1-p) draw of nn.Dropout, [B' x 128]."""
    other = "item" if data_mode == "user" else "user"
    rows = ae_rows(batch, data_mode)
    W1 = p["encoder_linear.weight"]
    X = torch.zeros(len(rows), W1.shape[1], dtype=W1.dtype)
    r = torch.searchsorted(rows, batch[data_mode])
    X.index_put_((r, batch[other]), batch["rating"], accumulate=True)
    h = torch.tanh(X @ W1.t() + p["encoder_linear.bias"])
    k = 0
    while "encoder.blocks.{}.weight".format(k) in p:
        h = torch.tanh(h @ p["encoder.blocks.{}.weight".format(k)].t() + p["encoder.blocks.{}.bias".format(k)])
        k += 2
    if "user_profile" in batch and "user_profile.blocks.0.weight" in p:
        h = h + _side_encoder(p, "user_profile", batch["user_profile"])
    if "item_attr" in batch and "item_attr.blocks.0.weight" in p:
        h = h + _side_encoder(p, "item_attr", batch["item_attr"])
    if training:
        h = h * keep_mask.to(h.dtype) / (1.0 - p_drop)
    k = 0
    while "decoder.blocks.{}.weight".format(k) in p:
        h = torch.tanh(h @ p["decoder.blocks.{}.weight".format(k)].t() + p["decoder.blocks.{}.bias".format(k)])
        k += 2
    rt = torch.searchsorted(rows, batch["target_" + data_mode])
    ct = batch["target_" + other]
    pred = (h[rt] * p["decoder_linear.weight"][ct]).sum(-1) + p["decoder_linear.bias"][ct]
    y = batch["target_rating"]
    loss = ((pred - y) ** 2).mean() if local else loss_fn(pred, y, target_mode)
    return pred, loss


def _side_encoder(p, name, x):
    k = 0
    while "{}.blocks.{}.weight".format(name, k) in p:
        x = torch.tanh(x @ p["{}.blocks.{}.weight".format(name, k)].t() + p["{}.blocks.{}.bias".format(name, k)])
        k += 2
    return x


class Base:
    """Round-0 predictor, reference src/models/base.py:9-60: per-index mean of the seen ratings.
    explicit: count per index, unseen -> mean of the seen means; implicit: the count of EVERY index grows by
    the number of distinct row-entities in each fitted batch (base.py:35-37,51-53)."""

    def __init__(self, size, target_mode):
        self.base = torch.zeros(size)
        self.count = torch.zeros(size)
        self.target_mode = target_mode

    def fit(self, idx, rating, row_entity):
        self.base.index_add_(0, idx, rating)
        if self.target_mode == "explicit":
            self.count.index_add_(0, idx, torch.ones_like(rating))
        else:
            self.count = self.count + float(len(torch.unique(row_entity)))

    def predict(self, target_idx):
        c = self.count[target_idx]
        if self.target_mode == "explicit":
            out = self.base[target_idx] / (c + 1e-10)
            seen = self.count != 0
            out[c == 0] = (self.base[seen] / self.count[seen]).mean()
            return out
        return self.base[target_idx] / c


def assist_forward(rate, weight, history, output, output_idx, target=None, target_mode=None):
    """reference src/models/assist.py:25-40: F = h + rate[idx] * sum_j softmax(w)_j O[:, j]; rows whose column 0
    is NaN (cold start) use softmax(w[1:]) over O[:, 1:] and are moved to the END of the result."""
    r = rate[output_idx]
    if torch.isnan(output).any():
        nan = torch.isnan(output[:, 0])
        s = history[~nan] + r[~nan] * (output[~nan] * torch.softmax(weight, -1)).sum(-1)
        c = history[nan] + r[nan] * (output[nan][:, 1:] * torch.softmax(weight[1:], -1)).sum(-1)
        pred = torch.cat([s, c], 0)
    else:
        pred = history + r * (output * torch.softmax(weight, -1)).sum(-1)
    loss = loss_fn(pred, target, target_mode) if target is not None else None
    return pred, loss
