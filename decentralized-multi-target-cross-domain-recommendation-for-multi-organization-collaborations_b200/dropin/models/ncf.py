"""``models.mlp`` and ``models.nmf`` (reference src/models/mlp.py:9-120, nmf.py:9-156): NCF on libdmt_b200 kernels.

Same parameters, names and construction order as the reference. Forward = fused embedding gathers written straight
into the tower input, dense ReLU layers, and (NMF) the GMF product + affine layer + loss in one kernel; backward =
dense-layer kernels and sort-by-index segmented reductions into the embedding tables.
"""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from . import _ops


def _tower(hidden_size, info_size):
    """fc stack: first layer takes [user, item (, profile, attr)] embeddings, ReLU after every layer
    (reference src/models/mlp.py:26-38, nmf.py:31-44)."""
    layers = []
    for i in range(len(hidden_size) - 1):
        if i == 0:
            n_in = 2 * hidden_size[0]
            if info_size is not None:
                n_in += hidden_size[0] * (('user_profile' in info_size) + ('item_attr' in info_size))
        else:
            n_in = hidden_size[i]
        layers += [nn.Linear(n_in, hidden_size[i + 1]), nn.ReLU()]
    return nn.Sequential(*layers)


def _run_tower(fc, x):
    for m in fc:
        if isinstance(m, nn.Linear):
            x = _ops.dense(x, m.weight, m.bias, 2)
    return x


def _pick(model, input):
    pre = '' if model.training else 'target_'
    user, item, rating = input[pre + 'user'], input[pre + 'item'], input[pre + 'rating']
    profile = attr = None
    if model.info_size is not None:
        profile = input.get(pre + 'user_profile')
        attr = input.get(pre + 'item_attr')
    if hasattr(model, 'num_matched'):
        raise NotImplementedError('shared-embedding MDR baseline (models/mdr.py) is out of scope')
    return user.to(torch.int32).contiguous(), item.to(torch.int32).contiguous(), rating.contiguous(), profile, attr


class MLP(nn.Module):
    def __init__(self, num_users, num_items, hidden_size, info_size):
        super().__init__()
        self.num_users, self.num_items = num_users, num_items
        self.hidden_size, self.info_size = hidden_size, info_size
        self.user_weight = nn.Embedding(num_users, hidden_size[0])
        self.item_weight = nn.Embedding(num_items, hidden_size[0])
        self.user_bias = nn.Embedding(num_users, 1)
        self.item_bias = nn.Embedding(num_items, 1)
        if info_size is not None:
            if 'user_profile' in info_size:
                self.user_profile = nn.Linear(info_size['user_profile'], hidden_size[0])
            if 'item_attr' in info_size:
                self.item_attr = nn.Linear(info_size['item_attr'], hidden_size[0])
        self.fc = _tower(hidden_size, info_size)
        self.affine = nn.Linear(hidden_size[-1], 1)
        nn.init.normal_(self.user_weight.weight, 0.0, 0.01)
        nn.init.normal_(self.item_weight.weight, 0.0, 0.01)
        nn.init.zeros_(self.user_bias.weight)
        nn.init.zeros_(self.item_bias.weight)
        for m in self.fc:
            if isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)
        nn.init.zeros_(self.affine.bias)

    def forward(self, input):
        user, item, rating, profile, attr = _pick(self, input)
        x = _ops.EmbedCatFn.apply(user, item, self.user_weight.weight, self.user_bias.weight, self.item_weight.weight,
                                  self.item_bias.weight)
        extra = []
        if profile is not None:
            extra.append(_ops.dense(profile, self.user_profile.weight, self.user_profile.bias))
        if attr is not None:
            extra.append(_ops.dense(attr, self.item_attr.weight, self.item_attr.bias))
        if extra:
            x = torch.cat([x] + extra, dim=-1)
        x = _run_tower(self.fc, x)
        pred = _ops.dense(x, self.affine.weight, self.affine.bias).view(-1)
        loss = _ops.LossFn.apply(pred, rating, native.LOSS_KIND[cfg['target_mode']])
        return {'target_rating': pred, 'loss': loss}


class NMF(nn.Module):
    def __init__(self, num_users, num_items, hidden_size, info_size):
        super().__init__()
        self.num_users, self.num_items = num_users, num_items
        self.hidden_size, self.info_size = hidden_size, info_size
        H = hidden_size[0]
        self.user_weight_mlp = nn.Embedding(num_users, H)
        self.item_weight_mlp = nn.Embedding(num_items, H)
        self.user_bias_mlp = nn.Embedding(num_users, 1)
        self.item_bias_mlp = nn.Embedding(num_items, 1)
        self.user_weight_mf = nn.Embedding(num_users, H)
        self.item_weight_mf = nn.Embedding(num_items, H)
        self.user_bias_mf = nn.Embedding(num_users, 1)
        self.item_bias_mf = nn.Embedding(num_items, 1)
        if info_size is not None:
            if 'user_profile' in info_size:
                self.user_profile_mf = nn.Linear(info_size['user_profile'], H)
                self.user_profile_mlp = nn.Linear(info_size['user_profile'], H)
            if 'item_attr' in info_size:
                self.item_attr_mf = nn.Linear(info_size['item_attr'], H)
                self.item_attr_mlp = nn.Linear(info_size['item_attr'], H)
        self.fc = _tower(hidden_size, info_size)
        self.affine = nn.Linear(hidden_size[-1] + H, 1)
        for w in (self.user_weight_mlp, self.item_weight_mlp):
            nn.init.normal_(w.weight, 0.0, 0.01)
        nn.init.zeros_(self.user_bias_mlp.weight)
        nn.init.zeros_(self.item_bias_mlp.weight)
        for w in (self.user_weight_mf, self.item_weight_mf):
            nn.init.normal_(w.weight, 0.0, 0.01)
        nn.init.zeros_(self.user_bias_mf.weight)
        nn.init.zeros_(self.item_bias_mf.weight)
        for m in self.fc:
            if isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)
        nn.init.zeros_(self.affine.bias)

    def forward(self, input):
        user, item, rating, profile, attr = _pick(self, input)
        x = _ops.EmbedCatFn.apply(user, item, self.user_weight_mlp.weight, self.user_bias_mlp.weight,
                                  self.item_weight_mlp.weight, self.item_bias_mlp.weight)
        extra = []
        pu = pi = None
        if profile is not None:
            pu = _ops.dense(profile, self.user_profile_mf.weight, self.user_profile_mf.bias)
            extra.append(_ops.dense(profile, self.user_profile_mlp.weight, self.user_profile_mlp.bias))
        if attr is not None:
            pi = _ops.dense(attr, self.item_attr_mf.weight, self.item_attr_mf.bias)
            extra.append(_ops.dense(attr, self.item_attr_mlp.weight, self.item_attr_mlp.bias))
        if extra:
            x = torch.cat([x] + extra, dim=-1)
        x = _run_tower(self.fc, x)
        n_tower = self.hidden_size[-1]
        # affine([tower, gmf]) = tower . a[:32] + a0  +  gmf . a[32:]: the second term is fused with the GMF product
        logit_tower = _ops.dense(x, self.affine.weight[:, :n_tower], self.affine.bias).view(-1)
        pred, loss = _ops.GMFLossFn.apply(user, item, rating, self.user_weight_mf.weight, self.item_weight_mf.weight,
                                          self.user_bias_mf.weight, self.item_bias_mf.weight, pu, pi,
                                          self.affine.weight[0, n_tower:], logit_tower,
                                          native.LOSS_KIND[cfg['target_mode']])
        return {'target_rating': pred, 'loss': loss}


def mlp(num_users=None, num_items=None):
    num_users = cfg['num_users']['data'] if num_users is None else num_users
    num_items = cfg['num_items']['data'] if num_items is None else num_items
    return MLP(num_users, num_items, cfg['mlp']['hidden_size'], cfg['info_size'])


def nmf(num_users=None, num_items=None):
    num_users = cfg['num_users']['data'] if num_users is None else num_users
    num_items = cfg['num_items']['data'] if num_items is None else num_items
    return NMF(num_users, num_items, cfg['nmf']['hidden_size'], cfg['info_size'])
