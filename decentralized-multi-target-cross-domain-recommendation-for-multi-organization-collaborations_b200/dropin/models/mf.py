"""``models.mf`` (reference src/models/mf.py:9-102): embedding MF whose forward/backward run in the fused
gather+dot+bias+loss kernel and the sort + segmented-reduction gradient kernels of libdmt_b200."""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from . import _ops


class MF(nn.Module):
    def __init__(self, num_users, num_items, hidden_size, info_size):
        super().__init__()
        self.num_users, self.num_items = num_users, num_items
        self.hidden_size, self.info_size = hidden_size, info_size
        # creation order and init calls follow the reference so that torch's generator is consumed identically
        # (src/models/mf.py:16-34): 4 embeddings, randn(1) bias, optional side-info Linears, then the re-init.
        self.user_weight = nn.Embedding(num_users, hidden_size)
        self.item_weight = nn.Embedding(num_items, hidden_size)
        self.user_bias = nn.Embedding(num_users, 1)
        self.item_bias = nn.Embedding(num_items, 1)
        self.bias = nn.Parameter(torch.randn(1))
        if info_size is not None:
            if 'user_profile' in info_size:
                self.user_profile = nn.Linear(info_size['user_profile'], hidden_size)
            if 'item_attr' in info_size:
                self.item_attr = nn.Linear(info_size['item_attr'], hidden_size)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.user_weight.weight, 0.0, 0.01)
        nn.init.normal_(self.item_weight.weight, 0.0, 0.01)
        for p in (self.user_bias.weight, self.item_bias.weight, self.bias):
            nn.init.zeros_(p)

    def forward(self, input):
        pre = '' if self.training else 'target_'
        user, item, rating = input[pre + 'user'], input[pre + 'item'], input[pre + 'rating']
        pu = pi = None
        if self.info_size is not None:
            if pre + 'user_profile' in input:
                pu = _ops.dense(input[pre + 'user_profile'], self.user_profile.weight, self.user_profile.bias)
            if pre + 'item_attr' in input:
                pi = _ops.dense(input[pre + 'item_attr'], self.item_attr.weight, self.item_attr.bias)
        if hasattr(self, 'num_matched'):
            raise NotImplementedError('shared-embedding MDR baseline (models/mdr.py) is out of scope')
        pred, loss = _ops.MFFn.apply(user.to(torch.int32).contiguous(), item.to(torch.int32).contiguous(),
                                     rating.contiguous(), self.user_weight.weight, self.item_weight.weight,
                                     self.user_bias.weight, self.item_bias.weight, self.bias, pu, pi,
                                     native.LOSS_KIND[cfg['target_mode']])
        return {'target_rating': pred, 'loss': loss}


def mf(num_users=None, num_items=None):
    num_users = cfg['num_users']['data'] if num_users is None else num_users
    num_items = cfg['num_items']['data'] if num_items is None else num_items
    return MF(num_users, num_items, cfg['mf']['hidden_size'], cfg['info_size'])
