"""``models.base``: the round-0 mean predictor (reference src/models/base.py:9-67) on libdmt_b200 kernels."""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from .utils import loss_fn


class Base(nn.Module):
    def __init__(self, num_users, num_items):
        super().__init__()
        self.num_users = num_users
        self.num_items = num_items
        if cfg['data_mode'] == 'user':
            size = num_items
        elif cfg['data_mode'] == 'item':
            size = num_users
        else:
            raise ValueError('Not valid data mode')
        self.register_buffer('base', torch.zeros(size))
        self.register_buffer('count', torch.zeros(size))

    def forward(self, input):
        mode = cfg['data_mode']
        if mode not in ('user', 'item'):
            raise ValueError('Not valid data mode')
        col_key = 'item' if mode == 'user' else 'user'
        implicit = cfg['target_mode'] == 'implicit'
        if cfg['target_mode'] not in ('explicit', 'implicit'):
            raise ValueError('Not valid target mode')
        if self.training:
            idx = input[col_key].to(torch.int32).contiguous()
            # count is only touched by the kernel in explicit mode; in implicit mode every entry grows by the number
            # of distinct row entities of the batch (reference src/models/base.py:35-37)
            scratch = self.count if not implicit else torch.zeros_like(self.count)
            native.base_fit(idx, input['rating'].contiguous(), self.base, scratch)
            if implicit:
                self.count = self.count + torch.unique(input[mode]).size(0)
        tidx = input['target_' + col_key].to(torch.int32).contiguous()
        imp_count = float(self.count[0]) if implicit and self.count.numel() else 0.0
        output = {'target_rating': native.base_predict(self.base, self.count, tidx, implicit, imp_count)}
        output['loss'] = loss_fn(output['target_rating'], input['target_rating'])
        return output


def base(num_users=None, num_items=None):
    num_users = cfg['num_users']['data'] if num_users is None else num_users
    num_items = cfg['num_items']['data'] if num_items is None else num_items
    return Base(num_users, num_items)
