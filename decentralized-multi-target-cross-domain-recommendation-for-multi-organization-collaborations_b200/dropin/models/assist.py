"""``models.assist`` (reference src/models/assist.py:7-49): assisted learning rate (one per owned column) and
assistance weights; forward = h + rate[idx] * (O . softmax(w)). The MTAL coordinator (dropin/assist.py) applies it
for all owners at once with dmt_assist_combine; this module is the per-owner form used by the fit and the tests."""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from .utils import loss_fn


class Assist(nn.Module):
    def __init__(self, ar, ar_mode, num_outputs, num_organizations, aw_mode):
        super().__init__()
        self.ar_mode, self.aw_mode = ar_mode, aw_mode
        rate = torch.full((num_outputs,), ar)
        weight = torch.ones(num_organizations) / num_organizations
        if ar_mode == 'optim':
            self.assist_rate = nn.Parameter(rate)
        elif ar_mode == 'constant':
            self.register_buffer('assist_rate', rate)
        else:
            raise ValueError('Not valid ar mode')
        if aw_mode == 'optim':
            self.assist_weight = nn.Parameter(weight)
        elif aw_mode == 'constant':
            self.register_buffer('assist_weight', weight)
        else:
            raise ValueError('Not valid aw mode')

    def forward(self, input):
        out = input['output']
        if torch.isnan(out).any():
            raise NotImplementedError("cold-start ('cs') NaN padding is out of scope (DESIGN.md)")
        h, idx = input['history'], input['output_idx']
        n, K = out.shape
        if not out.is_cuda:
            raise native.NativeError('models.assist runs on CUDA tensors only')
        # single-owner use of the all-owner kernel: one pseudo owner that owns every column
        O = out.t().contiguous()
        col = idx.to(torch.int32).contiguous()
        owner = torch.zeros(self.assist_rate.numel(), dtype=torch.int32, device=out.device)
        S = torch.softmax(self.assist_weight.detach(), -1).repeat(K, 1).contiguous()
        target = native.assist_combine(h.contiguous(), O, col, owner, self.assist_rate.detach().contiguous(), S)
        output = {'target': target}
        if 'target' in input:
            output['loss'] = loss_fn(output['target'], input['target'])
        return output


def assist(num_outputs):
    return Assist(cfg['assist']['ar'], cfg['assist']['ar_mode'], num_outputs, cfg['num_organizations'],
                  cfg['assist']['aw_mode'])
