"""``models.assist`` (reference src/models/assist.py:7-49): assisted learning rate (one per owned column) and
assistance weights; forward = h + rate[idx] * (O . softmax(w)), with the cold-start branch (rows whose slot 0 is NaN
combine slots 1.. under softmax(w[1:]) and move to the end of the result, :28-34).

The module is differentiable: the reference's fit (closure at src/assist.py:121-126) calls ``loss.backward()`` through
it, so ``assist_rate`` / ``assist_weight`` receive ``.grad`` when they are Parameters. Forward, loss and backward run in
libdmt_b200 kernels (dmt_assist_rows_fwd / dmt_loss_fwd / dmt_assist_rows_bwd); torch only carries the tensors. The
MTAL coordinator (dropin/assist.py) applies the same arithmetic for all owners at once with dmt_assist_combine."""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from . import _ops


class AssistRowsFn(torch.autograd.Function):
    """target = history + rate[idx] * sum_j softmax(w)_j out[:, j] (NaN-aware); grads for rate and weight."""

    @staticmethod
    def forward(ctx, out, history, idx32, rate, weight, seg_of):
        n, K = out.shape
        train = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        tgt, q = native.assist_rows_fwd(out, K, 1, history, idx32, rate.detach().contiguous(),
                                        weight.detach().contiguous(), n, K, want_q=train)
        if train:
            ctx.save_for_backward(out, idx32, rate.detach().contiguous(), weight.detach().contiguous(), q)
            ctx.seg_of = seg_of
        return tgt

    @staticmethod
    def backward(ctx, dtgt):
        out, idx32, rate, weight, q = ctx.saved_tensors
        n, K = out.shape
        want_rate, want_w = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        seg = ctx.seg_of() if want_rate else (None, None, None, None)
        d_rate, d_w = native.assist_rows_bwd(out, K, 1, idx32, rate, weight, q, dtgt.contiguous(), seg, n, K,
                                             want_rate=want_rate, want_w=want_w)
        return None, None, None, d_rate, d_w, None


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
        self._cache = {}  # per input tensor: int32 indices, their sorted segments, the cold-start ordering

    def _prepared(self, out, idx):
        """Index-derived helpers are cached per (output, output_idx) storage: the L-BFGS closure evaluates the same
        input dict up to 25 times per step (src/assist.py:118-129)."""
        key = (out.data_ptr(), idx.data_ptr(), out.shape[0])
        hit = self._cache.get(key)
        if hit is None:
            self._cache.clear()
            idx32 = idx.to(torch.int32).contiguous()
            out_c = out.to(torch.float32).contiguous()
            cold = torch.isnan(out_c[:, 0])
            order = None
            if bool(cold.any()):  # the reference returns cat(warm rows, cold rows): a stable partition
                order = torch.argsort(cold.to(torch.int8), stable=True)
                if bool((order[1:] > order[:-1]).all()):
                    order = None  # cold rows already at the end (the driver's layout): identity
            hit = {'idx32': idx32, 'out': out_c, 'order': order, 'seg': None}
            self._cache[key] = hit
        return hit

    def forward(self, input):
        out, h, idx = input['output'], input['history'], input['output_idx']
        if not out.is_cuda:
            raise native.NativeError('models.assist runs on CUDA tensors only')
        prep = self._prepared(out, idx)

        def seg_of():
            if prep['seg'] is None:
                prep['seg'] = native.sort_segments(prep['idx32'], self.assist_rate.numel())
            return prep['seg']

        target = AssistRowsFn.apply(prep['out'], h.to(torch.float32).contiguous(), prep['idx32'], self.assist_rate,
                                    self.assist_weight, seg_of)
        if prep['order'] is not None:
            target = target[prep['order']]
        output = {'target': target}
        if 'target' in input:
            output['loss'] = _ops.LossFn.apply(target, input['target'].to(torch.float32),
                                               native.LOSS_KIND[cfg['target_mode']])
        return output


def assist(num_outputs):
    return Assist(cfg['assist']['ar'], cfg['assist']['ar_mode'], num_outputs, cfg['num_organizations'],
                  cfg['assist']['aw_mode'])
