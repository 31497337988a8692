"""``models.ae`` — the Assisted AutoEncoder (reference src/models/ae.py:9-170) on libdmt_b200 kernels.

Same module tree and parameter names as the reference (``encoder_linear``, ``encoder.blocks.0``,
``decoder.blocks.0``, ``decoder_linear``, optional ``user_profile`` / ``item_attr`` encoders) so state_dicts,
checkpoints and ``models.distribute`` keep working; same construction order so torch's generator is consumed
identically. ``forward`` is a chain of three kernel-backed autograd Functions (CSR SpMM encoder, dense tanh
layers with fused dropout, SDDMM decoder fused with the loss); the device-resident engine
(``dmtcdr_b200.engine``) runs the same kernels without Python in the loop.
"""
import torch
import torch.nn as nn

from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from . import _ops


class _TanhStack(nn.Module):
    """Linear+Tanh blocks held in ``self.blocks`` (an nn.Sequential, as in the reference's Encoder/Decoder)."""

    def __init__(self, sizes):
        super().__init__()
        layers = []
        for n_in, n_out in zip(sizes[:-1], sizes[1:]):
            layers += [nn.Linear(n_in, n_out), nn.Tanh()]
        self.blocks = nn.Sequential(*layers)
        for m in self.blocks:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                m.bias.data.zero_()

    def linears(self):
        return [m for m in self.blocks if isinstance(m, nn.Linear)]

    def forward(self, x, keep_last=None, scale=1.0):
        lins = self.linears()
        for i, m in enumerate(lins):
            last = i == len(lins) - 1
            x = _ops.dense(x, m.weight, m.bias, 1, keep_last if last else None, scale if last else 1.0)
        return x


class Encoder(_TanhStack):
    def __init__(self, input_size, hidden_size):
        super().__init__([input_size] + list(hidden_size))
        self.input_size, self.hidden_size = input_size, hidden_size


class Decoder(_TanhStack):
    def __init__(self, output_size, hidden_size):
        super().__init__(list(hidden_size) + [output_size])
        self.output_size, self.hidden_size = output_size, hidden_size


class AE(nn.Module):
    def __init__(self, encoder_num_users, encoder_num_items, decoder_num_users, decoder_num_items, encoder_hidden_size,
                 decoder_hidden_size, info_size):
        super().__init__()
        self.info_size = info_size
        if len(encoder_hidden_size) > 1:
            self.encoder = Encoder(encoder_hidden_size[0], encoder_hidden_size[1:])
            self.decoder = Decoder(decoder_hidden_size[-1], decoder_hidden_size[:-1])
        else:
            self.encoder = nn.Identity()
            self.decoder = nn.Identity()
        if cfg['data_mode'] == 'user':
            n_in, n_out = encoder_num_items, decoder_num_items
        elif cfg['data_mode'] == 'item':
            n_in, n_out = encoder_num_users, decoder_num_users
        else:
            raise ValueError('Not valid data mode')
        self.encoder_linear = nn.Linear(n_in, encoder_hidden_size[0])
        self.decoder_linear = nn.Linear(decoder_hidden_size[-1], n_out)
        self.dropout = nn.Dropout(p=0.5)
        if info_size is not None:
            if 'user_profile' in info_size:
                self.user_profile = Encoder(info_size['user_profile'], encoder_hidden_size)
            if 'item_attr' in info_size:
                self.item_attr = Encoder(info_size['item_attr'], encoder_hidden_size)
        for lin in (self.encoder_linear, self.decoder_linear):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.zero_()
        self.keep_override = None  # tests inject the reference's Bernoulli draw here

    def _keep_mask(self, rows, width, device):
        """0/1 keep mask of nn.Dropout(p): Bernoulli(1-p) drawn with torch's generator of the model's device."""
        if self.keep_override is not None:
            k = self.keep_override
            self.keep_override = None
            return k.to(device=device, dtype=torch.uint8).contiguous()
        return torch.empty(rows, width, device=device).bernoulli_(1.0 - self.dropout.p).to(torch.uint8)

    def forward(self, input):
        mode = cfg['data_mode']
        if mode not in ('user', 'item'):
            raise ValueError('Not valid data mode')
        other = 'item' if mode == 'user' else 'user'
        ids, t_ids = input[mode], input['target_' + mode]
        rows = torch.unique(torch.cat([ids, t_ids]), sorted=True)  # row space of the batch (ae.py:101)
        d_ptr, d_idx, d_val, _, d_row = _ops.batch_csr(ids, input[other], input['rating'], rows)
        t_ptr, t_idx, t_val, t_order, t_row = _ops.batch_csr(t_ids, input['target_' + other], input['target_rating'],
                                                             rows)
        x = _ops.SparseEncoderFn.apply(self.encoder_linear.weight, self.encoder_linear.bias, d_ptr, d_idx, d_val, d_row)
        drop = self.training and self.dropout.p > 0
        has_info = self.info_size is not None and (('user_profile' in input and hasattr(self, 'user_profile')) or
                                                   ('item_attr' in input and hasattr(self, 'item_attr')))
        scale = 1.0 / (1.0 - self.dropout.p) if drop else 1.0
        stacked = isinstance(self.encoder, Encoder)
        width = self.encoder.linears()[-1].out_features if stacked else x.shape[1]
        keep = self._keep_mask(rows.numel(), width, x.device) if drop else None
        fuse_drop = drop and stacked and not has_info  # dropout rides in the last encoder GEMM's epilogue
        if stacked:
            x = self.encoder(x, keep if fuse_drop else None, scale)
        if has_info:
            if 'user_profile' in input and hasattr(self, 'user_profile'):
                x = x + self.user_profile(input['user_profile'])
            if 'item_attr' in input and hasattr(self, 'item_attr'):
                x = x + self.item_attr(input['item_attr'])
        if drop and not fuse_drop:
            x = x * (keep.to(x.dtype) * scale)
        if stacked:
            x = self.decoder(x)
        local = bool(input['local']) if 'local' in input else False
        kind = native.LOSS_KIND['explicit'] if local else native.LOSS_KIND[cfg['target_mode']]
        pred_sorted, loss = _ops.SparseDecoderLossFn.apply(x, self.decoder_linear.weight, self.decoder_linear.bias,
                                                           t_ptr, t_idx, t_val, t_row, kind)
        pred = torch.empty_like(pred_sorted)
        pred[t_order] = pred_sorted  # back to the order of input['target_*']
        return {'target_rating': pred, 'loss': loss}


def ae(encoder_num_users=None, encoder_num_items=None, decoder_num_users=None, decoder_num_items=None):
    encoder_num_users = cfg['num_users']['data'] if encoder_num_users is None else encoder_num_users
    encoder_num_items = cfg['num_items']['data'] if encoder_num_items is None else encoder_num_items
    decoder_num_users = cfg['num_users']['target'] if decoder_num_users is None else decoder_num_users
    decoder_num_items = cfg['num_items']['target'] if decoder_num_items is None else decoder_num_items
    return AE(encoder_num_users, encoder_num_items, decoder_num_users, decoder_num_items,
              cfg['ae']['encoder_hidden_size'], cfg['ae']['decoder_hidden_size'], cfg['info_size'])
