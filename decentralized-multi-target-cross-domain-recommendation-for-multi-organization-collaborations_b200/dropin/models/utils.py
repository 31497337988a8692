"""``models.loss_fn`` and ``models.distribute`` of the reference (src/models/utils.py:7-40)."""
import torch
import torch.nn.functional as F

from dmtcdr_b200.config import cfg


def loss_fn(output, target, reduction='mean'):
    """implicit -> BCE with logits, explicit -> MSE (reference src/models/utils.py:7-14). The fused model kernels
    compute their own losses; this entry point exists for the drivers' evaluation code, which calls it on small
    host tensors (src/train_recsys_assist.py:204)."""
    if cfg['target_mode'] == 'implicit':
        return F.binary_cross_entropy_with_logits(output, target, reduction=reduction)
    if cfg['target_mode'] == 'explicit':
        return F.mse_loss(output, target, reduction=reduction)
    raise ValueError('Not valid target mode')


def distribute(model, local_model, data_split):
    """Copy a jointly trained model into the per-organization models: tables indexed by the split entity are
    row-sliced with the organization's ids, everything else is copied whole (reference src/models/utils.py:17-40)."""
    name = cfg['model_name']
    split_key = {'user': 'item', 'item': 'user'}[cfg['data_mode']]
    with torch.no_grad():
        for k, v in model.state_dict().items():
            for i, local in enumerate(local_model):
                dst = local.state_dict()[k]
                if name == 'base' or (name in ('mf', 'mlp', 'nmf') and split_key in k):
                    dst.copy_(v[data_split[i].to(v.device)])
                elif name in ('mf', 'mlp', 'nmf', 'ae'):
                    dst.copy_(v)
                else:
                    raise ValueError('Not valid model')
