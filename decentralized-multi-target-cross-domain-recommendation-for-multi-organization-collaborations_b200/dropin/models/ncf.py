"""``models.mlp`` / ``models.nmf`` (reference src/models/mlp.py, nmf.py) — filled in by ncf kernels (see below)."""
from dmtcdr_b200.config import cfg


class MLP:  # replaced below once the NCF kernels are wired
    pass


class NMF:
    pass


def mlp(num_users=None, num_items=None):
    raise NotImplementedError('models.mlp: NCF tower kernels are not wired yet')


def nmf(num_users=None, num_items=None):
    raise NotImplementedError('models.nmf: NCF tower kernels are not wired yet')
