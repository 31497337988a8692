"""Drop-in ``models`` package: same constructors and forward contract as the reference's ``src/models``
(``forward(input: dict) -> {'target_rating' | 'target', 'loss'}``), arithmetic in libdmt_b200 kernels."""
from .utils import loss_fn, distribute
from .base import Base, base
from .mf import MF, mf
from .ncf import MLP, NMF, mlp, nmf
from .ae import AE, Encoder, Decoder, ae
from .assist import Assist, assist
