"""CPU oracle for the DMTCDR hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch restatement (torch-CPU fp32 / numpy, dense-matrix formulations) of the reference's
algorithm for per-organization training and the MTAL round. Every function cites the reference
file:line it follows. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or the timed CPU arm.
The product path (``dmtcdr_b200``) never imports it and fails loudly without its CUDA library.

Pinning: the reference has no tests or golden vectors of its own and its
arithmetic lives in PyTorch. The oracle is therefore pinned against outputs of the UNMODIFIED
reference run in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``):
model forward/loss/grads/optimizer steps, make_dataset residuals, every update() variant and whole
shortened experiments. ``tests/test_oracle_golden.py`` holds those checks.
"""
from . import models, mtal, train  # noqa: F401
