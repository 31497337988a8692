"""dmtcdr_b200 — B200-native hot path of DMTCDR (per-organization training + the MTAL round).

Layout (only what the path needs):
  csrc/      hand-written sm_100a CUDA kernels and the C-ABI (include/dmt_b200.h)
  native.py  ctypes binding of the C-ABI shared library (fails loudly if it is missing)
  dropin/    host-side mirror of the reference interface: ``models``, ``assist``, ``organization``
  engine.py  device-resident organization engine (batch assembly, train/predict rounds)
  dist.py    organization -> rank sharding and the per-round exchange
  synth.py   synthetic ML1M/Douban/Amazon-shaped inputs
"""
__version__ = "0.1.0"
