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

import os as _os0

# Organizations run on private streams (one captured graph each). The default of 8 hardware work queues would alias
# 18 streams onto 8 queues and serialise them; must be set before the CUDA context exists.
_os0.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import os as _os
import sys as _sys

DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def use_dropin():
    """Put the host-side mirror of the reference interface first on sys.path and import it. Returns the three
    modules the reference's drivers import by these exact top-level names: (models, organization, assist)."""
    if DROPIN_DIR not in _sys.path:
        _sys.path.insert(0, DROPIN_DIR)
    import importlib

    mods = []
    for name in ("models", "organization", "assist"):
        m = _sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(DROPIN_DIR):
            raise ImportError("a different '{}' module is already imported from {}".format(name, m.__file__))
        mods.append(importlib.import_module(name))
    return tuple(mods)
